"""Two-site DMRG driver (a14): CPU — the driver on the numpy oracle reproduces the reference's
FiniteDMRG energies (golden, from the real reference) and exact diagonalisation (dmrg_test.py:161-191);
GPU — the same driver on the CUDA backend reproduces them too."""
import numpy as np
import pytest
from conftest import load_golden
from tensornetwork_b200 import dmrg


def _case(z, ci, m):
  mps = [z["c%d_mps%d" % (ci, j)] for j in range(m["N"])]
  mpo = [z["c%d_mpo%d" % (ci, j)] for j in range(m["N"])]
  return mps, mpo


def test_xxz_mpo_matches_reference_tensors():
  meta, z = load_golden("dmrg")
  for ci, m in enumerate(meta):
    _, ref = _case(z, ci, m)
    mine = dmrg.xxz_mpo(np.ones(m["N"] - 1), np.ones(m["N"] - 1), np.zeros(m["N"]))
    for a, b in zip(ref, mine):
      np.testing.assert_array_equal(a, b)


def test_driver_on_oracle_matches_reference_energy():
  from oracle.np_ops import NumpyOps
  meta, z = load_golden("dmrg")
  for ci, m in enumerate(meta):
    mps, mpo = _case(z, ci, m)
    drv = dmrg.TwoSiteDMRG(NumpyOps(), mps, mpo, center_position=m["center"])
    e = drv.run_two_site(max_bond_dim=m["D"], num_sweeps=m["sweeps"], num_krylov_vecs=10)
    assert abs(e - m["energy"]) < 1e-10, (ci, e, m["energy"])
    assert abs(e - m["ed"]) < 1e-5


@pytest.mark.gpu
def test_driver_on_cuda_backend_matches_reference_energy():
  import tensornetwork_b200 as tb
  be = tb.get_backend()
  meta, z = load_golden("dmrg")
  for ci, m in enumerate(meta):
    mps, mpo = _case(z, ci, m)
    drv = dmrg.TwoSiteDMRG(dmrg.BackendOps(be), mps, mpo, center_position=m["center"])
    e = drv.run_two_site(max_bond_dim=m["D"], num_sweeps=m["sweeps"], num_krylov_vecs=10)
    assert abs(e - m["energy"]) < 1e-8, (ci, e, m["energy"])
    assert abs(e - m["ed"]) < 1e-5
