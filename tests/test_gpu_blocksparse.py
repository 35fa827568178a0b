"""GPU parity of the grouped block-sparse tensordot (a11) vs the reference's own result vectors
(golden) and vs dense np.tensordot (the reference's oracle style, block_sparse/tensordot_test.py:171-187)."""
import numpy as np
import pytest
from conftest import load_golden
from util import get_backend, rel_err
from tensornetwork_b200 import blocksparse as bs

pytestmark = pytest.mark.gpu


def test_golden_blocksparse_tensordot():
  be = get_backend()
  meta, z = load_golden("blocksparse")
  for ci, m in enumerate(meta):
    legs = [bs.Index(z["c%d_q%d" % (ci, li)], f) for li, f in enumerate(m["flows"])]
    A = bs.BlockSparseTensor.from_data(z["c%d_A" % ci], legs, backend=be)
    At = A if m["perm"] is None else A.transpose(m["perm"])
    l0 = be.lib.tnb200_launch_count()
    C = bs.tensordot(At, At.conj(), m["axes"])
    ref = z["c%d_C" % ci]
    got = C.data.to_host()
    assert got.shape == ref.shape, "case %d: data vector length (layout) differs" % ci
    assert rel_err(got, ref) < 1e-12, "case %d" % ci
    if z["c%d_Cdense" % ci].size:
      np.testing.assert_allclose(C.todense(), z["c%d_Cdense" % ci], atol=1e-12)
    assert be.lib.tnb200_last_kernel().decode() == "blocksparse_grouped"


@pytest.mark.parametrize("dtype", ["float64", "float32", "complex128"])
def test_dense_equivalence_random(dtype):
  be = get_backend()
  rng = np.random.default_rng(41)
  q = lambda n: rng.integers(-2, 3, size=n)
  la = [bs.Index(q(6), False), bs.Index(q(5), True), bs.Index(q(7), False), bs.Index(q(4), True)]
  lb = [la[2].flip_flow(), bs.Index(q(6), False), la[1].flip_flow(), bs.Index(q(3), True)]
  da = rng.standard_normal([l.dim for l in la]).astype(dtype)
  db = rng.standard_normal([l.dim for l in lb]).astype(dtype)
  if dtype.startswith("complex"):
    da = da + 1j * rng.standard_normal(da.shape)
    db = db - 1j * rng.standard_normal(db.shape)
  A = bs.BlockSparseTensor.fromdense(la, da, backend=be)
  B = bs.BlockSparseTensor.fromdense(lb, db, backend=be)
  C = bs.tensordot(A, B, ([2, 1], [0, 2]))
  ref = np.tensordot(A.todense(), B.todense(), ([2, 1], [0, 2]))
  assert C.shape == ref.shape
  tol = 1e-5 if dtype == "float32" else 1e-12
  assert rel_err(C.todense(), ref) < tol
  # transposed operands + full inner product
  At = A.transpose((3, 0, 2, 1))
  C2 = bs.tensordot(At, B, ([2, 3], [0, 2]))
  assert rel_err(C2.todense(), np.tensordot(At.todense(), B.todense(), ([2, 3], [0, 2]))) < tol
  ip = bs.tensordot(A, A.conj(), ([0, 1, 2, 3], [0, 1, 2, 3]))
  assert abs(ip.todense() - np.vdot(A.todense(), A.todense())) < tol * abs(np.vdot(A.todense(), A.todense())) + 1e-12
  with pytest.raises(ValueError):
    bs.tensordot(A, B, ([2, 1], [0, 1]))
