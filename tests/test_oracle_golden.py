"""Pin the oracle (oracle/np_*.py) against golden vectors produced by the real reference
(oracle/gen_golden.py, run in the build container).  CPU only."""
import numpy as np
import pytest
from oracle import np_backend as nb
from oracle import np_network as nn
from conftest import load_golden


def _same(x, y, exact=True):
  x, y = np.asarray(x), np.asarray(y)
  assert x.shape == y.shape and x.dtype == y.dtype, (x.shape, y.shape, x.dtype, y.dtype)
  if exact:
    np.testing.assert_array_equal(x, y)
  else:
    tol = 1e-5 if x.dtype in (np.float32, np.complex64) else 1e-12
    np.testing.assert_allclose(x, y, rtol=tol, atol=tol)


def test_tensordot_bit_exact():
  meta, z = load_golden("tensordot")
  for i, m in enumerate(meta):
    a, b = z["a%d" % i], z["b%d" % i]
    if m["perm_a"] is not None:
      a = nb.transpose(a, m["perm_a"])
    if m["perm_b"] is not None:
      b = nb.transpose(b, m["perm_b"])
    _same(nb.tensordot(a, b, m["axes"]), z["out%d" % i])


def test_ncon_bit_exact():
  meta, z = load_golden("ncon")
  for i, m in enumerate(meta):
    ts = [z["c%d_t%d" % (i, j)] for j in range(m["n"])]
    _same(nn.ncon(ts, m["net"], m["con"], m["out"]), z["c%d_out" % i], exact=False)


def test_decompositions_bit_exact():
  meta, z = load_golden("decomp")
  for i, m in enumerate(meta):
    res = getattr(nb, m["kind"])(z["in%d" % i], **m["kwargs"])
    assert len(res) == m["nout"]
    for j, x in enumerate(res):
      _same(x, z["out%d_%d" % (i, j)])


def test_greedy_contraction():
  meta, z = load_golden("greedy")
  for ci, m in enumerate(meta):
    if m.get("open"):
      ts = [z["open_a"], z["open_b"], z["open_c"]]
      labels, out = m["labels"], m["out"]
      ref = z["open_out"]
    else:
      kets = [z["c%d_k%d" % (ci, j)] for j in range(m["L"])]
      ts = kets + [np.conj(k) for k in kets]
      labels, out = m["labels"], []
      ref = z["c%d_out" % ci]
    sizes = {l: t.shape[ax] for t, labs in zip(ts, labels) for ax, l in enumerate(labs)}
    path = nn.greedy_path(labels, out, sizes)
    res = nn.contract_path(ts, labels, path, out)
    tol = 1e-5 if ref.dtype == np.float32 else 1e-12
    np.testing.assert_allclose(res, ref, rtol=tol, atol=tol)
    assert np.asarray(res).shape == ref.shape


def test_lanczos():
  meta, z = load_golden("lanczos")
  for i, m in enumerate(meta):
    h, x0 = z["h%d" % i], z["x%d" % i]
    ev, vecs = nb.eigsh_lanczos(lambda x, mat: mat @ x, [h], x0.copy(),
                                num_krylov_vecs=m["num_krylov_vecs"], numeig=m["numeig"],
                                reorthogonalize=m["reorthogonalize"], ndiag=m["ndiag"])
    np.testing.assert_allclose(ev, z["ev%d" % i], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.stack(vecs), z["vec%d" % i], rtol=1e-9, atol=1e-9)


def test_path_known_answers():
  """contractors/opt_einsum_paths/path_calculation_test.py:40-93, the three `greedy` rows:
  gemm_network -> [(0,2),(0,1)], inner_network -> [(0,1),(0,1)],
  matrix_chain -> [(0,1),(0,2),(0,1)]."""
  def run(shapes, labels):
    sizes = {l: s[ax] for s, labs in zip(shapes, labels) for ax, l in enumerate(labs)}
    return nn.greedy_path(labels, [], sizes)
  assert run([(1, 2, 4), (1, 3), (2, 4, 3)],
             [["xy", "xz0", "xz1"], ["xy", "yz"], ["xz0", "xz1", "yz"]]) == [(0, 2), (0, 1)]
  assert run([(5, 2, 3, 4), (5, 3), (2, 4)],
             [["a", "b", "c", "d"], ["a", "c"], ["b", "d"]]) == [(0, 1), (0, 1)]
  d = [10, 8, 6, 4, 2]
  shapes = list(zip(d[:-1], d[1:]))
  labels = [["e%d" % i, "e%d" % (i + 1)] for i in range(4)]
  sizes = {"e%d" % i: d[i] for i in range(5)}
  assert nn.greedy_path(labels, ["e0", "e4"], sizes) == [(0, 1), (0, 2), (0, 1)]


def test_blocksparse_oracle_matches_reference_data_vectors():
  """oracle/np_blocksparse.py (cpu_baseline of the cfg-4 bench leg) against the real reference's result data
  vectors for the trailing-axes cases of tests/golden/blocksparse.npz (incl. cfg 4 itself)."""
  from oracle import np_blocksparse as nbs
  meta, z = load_golden("blocksparse")
  checked = 0
  for ci, m in enumerate(meta):
    n = len(m["axes"][0])
    trailing = list(range(m["nlegs"] - n, m["nlegs"]))
    if m["perm"] is not None or m["axes"][0] != trailing or m["axes"][1] != trailing:
      continue
    charges = [z["c%d_q%d" % (ci, li)] for li in range(m["nlegs"])]
    flows = list(m["flows"])
    a = z["c%d_A" % ci]
    assert a.shape[0] == nbs.num_nonzero(charges, flows)
    c, _, _ = nbs.tensordot_trailing(a, charges, flows, np.conj(a), charges, [not f for f in flows], n)
    np.testing.assert_allclose(c, z["c%d_C" % ci], rtol=1e-12, atol=1e-12)
    checked += 1
  assert checked >= 3
