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
    assert be.lib.tnb200_last_kernel().decode().startswith("blocksparse_grouped")


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


def test_golden_symmetric_svd():
  """backends/symmetric/decompositions.py:27-216: kept / discarded singular values (sector-major order,
  cross-sector truncation) equal the reference's; U S V reconstructs the reference's truncated tensor."""
  be = get_backend()
  meta, z = load_golden("symsvd")
  for ci, m in enumerate(meta):
    legs = [bs.Index(z["c%d_q%d" % (ci, li)], f) for li, f in enumerate(m["flows"])]
    A = bs.BlockSparseTensor.from_data(z["c%d_A" % ci], legs, backend=be)
    U, S, V, Sd = bs.svd(A, m["pivot"], **m["kwargs"])
    ref_s = z["c%d_S" % ci]
    got_s = S["values"].to_host()
    assert got_s.shape == ref_s.shape, "case %d kept count %s vs %s" % (ci, got_s.shape, ref_s.shape)   # integer: bit exact
    assert S["ktot"] == m["k"]
    np.testing.assert_allclose(got_s, ref_s, atol=1e-10 * max(1.0, ref_s.max() if ref_s.size else 1.0))
    np.testing.assert_allclose(np.sort(Sd), np.sort(z["c%d_Sdisc" % ci]), atol=1e-10)
    ud, vd = U.todense(), V.todense()
    assert ud.shape[-1] == m["k"] and vd.shape[0] == m["k"]
    rec = np.tensordot(ud * got_s, vd, 1)
    assert rel_err(rec, z["c%d_rec" % ci]) < 1e-9, "case %d" % ci
    # isometries
    k = m["k"]
    u2 = ud.reshape(-1, k)
    np.testing.assert_allclose(u2.T @ u2, np.eye(k), atol=1e-10)


def test_batched_small_svd_direct():
  """tnb200_svd_batched on ragged problems (tall, wide, odd sizes, complex) vs numpy."""
  import torch
  from tensornetwork_b200 import _lib as L
  be = get_backend()
  rng = np.random.default_rng(43)
  for dtype, code, rcode, tol in (("float64", L.F64, L.F64, 1e-11), ("complex128", L.C128, L.F64, 1e-11), ("float32", L.F32, L.F32, 2e-5)):
    shapes = [(5, 9), (9, 5), (1, 4), (7, 7), (33, 20), (20, 33), (2, 2), (64, 17)]
    mats = []
    for (m_, n_) in shapes:
      x = rng.standard_normal((m_, n_))
      if dtype.startswith("complex"):
        x = x + 1j * rng.standard_normal((m_, n_))
      mats.append(x.astype(dtype))
    rs = [min(s) for s in shapes]
    a_off = np.insert(np.cumsum([m_ * n_ for m_, n_ in shapes]), 0, 0)
    u_off = np.insert(np.cumsum([s[0] * r for s, r in zip(shapes, rs)]), 0, 0)
    s_off = np.insert(np.cumsum(rs), 0, 0)
    v_off = np.insert(np.cumsum([r * s[1] for s, r in zip(shapes, rs)]), 0, 0)
    a = be.convert_to_tensor(np.concatenate([x.ravel() for x in mats]))
    u = be._new((int(u_off[-1]),), code); sv = be._new((int(s_off[-1]),), rcode); vh = be._new((int(v_off[-1]),), code)
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64)).to(be.device)
    d = [up(np.array(shapes).ravel()), up(a_off), up(u_off), up(s_off), up(v_off)]
    st = torch.zeros(1, dtype=torch.int32, device=be.device)
    L.check(be.lib.tnb200_svd_batched(a.t.data_ptr(), code, len(shapes), d[0].data_ptr(), d[1].data_ptr(), u.t.data_ptr(), d[2].data_ptr(),
                                      sv.t.data_ptr(), d[3].data_ptr(), vh.t.data_ptr(), d[4].data_ptr(), 64, 33, st.data_ptr(), be._stream()))
    assert int(st.item()) == 0
    uh, sh, vhh = u.to_host(), sv.to_host(), vh.to_host()
    for q, (x, (m_, n_), r) in enumerate(zip(mats, shapes, rs)):
      U = uh[u_off[q]:u_off[q + 1]].reshape(m_, r); S = sh[s_off[q]:s_off[q + 1]]; Vh = vhh[v_off[q]:v_off[q + 1]].reshape(r, n_)
      ref = np.linalg.svd(x.astype(np.complex128 if dtype.startswith("complex") else np.float64), compute_uv=False)
      np.testing.assert_allclose(S, ref, atol=tol * ref[0])
      assert rel_err((U * S) @ Vh, x) < 50 * tol
      np.testing.assert_allclose(U.conj().T @ U, np.eye(r), atol=200 * tol)


def test_device_built_maps_are_bit_identical():
  """f3: tnb200_blocksparse_maps (fuse / rank / scan / bucket / element kernels) against the host construction, which
  tests/test_blocksparse_maps.py pins bit-exactly to the reference's own `_find_transposed_diagonal_sparse_blocks` maps:
  random U(1) and Z_N leg structures, every order / partition, and the BASELINE cfg 4 structure (4 legs x 32, 33 sectors)."""
  be = get_backend()
  rng = np.random.default_rng(1)
  cases = []
  for _ in range(60):
    n = int(rng.integers(1, 6))
    mod = [None, None, None, 2, 3, 4][rng.integers(0, 6)]
    idx = [bs.Index(rng.integers(-3, 4, rng.integers(1, 9)) if mod is None else rng.integers(0, mod, rng.integers(1, 9)),
                    bool(rng.integers(0, 2)), mod) for _ in range(n)]
    cases.append((idx, [int(x) for x in rng.permutation(n)], int(rng.integers(0, n + 1))))
  r5 = np.random.RandomState(5)
  cfg4 = [bs.Index(r5.randint(-8, 9, 32).astype(np.int64), f) for f in (False, False, True, True)]
  cases += [(cfg4, [0, 1, 2, 3], 2), (cfg4, [2, 0, 3, 1], 2), (cfg4, [3, 2, 1, 0], 1)]
  for idx, order, part in cases:
    bs._MAP_CACHE.clear()
    q1, d1, m1 = bs._sector_maps(idx, order, part)
    q2, d2, dm, off = bs._device_sector_maps(be, idx, order, part)
    flat = np.concatenate(m1) if m1 else np.zeros(0, dtype=np.int64)
    np.testing.assert_array_equal(q1, q2)
    np.testing.assert_array_equal(d1, d2)
    np.testing.assert_array_equal(flat, dm.cpu().numpy()[:flat.shape[0]])
    np.testing.assert_array_equal(off, np.cumsum([0] + [len(x) for x in m1]))
  bs._MAP_CACHE.clear()


def test_tutorial_sized_legs_run_without_dense_enumeration():
  """the reference tutorial's (100,101,102,103) legs: 1.06e8 dense states, ~6e6 stored elements — maps on the device,
  nothing of dense size is ever allocated; checked against per-sector dense matmul of two sectors."""
  be = get_backend()
  np.random.seed(10)
  legs = [bs.Index(np.random.randint(-5, 6, d), f) for d, f in zip((100, 101, 102, 103), (False, False, True, True))]
  A = bs.BlockSparseTensor.randn(legs, dtype=np.float64, seed=3, backend=be)
  C = bs.tensordot(A, A.conj(), ([2, 3], [2, 3]))
  assert be.lib.tnb200_last_kernel().decode().startswith("blocksparse_grouped")
  # property check: C = M M^T per sector is symmetric positive semi-definite; trace(C) = |A|^2
  qn, dims, dmap, off = bs._device_sector_maps(be, C.indices, [0, 1, 2, 3], 2)
  data = C.data.to_host()
  tr = 0.0
  for s in range(len(qn)):
    m_, n_ = int(dims[s, 0]), int(dims[s, 1])
    blk = data[dmap.cpu().numpy()[off[s]:off[s + 1]]].reshape(m_, n_)
    assert m_ == n_
    np.testing.assert_allclose(blk, blk.T, atol=1e-9 * max(1.0, np.abs(blk).max()))
    tr += np.trace(blk)
  a = A.data.to_host()
  assert abs(tr - float(a @ a)) <= 1e-10 * float(a @ a)
