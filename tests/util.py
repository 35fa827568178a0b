"""Shared helpers for the parity tests (CUDA path vs oracle / golden vectors)."""
import numpy as np

# relative Frobenius tolerances per dtype (north_star: fp64 <= 1e-10; lower precisions stated here)
TOL = {"float64": 1e-10, "complex128": 1e-10, "float32": 2e-5, "complex64": 2e-5,
       "float16": 4e-3, "bfloat16": 2e-2, "tf32": 2e-3}


def rel_err(x, ref):
  x = np.asarray(x).astype(np.complex128 if np.iscomplexobj(ref) or np.iscomplexobj(x) else np.float64)
  ref = np.asarray(ref).astype(x.dtype)
  assert x.shape == ref.shape, (x.shape, ref.shape)
  d = np.linalg.norm((x - ref).ravel())
  n = np.linalg.norm(ref.ravel())
  return d / n if n > 0 else d


def assert_close(x, ref, dtype=None, tol=None, what=""):
  ref = np.asarray(ref)
  host = x.to_host() if hasattr(x, "to_host") else np.asarray(x)
  assert host.shape == ref.shape, "{} shape {} vs {}".format(what, host.shape, ref.shape)
  if ref.dtype.kind in "iu":
    np.testing.assert_array_equal(host, ref)
    return
  if tol is None:
    tol = TOL[dtype or str(ref.dtype)]
  e = rel_err(host, ref)
  assert e <= tol, "{}: rel err {:.3e} > {:.1e}".format(what, e, tol)


def get_backend():
  import tensornetwork_b200 as tb
  return tb.get_backend()
