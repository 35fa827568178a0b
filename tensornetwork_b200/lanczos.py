"""Lanczos eigensolver on device tensors — NumPyBackend.eigsh_lanczos
(backends/numpy/numpy_backend.py:415-534) with the same host control flow; the vector
arithmetic (norm / dot / axpy / scale) runs in libtnb200 kernels.  Host syncs: one scalar
read per iteration for the `norm < delta` break test (the reference reads it too) and one
batched read of the alpha/beta coefficients whenever the tridiagonal matrix is needed."""
import numpy as np
from .tensor import B200Tensor
from . import _lib as L
from . import tensor as T


def eigsh_lanczos(be, A, args=None, initial_state=None, shape=None, dtype=None,
                  num_krylov_vecs=20, numeig=1, tol=1e-8, delta=1e-8, ndiag=20,
                  reorthogonalize=False):
  if args is None:
    args = []
  if num_krylov_vecs < numeig:
    raise ValueError('`num_krylov_vecs` >= `numeig` required!')
  if numeig > 1 and not reorthogonalize:
    raise ValueError("Got numeig = {} > 1 and `reorthogonalize = False`. "
                     "Use `reorthogonalize=True` for `numeig > 1`".format(numeig))
  if initial_state is None:
    if shape is None or dtype is None:
      raise ValueError("if no `initial_state` is passed, then `shape` and"
                       "`dtype` have to be provided")
    initial_state = be.randn(shape, dtype)
  if not isinstance(initial_state, B200Tensor):
    raise TypeError("Expected a `B200Tensor`. Got {}".format(type(initial_state)))
  torch = be.torch
  code = initial_state.code
  cplx = T.is_complex_code(code)

  vector_n = be.copy(initial_state)
  vector_n /= be.norm(vector_n)
  norms_dev, diags_dev, krylov = [], [], []
  first = True
  eigvalsold = []

  def host_coeffs():
    d = torch.stack([x.t.reshape(()) for x in diags_dev]).cpu().numpy()
    n = torch.stack([x.t.reshape(()) for x in norms_dev]).cpu().numpy().astype(
        np.float64 if code in (L.F64, L.C128) else np.float32)
    return d, n

  def tridiag():
    d, n = host_coeffs()
    return np.diag(d) + np.diag(n[1:], 1) + np.diag(np.conj(n[1:]), -1)

  for it in range(num_krylov_vecs):
    nrm = be.norm(vector_n)
    if abs(float(nrm.item())) < delta:
      break
    norms_dev.append(nrm)
    v = be.copy(vector_n)
    v /= nrm
    vector_n = v
    if reorthogonalize:
      for kv in krylov:
        ov = be.vdot(kv, vector_n, conj_x=True)
        be.axpy_dev(vector_n, kv, ov, sign=-1.0)
    krylov.append(vector_n)
    A_vector_n = A(vector_n, *args)
    diags_dev.append(be.vdot(vector_n, A_vector_n, conj_x=True))
    if (it > 0) and (it % ndiag == 0) and (len(diags_dev) >= numeig):
      eigvals, _ = np.linalg.eigh(tridiag())
      if not first:
        if np.linalg.norm(eigvals[0:numeig] - eigvalsold[0:numeig]) < tol:
          break
      first = False
      eigvalsold = eigvals[0:numeig]
    # the matvec result may alias caller state: work on our own copy
    A_vector_n = be.copy(A_vector_n)
    be.axpy_dev(A_vector_n, krylov[-1], diags_dev[-1], sign=-1.0)
    if it > 0:
      be.axpy_dev(A_vector_n, krylov[-2], norms_dev[-1], sign=-1.0)
    vector_n = A_vector_n

  A_tridiag = tridiag()
  eigvals, u = np.linalg.eigh(A_tridiag)
  eigvals = np.array(eigvals).astype(A_tridiag.dtype)
  eigenvectors = []
  for n2 in range(min(numeig, len(eigvals))):
    state = be.zeros(initial_state.shape, initial_state.dtype)
    for n1, vec in enumerate(krylov):
      c = u[n1, n2]
      be.iadd(state, vec, complex(c) if cplx else float(np.real(c)))
    state /= be.norm(state)
    eigenvectors.append(state)
  return eigvals[0:numeig], eigenvectors
