"""numpy restatement of the reference numpy backend (test oracle, not product).

Each function cites the reference file:line (relative to /root/reference/) whose
behaviour it restates.  Arithmetic is plain numpy (the same library the reference
numpy backend calls), so at equal dtype results are bit-comparable with the
reference; see oracle/__init__.py for how this file is pinned.
"""
import numpy as np

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def tensordot(a, b, axes):
  """tensornetwork/backends/numpy/numpy_backend.py:35-54.

  Output axis order: free axes of `a` (original order) then free axes of `b`.
  When every axis of both operands is contracted the reference switches to an
  einsum and returns a 0-d array; the value equals np.tensordot's.
  """
  a = np.asarray(a)
  b = np.asarray(b)
  if not isinstance(axes, (int, np.integer)):
    if len(axes[0]) == a.ndim and len(axes[1]) == b.ndim:
      if len(axes[0]) != len(axes[1]):
        raise ValueError("shape-mismatch for sum")
      sub_a = [None] * a.ndim
      sub_b = [None] * b.ndim
      for n, (i, j) in enumerate(zip(axes[0], axes[1])):
        sub_a[i] = _LETTERS[n]
        sub_b[j] = _LETTERS[n]
      return np.array(np.einsum("".join(sub_a) + "," + "".join(sub_b), a, b,
                                optimize=True))
  return np.tensordot(a, b, axes)


def reshape(tensor, shape):
  """numpy_backend.py:56-57 — the target shape is cast to int32."""
  return np.reshape(tensor, np.asarray(shape).astype(np.int32))


def transpose(tensor, perm=None):
  """numpy_backend.py:59-62 — perm=None reverses the axes."""
  return np.transpose(tensor, perm)


def shape_concat(values, axis):
  """numpy_backend.py:74-75."""
  return np.concatenate(values, axis)


def shape_prod(values):
  """numpy_backend.py:86-87."""
  return np.prod(values)


def outer_product(t1, t2):
  """numpy_backend.py:99-100."""
  return np.tensordot(t1, t2, 0)


def einsum(expression, *tensors, optimize=True):
  """numpy_backend.py:102-106."""
  return np.einsum(expression, *tensors, optimize=optimize)


def norm(tensor):
  """numpy_backend.py:108-109 (Frobenius norm of the flattened tensor)."""
  return np.linalg.norm(tensor)


def trace(tensor, offset=0, axis1=-2, axis2=-1):
  """numpy_backend.py:684-707."""
  return np.trace(tensor, offset=offset, axis1=axis1, axis2=axis2)


def tsum(tensor, axis=None, keepdims=False):
  """numpy_backend.py:603-607 (`sum`)."""
  return np.sum(tensor, axis=None if axis is None else tuple(axis),
                keepdims=keepdims)


def matmul(t1, t2):
  """numpy_backend.py:609-612 — batched `...ab,...bc`; order-1 inputs rejected."""
  if t1.ndim <= 1 or t2.ndim <= 1:
    raise ValueError("inputs to `matmul` have to be a tensors of order > 1,")
  return np.matmul(t1, t2)


def diagflat(tensor, k=0):
  """numpy_backend.py:673-682."""
  return np.diagflat(tensor, k=k)


def diagonal(tensor, offset=0, axis1=-2, axis2=-1):
  """numpy_backend.py:643-671."""
  return np.diagonal(tensor, offset=offset, axis1=axis1, axis2=axis2)


def broadcast_right_multiplication(t1, t2):
  """numpy_backend.py:560-565."""
  if len(t2.shape) != 1:
    raise ValueError("only order-1 tensors are allowed for `tensor2`")
  return t1 * t2


def broadcast_left_multiplication(t1, t2):
  """numpy_backend.py:567-575."""
  if len(t1.shape) != 1:
    raise ValueError("only order-1 tensors are allowed for `tensor1`")
  return t2 * np.reshape(t1, t1.shape + (1,) * (t2.ndim - 1))


def svd(tensor, pivot_axis=-1, max_singular_values=None,
        max_truncation_error=None, relative=False):
  """tensornetwork/backends/numpy/decompositions.py:21-74.

  Returns (u, s, vh, s_rest).  The kept count is
  min(max_singular_values, #{ascending cumulative norms > eps}); `s` is cast to
  the input dtype; all discarded singular values are returned.
  """
  left_dims = tensor.shape[:pivot_axis]
  right_dims = tensor.shape[pivot_axis:]
  mat = np.reshape(tensor, [int(np.prod(left_dims)), int(np.prod(right_dims))])
  u, s, vh = np.linalg.svd(mat, full_matrices=False)
  keep = truncation_count(s, max_singular_values, max_truncation_error,
                          relative)
  s = s.astype(mat.dtype)
  s_rest = s[keep:]
  s = s[:keep]
  u = u[:, :keep]
  vh = vh[:keep, :]
  dim_s = s.shape[0]
  u = np.reshape(u, list(left_dims) + [dim_s])
  vh = np.reshape(vh, [dim_s] + list(right_dims))
  return u, s, vh, s_rest


def truncation_count(s, max_singular_values=None, max_truncation_error=None,
                     relative=False):
  """decompositions.py:38-57 — integer output, must match bit-exactly."""
  s = np.asarray(s)
  if max_singular_values is None:
    max_singular_values = s.size
  if max_truncation_error is not None:
    trunc_errs = np.sqrt(np.cumsum(np.square(s[::-1])))
    eps = max_truncation_error * s[0] if relative else max_truncation_error
    by_err = int(np.count_nonzero((trunc_errs > eps).astype(np.int32)))
  else:
    by_err = max_singular_values
  return int(min(max_singular_values, by_err))


def qr(tensor, pivot_axis=-1, non_negative_diagonal=False):
  """decompositions.py:77-98."""
  left_dims = tensor.shape[:pivot_axis]
  right_dims = tensor.shape[pivot_axis:]
  mat = np.reshape(tensor, [int(np.prod(left_dims)), int(np.prod(right_dims))])
  q, r = np.linalg.qr(mat)
  if non_negative_diagonal:
    phases = np.sign(np.diagonal(r))
    q = q * phases
    r = phases.conj()[:, None] * r
  center = q.shape[1]
  return (np.reshape(q, list(left_dims) + [center]),
          np.reshape(r, [center] + list(right_dims)))


def rq(tensor, pivot_axis=-1, non_negative_diagonal=False):
  """decompositions.py:101-124 — QR of the conjugate transpose."""
  left_dims = tensor.shape[:pivot_axis]
  right_dims = tensor.shape[pivot_axis:]
  mat = np.reshape(tensor, [int(np.prod(left_dims)), int(np.prod(right_dims))])
  q, r = np.linalg.qr(np.conj(np.transpose(mat)))
  if non_negative_diagonal:
    phases = np.sign(np.diagonal(r))
    q = q * phases
    r = phases.conj()[:, None] * r
  r, q = np.conj(np.transpose(r)), np.conj(np.transpose(q))
  center = r.shape[1]
  return (np.reshape(r, list(left_dims) + [center]),
          np.reshape(q, [center] + list(right_dims)))


def eigsh_lanczos(A, args=None, initial_state=None, shape=None, dtype=None,
                  num_krylov_vecs=20, numeig=1, tol=1e-8, delta=1e-8, ndiag=20,
                  reorthogonalize=False):
  """numpy_backend.py:415-534 — host-driven Lanczos, same control flow."""
  if args is None:
    args = []
  if num_krylov_vecs < numeig:
    raise ValueError("`num_krylov_vecs` >= `numeig` required!")
  if numeig > 1 and not reorthogonalize:
    raise ValueError("Use `reorthogonalize=True` for `numeig > 1`")
  if initial_state is None:
    if shape is None or dtype is None:
      raise ValueError("if no `initial_state` is passed, then `shape` and"
                       "`dtype` have to be provided")
    initial_state = np.random.randn(*shape).astype(dtype)
  if not isinstance(initial_state, np.ndarray):
    raise TypeError("Expected a `np.ndarray`. Got {}".format(
        type(initial_state)))
  vector_n = np.array(initial_state)
  vector_n = vector_n / np.linalg.norm(vector_n)
  norms, diags, kvecs = [], [], []
  first = True
  eigvalsold = []
  for it in range(num_krylov_vecs):
    nrm = np.linalg.norm(vector_n)
    if abs(nrm) < delta:
      break
    norms.append(nrm)
    vector_n = vector_n / norms[-1]
    if reorthogonalize:
      for v in kvecs:
        vector_n = vector_n - np.dot(np.ravel(np.conj(v)),
                                     np.ravel(vector_n)) * v
    kvecs.append(vector_n)
    Av = A(vector_n, *args)
    diags.append(np.dot(np.ravel(np.conj(vector_n)), np.ravel(Av)))
    if it > 0 and it % ndiag == 0 and len(diags) >= numeig:
      T = (np.diag(diags) + np.diag(norms[1:], 1) +
           np.diag(np.conj(norms[1:]), -1))
      eigvals, _ = np.linalg.eigh(T)
      if not first:
        if np.linalg.norm(eigvals[0:numeig] - eigvalsold[0:numeig]) < tol:
          break
      first = False
      eigvalsold = eigvals[0:numeig]
    Av = Av - kvecs[-1] * diags[-1]
    if it > 0:
      Av = Av - kvecs[-2] * norms[-1]
    vector_n = Av
  T = np.diag(diags) + np.diag(norms[1:], 1) + np.diag(np.conj(norms[1:]), -1)
  eigvals, u = np.linalg.eigh(T)
  eigvals = np.array(eigvals).astype(T.dtype)
  eigvecs = []
  for n2 in range(min(numeig, len(eigvals))):
    state = np.zeros(initial_state.shape, initial_state.dtype)
    for n1, vec in enumerate(kvecs):
      state = state + vec * u[n1, n2]
    eigvecs.append(state / np.linalg.norm(state))
  return eigvals[0:numeig], eigvecs
