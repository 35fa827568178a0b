"""TEST DOUBLE for libtnb200.so — lets the *host logic* of the cuda_b200 adapter (registration in the
reference's backend factory, axes bookkeeping, views, dtype promotion, error translation, Lanczos /
split control flow) run in the GPU-less build container against the REAL reference callers
(`tn.Node`, `tn.ncon`, `contractors.greedy`, `split_node*`, `FiniteDMRG`).

It implements the C-ABI entry points of include/tnb200.h on HOST memory with the numpy oracle.  It lives
under tests/ and is never importable from the product package; the product has no CPU path."""
import ctypes
import numpy as np
from oracle import np_backend as nb

_NP = {0: np.float64, 1: np.float32, 2: np.float16, 4: np.complex64, 5: np.complex128, 6: np.int32, 7: np.int64}


def _desc(arg):
  return arg._obj if hasattr(arg, "_obj") else arg


def _view(arg):
  d = _desc(arg)
  dt = np.dtype(_NP[d.dtype])
  nd = d.ndim
  shape = tuple(d.shape[i] for i in range(nd))
  strides = tuple(d.stride[i] * dt.itemsize for i in range(nd))
  if any(s == 0 for s in shape):
    return np.zeros(shape, dtype=dt)
  span = sum((s - 1) * abs(st) for s, st in zip(shape, strides)) + dt.itemsize
  buf = (ctypes.c_char * span).from_address(d.data)
  return np.ndarray(shape, dtype=dt, buffer=buf, strides=strides)


def _scalar_at(ptr, code):
  dt = np.dtype(_NP[code])
  return np.ndarray((), dtype=dt, buffer=(ctypes.c_char * dt.itemsize).from_address(ptr))


class FakeLib:
  """same callables as the ctypes library object"""

  def __init__(self):
    self._err = b""
    self._kernel = b"fake"
    self._launches = 0

  def _fail(self, code, msg):
    self._err = msg.encode()
    return code

  def tnb200_last_error(self):
    return self._err

  def tnb200_last_kernel(self):
    return self._kernel

  def tnb200_abi_version(self):
    return 1

  def tnb200_launch_count(self):
    return self._launches

  def tnb200_tensordot(self, a, b, c, naxes, axes_a, axes_b, nbatch, batch_a, batch_b, flags, stream):
    A, B, C = _view(a), _view(b), _view(c)
    ax_a = [axes_a[i] for i in range(naxes)]
    ax_b = [axes_b[i] for i in range(naxes)]
    ba = [batch_a[i] for i in range(nbatch)]
    bb = [batch_b[i] for i in range(nbatch)]
    for x, y in zip(ax_a, ax_b):
      if A.shape[x] != B.shape[y]:
        return self._fail(-1, "shape-mismatch for sum")
    if flags & 1:
      A = np.conj(A)
    if flags & 2:
      B = np.conj(B)
    self._launches += 1
    if nbatch == 0:
      C[...] = np.tensordot(A, B, (ax_a, ax_b))
      return 0
    fa = [i for i in range(A.ndim) if i not in ax_a and i not in ba]
    fb = [i for i in range(B.ndim) if i not in ax_b and i not in bb]
    L = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOP"
    sa, sb = [None] * A.ndim, [None] * B.ndim
    n = 0
    for x, y in zip(ba, bb):
      sa[x] = sb[y] = L[n]; n += 1
    for x, y in zip(ax_a, ax_b):
      sa[x] = sb[y] = L[n]; n += 1
    for x in fa:
      sa[x] = L[n]; n += 1
    for y in fb:
      sb[y] = L[n]; n += 1
    out = "".join(sa[x] for x in ba) + "".join(sa[x] for x in fa) + "".join(sb[y] for y in fb)
    C[...] = np.einsum("".join(sa) + "," + "".join(sb) + "->" + out, A, B)
    return 0

  def tnb200_chain_create(self, nsteps, steps, first_unsupported, handle):
    self._err = b"chained launches need the CUDA library"
    return -4

  def tnb200_chain_launch(self, handle, stream):
    return -1

  def tnb200_chain_destroy(self, handle):
    return 0

  def tnb200_copy(self, src, dst, conj, stream):
    s, d = _view(src), _view(dst)
    self._launches += 1
    if np.iscomplexobj(s) and not np.iscomplexobj(d):
      s = s.real
    d[...] = np.conj(s) if conj else s
    return 0

  def tnb200_binary(self, op, a, b, c, stream):
    A, B, C = _view(a), _view(b), _view(c)
    self._launches += 1
    with np.errstate(all="ignore"):
      C[...] = [np.add, np.subtract, np.multiply, np.divide, np.power][op](A, B)
    return 0

  def tnb200_unary(self, op, a, c, stream):
    A, C = _view(a), _view(c)
    self._launches += 1
    f = [np.conj, np.sqrt, np.abs, np.negative, np.exp, np.log, np.sin, np.cos, np.sign, np.real, np.imag][op]
    with np.errstate(all="ignore"):
      C[...] = f(A)
    return 0

  def tnb200_affine_inplace(self, x, ar, ai, br, bi, stream):
    X = _view(x)
    self._launches += 1
    if np.iscomplexobj(X):
      X[...] = X * complex(ar, ai) + complex(br, bi)
    else:
      X[...] = X * ar + br
    return 0

  def tnb200_scale_by_device_scalar(self, x, alpha_ptr, alpha_dtype, power, stream):
    X = _view(x)
    s = _scalar_at(alpha_ptr, alpha_dtype)[()]
    self._launches += 1
    if not np.iscomplexobj(X):
      s = np.real(s)
    X[...] = X / s if power < 0 else X * s
    return 0

  def tnb200_axpy(self, x, y, ar, ai, alpha_ptr, sign, stream):
    X, Y = _view(x), _view(y)
    self._launches += 1
    if alpha_ptr:
      alpha = sign * _scalar_at(alpha_ptr, _desc(x).dtype)[()]
    else:
      alpha = complex(ar, ai) if np.iscomplexobj(X) else ar
    Y[...] = Y + alpha * X
    return 0

  def tnb200_fill(self, c, re, im, stream):
    C = _view(c)
    self._launches += 1
    C[...] = complex(re, im) if np.iscomplexobj(C) else re
    return 0

  def tnb200_eye(self, c, k, stream):
    C = _view(c)
    C[...] = np.eye(C.shape[0], C.shape[1], k=k, dtype=C.dtype)
    return 0

  def tnb200_randn(self, c, seed, stream):
    C = _view(c)
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(C.shape)
    if np.iscomplexobj(C):
      v = v + 1j * rng.standard_normal(C.shape)
    C[...] = v
    return 0

  def tnb200_uniform(self, c, lo, hi, seed, stream):
    C = _view(c)
    rng = np.random.default_rng(seed)
    v = rng.uniform(lo, hi, C.shape)
    if np.iscomplexobj(C):
      v = v + 1j * rng.uniform(lo, hi, C.shape)
    C[...] = v
    return 0

  def tnb200_norm(self, a, out, stream):
    A = _view(a)
    code = {5: 0, 4: 1}.get(_desc(a).dtype, _desc(a).dtype)
    _scalar_at(out, code)[...] = np.linalg.norm(A)
    return 0

  def tnb200_dot(self, x, y, conj_x, out, stream):
    X, Y = _view(x), _view(y)
    v = np.sum((np.conj(X) if conj_x else X) * Y)
    _scalar_at(out, _desc(x).dtype)[...] = v
    return 0

  def tnb200_sum(self, a, c, naxes, axes, stream):
    A, C = _view(a), _view(c)
    C[...] = np.sum(A, axis=tuple(axes[i] for i in range(naxes)))
    return 0

  def tnb200_trace(self, a, c, offset, axis1, axis2, stream):
    A, C = _view(a), _view(c)
    C[...] = np.trace(A, offset=offset, axis1=axis1, axis2=axis2)
    return 0

  def tnb200_diagflat(self, a, c, k, stream):
    A, C = _view(a), _view(c)
    C[...] = np.diagflat(A, k=k)
    return 0

  def tnb200_svd(self, a, u, s, vh, info, stream):
    A = _view(a)
    U, S, Vh = np.linalg.svd(A, full_matrices=False)
    _view(u)[...] = U
    _view(s)[...] = S
    _view(vh)[...] = Vh
    return 0

  def tnb200_svd_truncation_count(self, s, max_sv, use_err, max_err, relative, keep_ptr, stream):
    S = _view(s)
    keep = nb.truncation_count(S, None if max_sv < 0 else max_sv, max_err if use_err else None, bool(relative))
    np.ndarray((), dtype=np.int64, buffer=(ctypes.c_char * 8).from_address(keep_ptr))[...] = keep
    return 0

  def tnb200_qr(self, a, q, r, nonneg, stream):
    A = _view(a)
    Q, R = np.linalg.qr(A)
    if nonneg:
      ph = np.sign(np.diagonal(R))
      Q = Q * ph
      R = ph.conj()[:, None] * R
    _view(q)[...] = Q
    _view(r)[...] = R
    return 0

  # ---- block-sparse entry points (host memory): raw pointers + element counts
  @staticmethod
  def _vec(ptr, n, dt):
    dt = np.dtype(dt)
    if n == 0:
      return np.zeros(0, dtype=dt)
    return np.ndarray((n,), dtype=dt, buffer=(ctypes.c_char * (n * dt.itemsize)).from_address(int(ptr)))

  def tnb200_gather(self, src, idx, dst, n, dtype, scatter, stream):
    n = int(n)
    if n == 0:
      return 0
    ix = self._vec(idx, n, np.int64)
    hi = int(ix.max()) + 1
    if scatter:
      self._vec(dst, hi, _NP[dtype])[ix] = self._vec(src, n, _NP[dtype])
    else:
      self._vec(dst, n, _NP[dtype])[...] = self._vec(src, hi, _NP[dtype])[ix]
    self._launches += 1
    return 0

  def tnb200_blocksparse_tensordot(self, a, b, c, dtype, nsect, dims, am, ao, bm, bo, cm, co, max_m, max_n, conj_b, stream):
    nsect = int(nsect)
    d = self._vec(dims, 3 * nsect, np.int64).reshape(nsect, 3)
    aoff, boff, coff = (self._vec(p, nsect + 1, np.int64) for p in (ao, bo, co))
    amap, bmap, cmap = self._vec(am, int(aoff[-1]), np.int64), self._vec(bm, int(boff[-1]), np.int64), self._vec(cm, int(coff[-1]), np.int64)
    A = self._vec(a, int(amap.max()) + 1, _NP[dtype])
    B = self._vec(b, int(bmap.max()) + 1, _NP[dtype])
    C = self._vec(c, int(cmap.max()) + 1, _NP[dtype])
    for q in range(nsect):
      m, k, n = (int(x) for x in d[q])
      x = A[amap[aoff[q]:aoff[q + 1]]].reshape(m, k)
      y = B[bmap[boff[q]:boff[q + 1]]].reshape(k, n)
      C[cmap[coff[q]:coff[q + 1]]] = (x @ (np.conj(y) if conj_b else y)).ravel()
    self._launches += 1
    return 0

  def tnb200_svd_batched(self, a, dtype, nprob, dims, aoff, u, uoff, s, soff, vh, voff, max_m, max_n, status, stream):
    nprob = int(nprob)
    d = self._vec(dims, 2 * nprob, np.int64).reshape(nprob, 2)
    ao, uo, so, vo = (self._vec(p, nprob + 1, np.int64) for p in (aoff, uoff, soff, voff))
    rdt = np.zeros(0, dtype=_NP[dtype]).real.dtype
    A, U = self._vec(a, int(ao[-1]), _NP[dtype]), self._vec(u, int(uo[-1]), _NP[dtype])
    S, V = self._vec(s, int(so[-1]), rdt), self._vec(vh, int(vo[-1]), _NP[dtype])
    for q in range(nprob):
      m, n = int(d[q, 0]), int(d[q, 1])
      uu, ss, vv = np.linalg.svd(A[ao[q]:ao[q + 1]].reshape(m, n), full_matrices=False)
      U[uo[q]:uo[q + 1]] = uu.ravel()
      S[so[q]:so[q + 1]] = ss
      V[vo[q]:vo[q + 1]] = vv.ravel()
    if status:
      self._vec(status, 1, np.int32)[0] = 0
    self._launches += 1
    return 0

  def tnb200_blocksparse_maps(self, nlegs, dims, charges, leg_off, order, partition, split, modulus, shift, nbins, tables, nnz, map_out, stream):
    """numpy transcription of csrc/blocksparse_maps.cu (same five stages: fuse, rank, first, bucket, element)"""
    nlegs, partition, split, modulus, shift, nbins, nnz = int(nlegs), int(partition), int(split), int(modulus), int(shift), int(nbins), int(nnz)
    dims = [int(dims[i]) for i in range(nlegs)]
    leg_off = [int(leg_off[i]) for i in range(nlegs)]
    order = [int(order[i]) for i in range(nlegs)]
    ch = self._vec(charges, sum(dims), np.int64)
    tab = self._vec(tables, 3 * nbins, np.int64)
    start_right, sect_off, ncols = tab[:nbins], tab[nbins:2 * nbins], tab[2 * nbins:]
    if nnz == 0:
      return 0
    out = self._vec(map_out, nnz, np.int64)

    def digits(legs):
      shape = [dims[t] for t in legs] or [1]
      idx = np.indices(shape).reshape(len(shape), -1) if legs else np.zeros((0, 1), dtype=np.int64)
      return idx

    def fuse(legs):
      idx = digits(legs)
      q = np.zeros(idx.shape[1] if legs else 1, dtype=np.int64)
      for k, t in enumerate(legs):
        q += ch[leg_off[t] + idx[k]]
      return (np.mod(q, modulus) if modulus else q + shift).astype(np.int64)

    def rank(b):
      r = np.zeros(b.shape[0], dtype=np.int64)
      cnt = np.zeros(nbins, dtype=np.int64)
      for v in range(nbins):
        m = b == v
        r[m] = np.arange(int(m.sum()))
        cnt[v] = m.sum()
      return r, cnt
    stored = list(range(nlegs))
    L_, R_ = stored[:split], stored[split:]
    bl, br = fuse(L_), fuse(R_)
    bro, bco = fuse(order[:partition]), fuse(order[partition:])
    rr, cr = rank(br)
    rro, _ = rank(bro)
    rco, _ = rank(bco)
    pb = (modulus - bl) % modulus if modulus else 2 * shift - bl
    first = np.zeros(bl.shape[0] + 1, dtype=np.int64)
    first[1:] = np.cumsum(cr[pb])
    assert first[-1] == nnz
    bucket = np.zeros(br.shape[0], dtype=np.int64)
    bucket[start_right[br] + rr] = np.arange(br.shape[0])
    e = np.arange(nnz)
    l = np.searchsorted(first, e, side="right") - 1
    j = e - first[l]
    r = bucket[start_right[pb[l]] + j]
    row_mul, col_mul, is_row = [0] * nlegs, [0] * nlegs, [0] * nlegs
    m = 1
    for i in range(partition - 1, -1, -1):
      row_mul[order[i]] = m; is_row[order[i]] = 1; m *= dims[order[i]]
    m = 1
    for i in range(nlegs - 1, partition - 1, -1):
      col_mul[order[i]] = m; m *= dims[order[i]]
    Rr = np.zeros(nnz, dtype=np.int64); Cc = np.zeros(nnz, dtype=np.int64); rq = np.zeros(nnz, dtype=np.int64)
    for legs, state in ((L_, l), (R_, r)):
      rem = state.copy()
      for t in reversed(legs):
        d = rem % dims[t]; rem //= dims[t]
        Rr += d * row_mul[t]; Cc += d * col_mul[t]
        if is_row[t]:
          rq += ch[leg_off[t] + d]
    qb = np.mod(rq, modulus) if modulus else rq + shift
    out[sect_off[qb] + rro[Rr] * ncols[qb] + rco[Cc]] = e
    self._launches += 10
    return 0
