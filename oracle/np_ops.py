"""numpy `ops` adapter for the generic drivers (test infrastructure): presents oracle.np_backend /
oracle.np_network with the backend-like surface `tensornetwork_b200.dmrg.TwoSiteDMRG` expects, so the
SAME driver can be run on the CPU oracle and on the CUDA backend and compared (plus golden energies
from the real reference's FiniteDMRG)."""
import numpy as np
from . import np_backend as nb
from . import np_network as nn


class NumpyOps:
  name = "numpy-oracle"

  def convert_to_tensor(self, t):
    return np.array(t)

  def ncon(self, tensors, net):
    return nn.ncon(list(tensors), net)

  def conj(self, t):
    return np.conj(t)

  def norm(self, t):
    return nb.norm(t)

  def qr(self, t, pivot_axis=-1, non_negative_diagonal=False):
    return nb.qr(t, pivot_axis, non_negative_diagonal)

  def rq(self, t, pivot_axis=-1, non_negative_diagonal=False):
    return nb.rq(t, pivot_axis, non_negative_diagonal)

  def svd(self, t, pivot_axis=-1, max_singular_values=None, max_truncation_error=None, relative=False):
    return nb.svd(t, pivot_axis, max_singular_values, max_truncation_error, relative)

  def diagflat(self, t, k=0):
    return nb.diagflat(t, k)

  def ones(self, shape, dtype=None):
    return np.ones(shape, dtype=dtype if dtype is not None else np.float64)

  def eigsh_lanczos(self, **kw):
    return nb.eigsh_lanczos(**kw)
