"""Import the real reference (google/TensorNetwork at /root/reference) in the BUILD
container only, through a 3-module import shim.

Test infrastructure.  `/root/reference` does not exist on the GPU box, so nothing that
runs there may call `load()`; it is used by oracle/gen_golden.py (fixture generation)
and by the `refhost` tests that check our adapter against the reference's own callers.
"""
import os
import sys
import types

REF_ROOT = "/root/reference"


def available() -> bool:
  return os.path.isdir(os.path.join(REF_ROOT, "tensornetwork"))


def load():
  """Returns the imported reference `tensornetwork` module."""
  if "tensornetwork" in sys.modules:
    return sys.modules["tensornetwork"]
  if not available():
    raise ImportError("reference not present at " + REF_ROOT)
  if "h5py" not in sys.modules:
    h5 = types.ModuleType("h5py")
    h5.Group = object
    h5.File = object
    h5.string_dtype = lambda encoding=None: object
    sys.modules["h5py"] = h5
  if "graphviz" not in sys.modules:
    gv = types.ModuleType("graphviz")
    gv.Graph = object
    sys.modules["graphviz"] = gv
  if "opt_einsum" not in sys.modules:
    from numpy._core.einsumfunc import _greedy_path, _optimal_path
    oe = types.ModuleType("opt_einsum")
    big = 2**62

    def greedy(i, o, s, memory_limit=None):
      return _greedy_path(i, o, s, big if memory_limit is None else memory_limit)

    def optimal(i, o, s, memory_limit=None):
      return _optimal_path(i, o, s, big if memory_limit is None else memory_limit)
    oe.paths = types.SimpleNamespace(greedy=greedy, optimal=optimal,
                                     dynamic_programming=optimal, branch=greedy)
    sys.modules["opt_einsum"] = oe
  if REF_ROOT not in sys.path:
    sys.path.insert(0, REF_ROOT)
  import tensornetwork  # pylint: disable=import-outside-toplevel
  return tensornetwork
