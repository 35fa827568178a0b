"""`CudaB200Backend.jit`: whole-function CUDA-graph capture behind the reference's jit hook.

The reference routes every `tn.ncon` through `backend.jit(_jittable_ncon, static_argnums=(1,..,5))`
(ncon_interface.py:654-660; hook: abstract_backend.py:798, numpy's is the identity
numpy_backend.py:600-601).  Here the hook returns a `JitFunction`:

  call 1 of a (static args, tensor shapes/dtypes/strides) key   eager  (warms the library's lazy
                                                                 state: attributes, memory pools)
  call 2                                                         the SAME python function runs once
                                                                 more under CUDA stream capture on
                                                                 private copies of the inputs; the
                                                                 graph is replayed for the result
  call 3..                                                       copy the inputs into the captured
                                                                 buffers, ONE graph launch, clone
                                                                 the outputs (value semantics: the
                                                                 caller may keep them)

so the python label bookkeeping of `_jittable_ncon` and its per-step launches are paid once per
network structure.  Functions that synchronise with the host inside (truncating SVD: the kept
count is data dependent) cannot be captured: the failed capture marks the key eager for good.
Keys whose tensors exceed `max_elements` stay eager too (kernel time dominates there and the
captured intermediates would stay resident).
"""
import collections

from .tensor import B200Tensor


class CaptureUnsupported(RuntimeError):
  """raised by the adapter when an operation that must synchronise with the host is reached during graph capture"""


def _flatten(obj, out):
  """replaces every B200Tensor in a nest of lists/tuples by a slot index; returns the skeleton"""
  if isinstance(obj, B200Tensor):
    out.append(obj)
    return ("t", len(out) - 1)
  if isinstance(obj, (list, tuple)):
    return ("l" if isinstance(obj, list) else "u", tuple(_flatten(o, out) for o in obj))
  return ("c", obj)


def _unflatten(skel, tensors):
  kind, val = skel
  if kind == "t":
    return tensors[val]
  if kind == "l":
    return [_unflatten(s, tensors) for s in val]
  if kind == "u":
    return tuple(_unflatten(s, tensors) for s in val)
  return val


def _skeleton_key(skel):
  kind, val = skel
  if kind in ("l", "u"):
    return (kind, tuple(_skeleton_key(s) for s in val))
  if kind == "t":
    return ("t", val)
  hash(val)            # TypeError for unhashable constants -> the caller falls back to eager
  return ("c", val)


class _Entry:
  __slots__ = ("calls", "graph", "static_in", "static_out", "out_skel", "eager")

  def __init__(self):
    self.calls = 0
    self.graph = None
    self.static_in = None
    self.static_out = None
    self.out_skel = None
    self.eager = False


class JitFunction:
  """Callable returned by `CudaB200Backend.jit(fun, static_argnums=...)`."""

  def __init__(self, backend, fun, static_argnums=None, max_elements=1 << 22, max_entries=128):
    self.backend = backend
    self.fun = fun
    self.static_argnums = tuple(static_argnums or ())
    self.max_elements = max_elements
    self.max_entries = max_entries
    self.cache = collections.OrderedDict()
    self.stats = backend.jit_stats
    self.__name__ = getattr(fun, "__name__", "jitted")
    self.__doc__ = getattr(fun, "__doc__", None)

  def __call__(self, *args, **kwargs):
    be = self.backend
    if kwargs or not be._on_cuda or not be.jit_graphs or be._capturing:   # (inside another capture: just part of it)
      return self.fun(*args, **kwargs)
    tensors = []
    try:
      skel = tuple(("s", a) if i in self.static_argnums else _flatten(a, tensors) for i, a in enumerate(args))
      key = (tuple(("s", a) if k == "s" else _skeleton_key((k, a)) for k, a in skel),
             tuple((t.shape, t.code, tuple(t.t.stride())) for t in tensors))
      hash(key)
    except TypeError:
      self.stats["eager"] += 1
      return self.fun(*args)
    if not tensors or sum(t.size for t in tensors) > self.max_elements:
      self.stats["eager"] += 1
      return self.fun(*args)
    ent = self.cache.get(key)
    if ent is None:
      ent = self.cache[key] = _Entry()
      while len(self.cache) > self.max_entries:
        self.cache.popitem(last=False)
    else:
      self.cache.move_to_end(key)
    ent.calls += 1
    if ent.eager or ent.calls == 1:
      self.stats["eager"] += 1
      return self.fun(*args)
    torch = be.torch
    if ent.graph is None:
      if not self._capture(ent, skel, tensors):
        self.stats["eager"] += 1
        return self.fun(*args)
    else:
      for dst, src in zip(ent.static_in, tensors):
        dst.t.copy_(src.t)
    ent.graph.replay()
    self.stats["replays"] += 1
    outs = [B200Tensor(o.t.clone(memory_format=torch.preserve_format), o.code) for o in ent.static_out]
    return _unflatten(ent.out_skel, outs)

  def _capture(self, ent, skel, tensors):
    be, torch = self.backend, self.backend.torch
    static_in = []
    for t in tensors:
      buf = torch.empty_strided(tuple(t.t.shape), tuple(t.t.stride()), dtype=t.t.dtype, device=t.t.device)
      buf.copy_(t.t)
      static_in.append(B200Tensor(buf, t.code))
    call_args = [a if k == "s" else _unflatten((k, a), static_in) for k, a in skel]
    graph = torch.cuda.CUDAGraph()
    cur = torch.cuda.current_stream()
    side = self.backend._capture_stream()  # pylint: disable=protected-access
    torch.cuda.synchronize()
    ok, out = True, None
    side.wait_stream(cur)
    # explicit begin / end (not the torch.cuda.graph context manager): the capture is ALWAYS ended and the current stream
    # ALWAYS restored, also when fun raises.  Operations that synchronise with the host raise CaptureUnsupported from the
    # adapter before any CUDA call is made (backend._no_capture), so a refused capture leaves no CUDA error behind.
    with torch.cuda.stream(side):
      be._capturing += 1  # pylint: disable=protected-access
      try:
        graph.capture_begin(capture_error_mode="thread_local")
        try:
          out = self.fun(*call_args)
        except Exception:  # pylint: disable=broad-except
          ok = False
        finally:
          try:
            graph.capture_end()
          except Exception:  # pylint: disable=broad-except
            ok = False
      except Exception:  # pylint: disable=broad-except
        ok = False
      finally:
        be._capturing -= 1  # pylint: disable=protected-access
    cur.wait_stream(side)
    if not ok:
      ent.eager = True
      self.stats["capture_failures"] += 1
      try:
        torch.cuda.synchronize()
      except Exception:  # pylint: disable=broad-except
        pass
      return False
    outs = []
    ent.out_skel = _flatten(out, outs)
    ent.static_in, ent.static_out, ent.graph = static_in, outs, graph
    self.stats["captures"] += 1
    return True
