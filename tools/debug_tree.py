"""Per-step error audit of the tree32 network: every pairwise contraction on the GPU is compared with a float64
numpy contraction of the SAME device operands (copied to host)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import tensornetwork_b200 as tb
from tensornetwork_b200 import drivers
from multigpu_check import tree_network

be = tb.get_backend()
dtype = sys.argv[1] if len(sys.argv) > 1 else "float32"
tensors, labels, sizes = tree_network(chi=128)
dev = [be.astype(be.convert_to_tensor(t.astype(np.float32)), dtype) for t in tensors]
path = drivers.greedy_path(labels, [], sizes)
steps, res = drivers.plan_path([t.shape for t in dev], labels, path, [], 0)
vals = list(dev)
for i, st in enumerate(steps):
  if st[0] == "tensordot":
    a, b = vals[st[1]], vals[st[2]]
    out = be.tensordot(a, b, (st[3], st[4]))
    kern = be.lib.tnb200_last_kernel().decode()
    ref = np.tensordot(a.to_host().astype(np.float64), b.to_host().astype(np.float64), (list(st[3]), list(st[4])))
    o = out.to_host().astype(np.float64)
    err = np.linalg.norm((o - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300)
    print(i, kern, a.shape, b.shape, st[3], st[4], "strides", a.t.stride(), b.t.stride(), "err %.3e" % err, flush=True)
    vals.append(out)
  else:
    vals.append(be.transpose(vals[st[1]], st[2]))
