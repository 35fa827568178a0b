"""Stress of the chained launch's dependency protocol: many replays under several tile schedules (short producer->
consumer distances included) must reproduce the step-by-step plan bit for bit every time."""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensornetwork_b200 as tb
from tensornetwork_b200 import drivers

be = tb.get_backend()
L, D, NB = 24, 256, 9
dims = [1] + [min(D, 2**min(i, L - i)) for i in range(1, L)] + [1]
labels = []
for side in "kb":
  for i in range(L):
    labels.append(["e0" if i == 0 else "%s%d" % (side, i), "p%d" % i, "eL" if i == L - 1 else "%s%d" % (side, i + 1)])
core = [(dims[i], 2, dims[i + 1]) for i in range(L)] * 2
shapes = [(NB,) + s for s in core]
sizes = {l: s[ax] for s, labs in zip(core, labels) for ax, l in enumerate(labs)}
path = drivers.greedy_path(labels, [], sizes)
rng = np.random.default_rng(5)
dev = [be.astype(be.convert_to_tensor((rng.standard_normal((NB,) + core[i]) / np.sqrt(core[i][0] * 2)).astype(np.float32)), "bfloat16")
       for i in range(L)]
al = {L + i: i for i in range(L)}
os.environ["TNB200_CHAIN_FORCE"] = "1"
ref_net = drivers.CompiledNetwork(be, shapes, "bfloat16", labels, [], path=path, nbatch=1, conj_aliases=al, use_chains=False)
ref_net.load(dev + [None] * L)
ref = ref_net().to_host().copy()
bad = 0
for G, rot, ring in itertools.product((1, 2, 3, 9), (1, 2, 3), (0, 2)):
  os.environ["TNB200_CHAIN_G"], os.environ["TNB200_CHAIN_ROT"], os.environ["TNB200_CHAIN_RING"] = str(G), str(rot), str(ring)
  net = drivers.CompiledNetwork(be, shapes, "bfloat16", labels, [], path=path, nbatch=1, conj_aliases=al, use_chains=True)
  assert net.chains, "no chain"
  net.load(dev + [None] * L)
  n_bad = 0
  for rep in range(60):
    out = net().to_host()
    if not np.array_equal(out, ref):
      n_bad += 1
  print("G=%d rot=%d ring=%d chain_steps=%d mismatching replays: %d / 60" % (G, rot, ring, max(len(c.steps) for c in net.chains), n_bad), flush=True)
  bad += n_bad
  del net
print("TOTAL mismatches", bad)
sys.exit(1 if bad else 0)
