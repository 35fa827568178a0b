"""Host-side planning logic (no GPU): operand-order selection, output shapes and chain discovery of path plans."""
import numpy as np
from tensornetwork_b200 import drivers
from oracle import np_network as nn


def _norm_network(L, D, d=2):
  dims = [1] + [min(D, d**min(i, L - i)) for i in range(1, L)] + [1]
  labels = []
  for side in "kb":
    for i in range(L):
      labels.append(["e0" if i == 0 else "%s%d" % (side, i), "p%d" % i, "eL" if i == L - 1 else "%s%d" % (side, i + 1)])
  return dims, labels


def _replay(shapes, steps):
  """numpy replay of a plan on random data (reference semantics of each step kind)"""
  rng = np.random.default_rng(0)
  vals = [rng.standard_normal(s) for s in shapes]
  for st in steps:
    if st[0] == "tensordot":
      vals.append(np.tensordot(vals[st[1]], vals[st[2]], (list(st[3]), list(st[4]))))
    elif st[0] == "batched":
      a, b = vals[st[1]], vals[st[2]]
      nb = len(st[5])
      assert tuple(st[5]) == tuple(range(nb)) == tuple(st[6])
      out = np.stack([np.tensordot(a[i], b[i], ([x - 1 for x in st[3]], [x - 1 for x in st[4]])) for i in range(a.shape[0])]) \
          if nb == 1 else None
      vals.append(out)
    elif st[0] == "transpose":
      vals.append(np.transpose(vals[st[1]], st[2]))
  return vals


def test_plan_shapes_and_operand_order_on_cfg2():
  """cfg 2 (L=64, D=512): after operand-order selection every pairwise step is a plain row-major GEMM view
  (contracted axes trail the first operand and lead the second, or both lead / both trail), the plan's shapes match
  a numpy replay at a small bond dimension, and the result is unchanged."""
  L = 64
  dims, labels = _norm_network(L, 512)
  core = [(dims[i], 2, dims[i + 1]) for i in range(L)] * 2
  sizes = {l: s[ax] for s, labs in zip(core, labels) for ax, l in enumerate(labs)}
  path = nn.greedy_path(labels, [], sizes)
  steps, res = drivers.plan_path(core, labels, path, [])
  shp = drivers.plan_shapes(core, steps)
  assert len([s for s in steps if s[0] == "tensordot"]) == 127 and shp[res] == ()
  for st in steps:
    if st[0] != "tensordot":
      continue
    na, nb_ = len(shp[st[1]]), len(shp[st[2]])
    n = len(st[3])
    trail_a = tuple(st[3]) == tuple(range(na - n, na))
    lead_a = tuple(sorted(st[3])) == tuple(range(n))
    lead_b = tuple(sorted(st[4])) == tuple(range(n))
    trail_b = tuple(sorted(st[4])) == tuple(range(nb_ - n, nb_))
    assert (trail_a or lead_a) and (lead_b or trail_b), st
  # small-D replay: shapes and value
  dims, labels = _norm_network(12, 8)
  core = [(dims[i], 2, dims[i + 1]) for i in range(12)] * 2
  sizes = {l: s[ax] for s, labs in zip(core, labels) for ax, l in enumerate(labs)}
  path = nn.greedy_path(labels, [], sizes)
  steps, res = drivers.plan_path(core, labels, path, [])
  vals = _replay(core, steps)
  assert [v.shape for v in vals] == list(drivers.plan_shapes(core, steps))
  rng = np.random.default_rng(0)
  ts = [rng.standard_normal(s) for s in core]
  np.testing.assert_allclose(vals[res], nn.contract_path(ts, labels, path, []), rtol=1e-12)


def test_find_chains_on_cfg2():
  """the zipper of cfg 2 is one run of consecutive dependent steps (what CompiledNetwork offers to tnb200_chain_create)"""
  L = 64
  dims, labels = _norm_network(L, 512)
  core = [(dims[i], 2, dims[i + 1]) for i in range(L)] * 2
  sizes = {l: s[ax] for s, labs in zip(core, labels) for ax, l in enumerate(labs)}
  path = nn.greedy_path(labels, [], sizes)
  steps, _ = drivers.plan_path([(3,) + c for c in core], labels, path, [], nbatch=1)
  runs = drivers.find_chains(steps, len(core))
  longest = max(runs, key=len)
  assert len(longest) >= 88
  n_in = len(core)
  for a, b in zip(longest[:-1], longest[1:]):
    assert b == a + 1 and (n_in + a) in (steps[b][1], steps[b][2])
  assert all(len(r) >= 2 for r in runs)
  flat = [i for r in runs for i in r]
  assert len(flat) == len(set(flat))


def test_local_subtrees_cover_the_partition_exactly():
  """parallel.local_subtrees / ssa_to_linear (the graph-replayed shards of ShardedNetwork): on every rank the local
  subtrees plus the steps above the cut are exactly the steps the rank owns, and replaying each subtree's linear
  sub-path with numpy reproduces the intermediate the full path produces."""
  import sys, os
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
  from tensornetwork_b200 import parallel
  from oracle import np_network as nn
  rng = np.random.default_rng(5)
  n_nodes, chi, d = 10, 4, 2
  parent = [-1] + [int(rng.integers(0, i)) for i in range(1, n_nodes)]
  lk = [["p%d" % i] for i in range(n_nodes)]
  lb = [["p%d" % i] for i in range(n_nodes)]
  sizes = {"p%d" % i: d for i in range(n_nodes)}
  for i in range(1, n_nodes):
    for tag, L in (("k", lk), ("b", lb)):
      e = "%s%d_%d" % (tag, parent[i], i)
      L[i].append(e); L[parent[i]].append(e); sizes[e] = chi
  labels = lk + lb
  tensors = [rng.standard_normal([sizes[l] for l in labs]) for labs in labels]
  path = nn.greedy_path(labels, [], sizes)
  n = len(labels)
  ssa = parallel.path_to_ssa(n, path)
  flops = [2.0 * m * k * nn_ for m, k, nn_ in nn.network_flops(labels, path, sizes)]
  # full replay: value and labels of every intermediate
  vals = {i: (tensors[i], list(labels[i])) for i in range(n)}
  for a, b, o in ssa:
    vals[o] = nn.contract_between(vals[a][0], vals[a][1], vals[b][0], vals[b][1])
  for world in (2, 3, 4):
    owner, transfers, info = parallel.partition_tree(n, path, flops, world)
    seen = set()
    for rank in range(world):
      roots, pure = parallel.local_subtrees(n, ssa, owner, rank)
      assert all(owner[s] == rank for s in pure)
      covered = set()
      for root, (leaves, steps) in roots.items():
        covered |= set(steps)
        sub = parallel.ssa_to_linear(leaves, steps, ssa)
        got = nn.contract_path([tensors[i] for i in leaves], [labels[i] for i in leaves], sub, vals[root][1])
        np.testing.assert_allclose(got, vals[root][0], rtol=1e-12, atol=1e-12)
      assert covered == pure
      seen |= {s for s in range(len(ssa)) if owner[s] == rank}
    assert seen == set(range(len(ssa)))
