"""world_size-2 gloo worker (CPU): exercises tensornetwork_b200.parallel with the numpy oracle as
the per-pair contraction and torch.distributed (gloo) as the transport."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensornetwork_b200 import parallel
from oracle import np_network as nn


def main():
  dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"],
                          rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
  rank, world = dist.get_rank(), dist.get_world_size()
  rng = np.random.default_rng(77)   # same inputs on every rank
  # a small tree tensor network: 8 ket nodes + 8 conj bra nodes, closed -> scalar
  nn_nodes, chi, d = 8, 6, 2
  parent = [-1] + [int(rng.integers(0, i)) for i in range(1, nn_nodes)]
  labels_k = [["p%d" % i] for i in range(nn_nodes)]
  labels_b = [["p%d" % i] for i in range(nn_nodes)]
  sizes = {"p%d" % i: d for i in range(nn_nodes)}
  for i in range(1, nn_nodes):
    for tag, L in (("k", labels_k), ("b", labels_b)):
      e = "%s%d_%d" % (tag, parent[i], i)
      L[i].append(e); L[parent[i]].append(e); sizes[e] = chi
  kets = [rng.standard_normal([sizes[l] for l in labs]) / 3.0 for labs in labels_k]
  tensors = kets + [np.conj(k) for k in kets]
  labels = labels_k + labels_b
  path = nn.greedy_path(labels, [], sizes)
  flops = [2.0 * m * k * n for m, k, n in nn.network_flops(labels, path, sizes)]
  serial = nn.contract_path(tensors, labels, path, [])

  lab = {}

  def send(t, dst):
    dist.send(torch.from_numpy(np.ascontiguousarray(t)), dst)

  def recv(tid, src):
    # shape from the symbolic labels of the intermediate
    shape = [sizes[l] for l in lab[tid]]
    buf = torch.empty(shape, dtype=torch.float64)
    dist.recv(buf, src)
    return buf.numpy()
  # symbolic labels of all intermediates (same replay as inside contract_tree_parallel)
  ssa = parallel.path_to_ssa(len(tensors), path)
  for i, l in enumerate(labels):
    lab[i] = list(l)
  for a, b, o in ssa:
    sh = [l for l in lab[a] if l in lab[b]]
    lab[o] = [l for l in lab[a] if l not in sh] + [l for l in lab[b] if l not in sh]
  res, root_rank, info = parallel.contract_tree_parallel(tensors, labels, [], path, rank, world,
                                                         nn.contract_between, send, recv, flops)
  owner, transfers, _ = parallel.partition_tree(len(tensors), path, flops, world)
  assert set(owner) == set(range(world)), "both ranks must get work: %s" % owner
  assert len(transfers) >= 1
  if rank == root_rank:
    assert abs(float(res) - float(serial)) <= 1e-12 * abs(float(serial)), (res, serial)
  # independent units (MPS batch samples): shard, contract, all-gather
  units = list(range(7))
  got = parallel.contract_independent(units, lambda u: float(u) ** 2, rank, world,
                                      gather=lambda mine: _all_gather(mine, world))
  assert got == [float(u) ** 2 for u in units], got
  assert list(parallel.shard_range(7, 0, 2)) == [0, 1, 2, 3] and list(parallel.shard_range(7, 1, 2)) == [4, 5, 6]
  dist.barrier()
  if rank == 0:
    print("PARALLEL OK total=%.3g critical=%.3g per_rank=%s" % (info["total"], info["critical"], info["per_rank"]))
  dist.destroy_process_group()


def _all_gather(mine, world):
  out = [None] * world
  dist.all_gather_object(out, mine)
  return out


if __name__ == "__main__":
  main()
