"""torchrun -n N tools/multigpu_check.py — NCCL check of the subtree-sharded contraction of the seed-7
32-node tree tensor network against the single-GPU result, plus timing at 1 and N GPUs."""
import json
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def tree_network(n_nodes=16, chi=128, d=2, seed=7, maxdeg=3):
  rng = np.random.default_rng(seed)
  parent, deg = [-1], [0]
  for i in range(1, n_nodes):
    while True:
      p = int(rng.integers(0, i))
      if deg[p] < maxdeg:
        break
    parent.append(p); deg[p] += 1; deg.append(1)
  lk = [["p%d" % i] for i in range(n_nodes)]
  lb = [["p%d" % i] for i in range(n_nodes)]
  sizes = {"p%d" % i: d for i in range(n_nodes)}
  for i in range(1, n_nodes):
    for tag, L in (("k", lk), ("b", lb)):
      e = "%s%d_%d" % (tag, parent[i], i)
      L[i].append(e); L[parent[i]].append(e); sizes[e] = chi
  kets = [rng.standard_normal([sizes[l] for l in labs]) / np.sqrt(np.prod([sizes[l] for l in labs[1:]]) or 1.0) for labs in lk]
  return kets + [np.conj(k) for k in kets], lk + lb, sizes


def main():
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  rank, world = dist.get_rank(), dist.get_world_size()
  import tensornetwork_b200 as tb
  from tensornetwork_b200 import drivers, parallel
  be = tb.get_backend()
  chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
  dtype = sys.argv[2] if len(sys.argv) > 2 else "float32"
  tensors, labels, sizes = tree_network(chi=chi)
  dev = [be.astype(be.convert_to_tensor(t.astype(np.float32)), dtype) for t in tensors]
  path = drivers.greedy_path(labels, [], sizes)
  # single-GPU reference on every rank
  for _ in range(2):
    ref = drivers.contract_network(dev, labels, [], path=path, backend=be)
  torch.cuda.synchronize(); dist.barrier()
  t0 = time.perf_counter()
  for _ in range(5):
    ref = drivers.contract_network(dev, labels, [], path=path, backend=be)
  torch.cuda.synchronize(); t1 = (time.perf_counter() - t0) / 5
  for _ in range(2):
    res, root, info = parallel.contract_network_parallel(be, dev, labels, [], path=path)
  torch.cuda.synchronize(); dist.barrier()
  t0 = time.perf_counter()
  for _ in range(5):
    res, root, info = parallel.contract_network_parallel(be, dev, labels, [], path=path)
  torch.cuda.synchronize(); dist.barrier(); tn_ = (time.perf_counter() - t0) / 5
  if rank == root:
    a, b = float(res.to_host().astype(np.float64)), float(ref.to_host().astype(np.float64))
    ok = abs(a - b) <= 2e-3 * abs(b)
    print(json.dumps({"world": world, "chi": chi, "dtype": dtype, "single_gpu_s": t1, "sharded_s": tn_, "speedup": t1 / tn_,
                      "bound_total_over_critical": info["total"] / info["critical"], "per_rank_gflop": [x / 1e9 for x in info["per_rank"]],
                      "result": a, "single": b, "match": bool(ok)}))
    assert ok
  dist.barrier()
  dist.destroy_process_group()


if __name__ == "__main__":
  main()
