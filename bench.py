#!/usr/bin/env python
"""bench.py — headline benchmark of the cuda_b200 hot path (contract: see DESIGN.md §Measurement).

A *step* is one full greedy contraction of the closed <psi|psi> network of an L=64,
bond-dim-512, phys-dim-2 MPS (BASELINE.json configs[1]; 128 tensors -> 127 pairwise
contractions, SURVEY.md 8(d) cfg 2, seed 3, tensors scaled by 1/sqrt(contracted dims)).
`value` = pairwise contractions per second with inputs resident in HBM; `e2e` = the same
metric through the public API with HOST (pinned) input buffers, H2D + D2H inside the timed
region.  `--impl reference` times the reference's own CPU algorithm (numpy backend restated
in oracle/, all host threads) on the same workload.  N > 1 (torchrun): every rank contracts
its own independent MPS sample (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_SITES, BOND, PHYS = 64, 512, 2


# ----------------------------------------------------------------------------- workload
def mps_dims(L, D, d):
  return [1] + [min(D, d**min(i, L - i)) for i in range(1, L)] + [1]


def make_kets(L, D, d, seed):
  rng = np.random.default_rng(seed)
  dims = mps_dims(L, D, d)
  return [rng.standard_normal((dims[i], d, dims[i + 1])) / np.sqrt(dims[i] * d) for i in range(L)]


def norm_labels(L):
  labels = []
  for side in "kb":
    for i in range(L):
      labels.append(["e0" if i == 0 else "%s%d" % (side, i), "p%d" % i,
                     "eL" if i == L - 1 else "%s%d" % (side, i + 1)])
  return labels


def path_and_work(shapes, labels):
  from tensornetwork_b200 import drivers
  sizes = {l: s[ax] for s, labs in zip(shapes, labels) for ax, l in enumerate(labs)}
  path = drivers.greedy_path(labels, [], sizes)
  # algorithmic work per pairwise step: flops = 2MNK, bytes = (MK + KN + MN) * sizeof
  labs = [list(l) for l in labels]
  steps = []
  for a, b in path:
    l1, l2 = labs[a], labs[b]
    sh = [l for l in l1 if l in l2]
    K = int(np.prod([sizes[l] for l in sh])) if sh else 1
    M = int(np.prod([sizes[l] for l in l1 if l not in sh] or [1]))
    N = int(np.prod([sizes[l] for l in l2 if l not in sh] or [1]))
    steps.append((M, K, N))
    new = [l for l in l1 if l not in sh] + [l for l in l2 if l not in sh]
    for i in sorted([a, b], reverse=True):
      del labs[i]
    labs.append(new)
  return path, steps


# ------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
  QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
           "clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index=0):
    super().__init__(daemon=True)
    self.gpu = gpu_index
    self.samples = []
    self.stop_flag = False

  def run(self):
    while not self.stop_flag:
      try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                              "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 8:
          self.samples.append(f)
      except Exception:  # pylint: disable=broad-except
        pass
      time.sleep(0.05)

  def summary(self):
    if not self.samples:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(float(s[1]) for s in self.samples)
    reasons = []
    for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                      ("sw_power_cap", 7)):
      if any(s[col].lower().startswith("active") for s in self.samples):
        reasons.append(name)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][2]), "reasons": reasons,
            "samples": len(self.samples)}


# ------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
  """The reference's CPU algorithm for the same workload: numpy backend restated in oracle/
  (tensordot = numpy_backend.py:35-54 -> np.tensordot/OpenBLAS, path = greedy, pairwise loop
  = path_contractors.py:87-90), with all host threads."""
  if rank != 0:
    return
  from oracle import np_network as nn
  np_dtype = {"bf16": np.float32, "f32": np.float32, "f64": np.float64}[args.dtype]
  # bounded sample: each reference step contracts `nsamp` of the step's networks (same shapes, same path)
  nsamp = min(max(1, args.networks), 2)
  nets = []
  for b in range(nsamp):
    kets = [k.astype(np_dtype) for k in make_kets(L_SITES, BOND, PHYS, 3 + b)]
    nets.append(kets + [np.conj(k).copy() for k in kets])
  tensors = nets[0]
  labels = norm_labels(L_SITES)
  sizes = {l: t.shape[ax] for t, labs in zip(tensors, labels) for ax, l in enumerate(labs)}

  def step():
    out = None
    for ts in nets:
      path = nn.greedy_path(labels, [], sizes)   # the reference searches the path on every call
      out = nn.contract_path(ts, labels, path, [])
    return out
  for _ in range(args.warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    res = step()
  dt = time.perf_counter() - t0
  npair = len(tensors) - 1
  val = nsamp * npair * args.steps / dt
  cores = os.cpu_count()
  line = {
      "impl": "reference", "metric": "pairwise contractions/s", "value": val, "unit": "contractions/s",
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f32" if np_dtype == np.float32 else "f64", "data": "synthetic",
      "config": workload_config(args, 1),
      "cpu_baseline": {"value": val, "unit": "contractions/s", "cores": cores, "kind": "port",
                       "sample": "%d networks per step x %d steps (127 pairwise each), numpy %s, OpenBLAS threads=all"
                                 % (nsamp, args.steps, np.dtype(np_dtype).name)},
      "e2e": {"value": val, "unit": "contractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "result_check": float(np.real(res)),
  }
  print(json.dumps(line))


def workload_config(args, world):
  return {"workload": "cfg2: <psi|psi> of MPS L=%d bond-dim %d phys-dim %d, greedy path, 127 pairwise contractions per network"
                      % (L_SITES, BOND, PHYS),
          "networks_per_step_per_gpu": max(1, args.networks), "launch_mode": "eager" if args.no_graph else "cuda-graph replay", "compute_dtype": args.dtype, "path_provider": "numpy greedy (opt_einsum stand-in)",
          "parallelism": "replicas x%d (independent MPS samples, no collective)" % world,
          "l2_policy": "inputs (128 tensors) exceed the 126 MB L2 for f32/f64; bf16 inputs are 134 MB"}


# ------------------------------------------------------------------------------ our arm
def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="cuda_b200", choices=["cuda_b200", "reference"])
  ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32", "f64"])
  ap.add_argument("--networks", type=int, default=8,
                  help="independent MPS samples contracted in lock-step per step per GPU (batched kernels)")
  ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
  ap.add_argument("--cpu-baseline-steps", type=int, default=2)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "cuda_b200" else max(args.warmup, 1)
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))

  if args.impl == "reference":
    run_reference(args, rank, world)
    return

  import torch
  import torch.distributed as dist
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  torch.cuda.set_device(local)
  import tensornetwork_b200 as tb
  from tensornetwork_b200 import drivers, _lib
  be = tb.get_backend()
  lib = be.lib

  code = {"bf16": _lib.BF16, "f32": _lib.F32, "f64": _lib.F64}[args.dtype]
  tdtype = {"bf16": torch.bfloat16, "f32": torch.float32, "f64": torch.float64}[args.dtype]
  esize = {"bf16": 2, "f32": 4, "f64": 8}[args.dtype]
  NB = max(1, args.networks)
  nbatch = 1 if NB > 1 else 0
  # NB independent MPS samples per rank, generated ON THE DEVICE by the library's Philox kernel (seeds differ
  # per rank / site; numpy would need ~30 s to draw 2.4e9 normals for NB = 74), scaled by 1/sqrt(contracted dims)
  dims = mps_dims(L_SITES, BOND, PHYS)
  labels = norm_labels(L_SITES)
  core_shapes = [(dims[i], PHYS, dims[i + 1]) for i in range(L_SITES)] * 2
  shapes = [((NB,) + cs) if nbatch else cs for cs in core_shapes]
  path, work = path_and_work(core_shapes, labels)
  npair = len(path)
  flops_step = NB * sum(2.0 * m * k * n for m, k, n in work)
  bytes_step = NB * sum((m * k + k * n + m * n) * esize for m, k, n in work)
  kets = []
  for i in range(L_SITES):
    t = be.randn(shapes[i], np.float32, seed=1 + 7919 * rank + i)
    t *= 1.0 / np.sqrt(dims[i] * PHYS)
    kets.append(be.astype(t, {"bf16": "bfloat16", "f32": np.float32, "f64": np.float64}[args.dtype]))
  dev = kets + [be.copy(k) for k in kets]                       # bra = conj(ket) (real data): its own 64 tensors
  h2d_bytes = sum(int(np.prod(s)) * esize for s in shapes)

  net = drivers.CompiledNetwork(be, shapes, {"bf16": "bfloat16", "f32": np.float32, "f64": np.float64}[args.dtype],
                                labels, [], path=path, nbatch=nbatch) if not args.no_graph else None
  host = None
  if net is not None:
    net.load(dev)
    # the public API's pinned staging arena holds the step's 128 host tensors (filled once, outside the timed
    # region, with the same synthetic data; a user would generate / load their data straight into these views)
    host = net.host_staging()
    for dst, src in zip(host, dev):
      dst.copy_(src.t)
  else:
    host = [d.t.cpu().pin_memory() for d in dev]
  torch.cuda.synchronize()

  def step_resident():
    if net is not None:
      return net()
    return drivers.contract_network(dev, labels, [], path=path, backend=be, nbatch=nbatch)

  def step_e2e():
    if net is not None:
      out = net.run_staged()    # ONE H2D of the pinned staging arena (all inputs), then graph replay
    else:
      ts = [tb.B200Tensor(h.to(be.device, non_blocking=True), code) for h in host]
      out = drivers.contract_network(ts, labels, [], path=path, backend=be, nbatch=nbatch)
    return out.t.to("cpu")      # D2H of the result (one scalar per network; syncs)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident timing ------------------------------------------------------
  for _ in range(args.warmup):
    res = step_resident()
  barrier()
  sampler = ClockSampler(local)
  sampler.start()
  l0 = lib.tnb200_launch_count()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    res = step_resident()
  e1.record()
  barrier()
  launches = (net.launches_per_replay * args.steps) if net is not None else (lib.tnb200_launch_count() - l0)
  ms = e0.elapsed_time(e1)
  sampler.stop_flag = True
  sampler.join(timeout=2)
  result_value = [float(x) for x in np.atleast_1d(res.to_host().astype(np.float64))]

  # ---- per-kernel device times of one step (live, CUDA events) -> dominant kernel and its roofline
  kstats = kernel_profile(be, dev, labels, path, work, nbatch, NB, esize)

  # ---- end-to-end timing (host buffers) ----------------------------------------------
  for _ in range(args.warmup):
    step_e2e()
  barrier()
  t0 = torch.cuda.Event(enable_timing=True)
  t1 = torch.cuda.Event(enable_timing=True)
  t0.record()
  for _ in range(args.steps):
    out = step_e2e()
  t1.record()
  barrier()
  ms_e2e = t0.elapsed_time(t1)

  if world > 1:
    tt = torch.tensor([ms, ms_e2e], device=be.device, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(tt[0]), float(tt[1])

  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # pylint: disable=broad-except
      pass
    if args.dtype == "bf16":
      peak, peak_src = peaks.get("bf16_tflops", 1590.0), ("measured" if peaks else "fallback")
      peak_note = "bf16 dense (cuBLAS burst), MEASURED_PEAKS.json" if peaks else "fallback 1.59 PF"
    elif args.dtype == "f32":
      peak = peaks.get("bf16_tflops", 1590.0) / 2.0
      peak_src, peak_note = "derived", "tf32 = measured bf16 / 2 (no measured tf32 figure)"
    else:
      peak, peak_src, peak_note = 40.0, "nominal", "B200 FP64 nominal 40 TFLOP/s (no measured fp64 figure)"
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    if args.dtype == "bf16" and "bf16_tflops_sustained" in peaks:
      # the dominant kernel is timed inside a long step: the sustained figure is its tensor roof
      peak, peak_note = peaks["bf16_tflops_sustained"], "bf16 dense sustained (cuBLAS back to back), MEASURED_PEAKS.json"
    # dominant kernel = the family with the largest share of the step's device time
    ktot = sum(d["us"] for d in kstats.values())
    kern_name = max(kstats, key=lambda k: kstats[k]["us"])
    kd = kstats[kern_name]
    k_tf = kd["flops"] / (kd["us"] * 1e-6) / 1e12
    k_gbs = kd["bytes"] / (kd["us"] * 1e-6) / 1e9
    t_tensor, t_hbm = kd["flops"] / (peak * 1e12), kd["bytes"] / (hbm_peak * 1e9)
    if t_hbm >= t_tensor:
      roof = {"bound": "hbm", "achieved": k_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k_gbs / hbm_peak}
    else:
      roof = {"bound": "tensor", "achieved": k_tf, "peak": peak, "unit": "TFLOP/s", "frac": k_tf / peak}
    traffic = None
    try:
      tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
      if tr.get("kernel") == kern_name and tr.get("dtype") == args.dtype and tr.get("networks") == NB:
        traffic = tr.get("dram_bytes_per_launch")
    except Exception:  # pylint: disable=broad-except
      pass
    roof.update({
        "traffic": traffic, "kernel": kern_name, "kernel_launches_per_step": kd["launches"],
        "kernel_us_per_launch": kd["us"] / kd["launches"], "kernel_share_of_step_time": kd["us"] / ktot,
        "algorithmic_mb_per_launch": kd["bytes"] / kd["launches"] / 1e6,
        "algorithmic_gflop_per_launch": kd["flops"] / kd["launches"] / 1e9,
        "kernel_tflops": k_tf, "kernel_gbs": k_gbs, "tensor_peak_tflops": peak, "hbm_peak_gbs": hbm_peak,
        "peak_source": peak_src, "peak_note": peak_note,
        "arithmetic_intensity_flop_per_byte": kd["flops"] / kd["bytes"], "ridge_flop_per_byte": peak * 1e3 / hbm_peak,
        "note": "dominant kernel timed live with CUDA events around each of its launches (eager replay of the "
                "step); algorithmic bytes = operands + result once, algorithmic flops = 2MNK; the binding roof "
                "is the one with the larger minimum time",
        "families": {k: {"launches": round(v["launches"], 2), "us": round(v["us"], 1),
                         "tflops": round(v["flops"] / (v["us"] * 1e-6) / 1e12, 1),
                         "gbs": round(v["bytes"] / (v["us"] * 1e-6) / 1e9, 1)} for k, v in kstats.items()},
    })
    achieved = flops_step * args.steps / (ms * 1e-3) / 1e12
    value = world * NB * npair * args.steps / (ms * 1e-3)
    line = {
        "metric": "pairwise contractions/s", "value": value, "unit": "contractions/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": workload_config(args, world),
        "step_tflops": world * achieved,
        "step_hbm_gbs": world * bytes_step * args.steps / (ms * 1e-3) / 1e9,
        "launches_per_step": launches / args.steps,
        "algorithmic_gflop_per_step": flops_step / 1e9, "algorithmic_mb_per_step": bytes_step / 1e6,
        "roofline": roof,
        "e2e": {"value": world * NB * npair * args.steps / (ms_e2e * 1e-3), "unit": "contractions/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": esize * NB,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "result_check": result_value[:4],
    }
    if not args.no_cpu_baseline and world == 1:
      line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def kernel_profile(be, dev, labels, path, work, nbatch, nb, esize, reps=3):
  """Per-kernel device time of one step, measured LIVE with CUDA events on the launching stream: the plan is
  replayed eagerly `reps` times with an event pair around every pairwise contraction; the library reports which
  kernel family served it.  Returns {family: {"launches", "us", "flops", "bytes"}} averaged per step
  (bytes = algorithmic operand + result bytes of the contraction, each counted once)."""
  import torch
  from tensornetwork_b200 import drivers
  steps, _ = drivers.plan_path([t.shape for t in dev], labels, path, [], nbatch)
  stats = {}
  for rep in range(reps + 1):
    vals = list(dev)
    evs = []
    wi = 0
    for st in steps:
      if st[0] in ("tensordot", "batched"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if st[0] == "tensordot":
          vals.append(be.tensordot(vals[st[1]], vals[st[2]], (st[3], st[4])))
        else:
          vals.append(be._contract(vals[st[1]], vals[st[2]], list(st[3]), list(st[4]), list(st[5]), list(st[6])))
        e1.record()
        evs.append((be.lib.tnb200_last_kernel().decode(), work[wi], e0, e1))
        wi += 1
      else:
        vals.append(be.transpose(vals[st[1]], st[2]))
    torch.cuda.synchronize()
    del vals
    if rep == 0:
      continue                                  # warm-up pass (allocator, descriptors)
    for name, (m, k, n), e0, e1 in evs:
      d = stats.setdefault(name, {"launches": 0, "us": 0.0, "flops": 0.0, "bytes": 0.0})
      d["launches"] += 1.0 / reps
      d["us"] += e0.elapsed_time(e1) * 1e3 / reps
      d["flops"] += nb * 2.0 * m * k * n / reps
      d["bytes"] += nb * float(m * k + k * n + m * n) * esize / reps
  return stats


def cpu_baseline(args):
  """oracle (numpy restatement of the reference numpy backend) on the host cores, bounded sample."""
  from oracle import np_network as nn
  np_dtype = np.float64 if args.dtype == "f64" else np.float32
  kets = [k.astype(np_dtype) for k in make_kets(L_SITES, BOND, PHYS, 3)]
  tensors = kets + [np.conj(k).copy() for k in kets]
  labels = norm_labels(L_SITES)
  sizes = {l: t.shape[ax] for t, labs in zip(tensors, labels) for ax, l in enumerate(labs)}
  path = nn.greedy_path(labels, [], sizes)
  nn.contract_path(tensors, labels, path, [])
  t0 = time.perf_counter()
  n = args.cpu_baseline_steps
  for _ in range(n):
    nn.contract_path(tensors, labels, path, [])
  dt = time.perf_counter() - t0
  return {"value": (len(tensors) - 1) * n / dt, "unit": "contractions/s", "cores": os.cpu_count(), "kind": "port",
          "sample": "%d full networks (127 pairwise each) in numpy %s, all host threads" % (n, np.dtype(np_dtype).name)}


if __name__ == "__main__":
  main()
