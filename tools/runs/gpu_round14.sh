mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err ) 2>&1 | tail -4; tail -2 gpurun_out/BENCH_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_reference.json 2> gpurun_out/BENCH_reference.err ) 2>&1 | tail -4
python - <<PY
import json
for f in ("gpurun_out/BENCH_default.json","gpurun_out/BENCH_reference.json"):
  txt=open(f).read().strip().splitlines(); print(f, "stdout lines:", len(txt))
  d=json.loads(txt[-1])
  print({k:d.get(k) for k in ("impl","value","unit","n_gpus","ms_per_step","steps","warmup","dtype","gpu_launches")})
  print("   e2e", d.get("e2e")); print("   cpu", d.get("cpu_baseline")); print("   clocks", d.get("clocks"))
  r=d.get("roofline") or {}; print("   roof", {k:r.get(k) for k in ("bound","achieved","peak","unit","frac","traffic","kernel","kernel_share_of_step_time")})
PY
