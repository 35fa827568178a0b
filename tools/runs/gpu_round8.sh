mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err ) 2>&1 | tail -4; tail -3 gpurun_out/BENCH_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/BENCH_default.json"))
print({k:d.get(k) for k in ("impl","value","unit","n_gpus","ms_per_step","steps","warmup","dtype","gpu_launches")})
print("   e2e", d.get("e2e")); print("   cpu", d.get("cpu_baseline")); print("   clocks", d.get("clocks"), d.get("result_check"))
r=d.get("roofline") or {}; print("   roof", {k:r.get(k) for k in ("bound","achieved","peak","unit","frac","traffic","kernel","kernel_share_of_step_time")})
PY
