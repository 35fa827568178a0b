# full verification on one GPU: every gpu-marked test, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) 2>&1 | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err ) 2>&1 | tail -4; tail -2 gpurun_out/BENCH_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_reference.json 2> gpurun_out/BENCH_reference.err ) 2>&1 | tail -4
python - <<PY
import json
for f in ("gpurun_out/BENCH_default.json","gpurun_out/BENCH_reference.json"):
  d=json.load(open(f))
  print(f, {k:d.get(k) for k in ("impl","value","unit","n_gpus","ms_per_step","steps","warmup","dtype","gpu_launches")})
  print("   e2e", d.get("e2e")); print("   cpu", d.get("cpu_baseline")); print("   clocks", d.get("clocks"), "single", d.get("single_network"))
  r=d.get("roofline") or {}; print("   roof", {k:r.get(k) for k in ("bound","achieved","peak","unit","frac","traffic","kernel","kernel_share_of_step_time")})
PY
( time timeout 900 python bench.py --dtype f64 --steps 3 --no-cpu-baseline > gpurun_out/BENCH_f64.json 2> gpurun_out/BENCH_f64.err ) 2>&1 | tail -3; tail -2 gpurun_out/BENCH_f64.err
python - <<PY
import json
d=json.loads(open("gpurun_out/BENCH_f64.json").read().strip().splitlines()[-1])
print("f64", {k:d.get(k) for k in ("value","ms_per_step")}, d["e2e"])
PY
