mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_tensordot.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --dtype f64 --networks 74 --steps 3 --no-cpu-baseline --no-e2e-overlap > gpurun_out/bench_f64_v12.json 2> gpurun_out/bench_f64.err; tail -2 gpurun_out/bench_f64.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_f64_v12.json").read().strip().splitlines()[-1])
print("f64", {k:d.get(k) for k in ("value","ms_per_step")}, d.get("result_check")); [print("   ",k,v) for k,v in d["roofline"]["families"].items()]
PY
