mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_tensordot.py -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --dtype f32 --networks 74 --steps 5 --no-cpu-baseline > gpurun_out/bench_f32_n74_v12.json 2> gpurun_out/bench_f32.err; tail -2 gpurun_out/bench_f32.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_f32_n74_v12.json"))
print("f32", {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d.get("result_check"))
r=d["roofline"]; [print("   ",k,v) for k,v in r["families"].items()]
PY
