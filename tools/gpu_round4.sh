mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_drivers.py -x -q -m gpu -k "chained or compiled" 2>&1 | tail -15
timeout 300 python bench.py --networks 74 --steps 10 --no-cpu-baseline > gpurun_out/bench_chain_n74.json 2> gpurun_out/bench_chain_n74.err; tail -5 gpurun_out/bench_chain_n74.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_chain_n74.json"))
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","steps","gpu_launches")}, "e2e", d.get("e2e"), "single", d.get("single_network"), "clocks", d.get("clocks"), d.get("result_check"))
r=d["roofline"]; print("  roof", {k:r.get(k) for k in ("bound","achieved","peak","frac","kernel","kernel_share_of_step_time")}); [print(k,v) for k,v in r["families"].items()]
PY
