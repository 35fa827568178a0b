mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_drivers.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --networks 74 --steps 10 --no-cpu-baseline > gpurun_out/bench_chain_pf.json 2> gpurun_out/bench_chain_pf.err; tail -3 gpurun_out/bench_chain_pf.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_chain_pf.json"))
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "clocks", d.get("clocks"), d.get("result_check"))
r=d["roofline"]; print("  roof", {k:r.get(k) for k in ("bound","achieved","peak","frac","kernel","kernel_share_of_step_time")}); [print("   ",k,v) for k,v in r["families"].items() if "chain" in k or "2cta" in k]
PY
