mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_tcgen05.py -x -q -m gpu 2>&1 | tail -5
for nb in 74 37 148; do
timeout 300 python bench.py --networks $nb --steps 10 --no-cpu-baseline > gpurun_out/bench_chain_n$nb.json 2> gpurun_out/bench_chain_n$nb.err; tail -3 gpurun_out/bench_chain_n$nb.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_chain_n$nb.json"))
print($nb, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "single", d.get("single_network"), "clocks", d.get("clocks"), d.get("result_check"))
r=d["roofline"]; print("  roof", {k:r.get(k) for k in ("bound","achieved","peak","frac","kernel","kernel_share_of_step_time")}); [print("   ",k,v) for k,v in r["families"].items() if "chain" in k or "2cta" in k]
PY
done
