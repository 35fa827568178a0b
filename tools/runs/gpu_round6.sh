mkdir -p gpurun_out
for nomma in 0 1; do
TNB200_CHAIN_NOMMA=$nomma timeout 300 python bench.py --networks 74 --steps 10 --no-cpu-baseline > gpurun_out/bench_nomma$nomma.json 2> gpurun_out/bench_nomma.err; tail -3 gpurun_out/bench_nomma.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_nomma$nomma.json"))
print("nomma=$nomma", {k:d.get(k) for k in ("value","ms_per_step")}, d.get("result_check"))
r=d["roofline"]; [print("   ",k,v) for k,v in r["families"].items() if "chain" in k]
PY
done
