mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_drivers.py -x -q -m gpu 2>&1 | tail -5
( time timeout 600 python bench.py --steps 20 > gpurun_out/bench_default_v9.json 2> gpurun_out/bench_default_v9.err ) 2>&1 | tail -4
tail -2 gpurun_out/bench_default_v9.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default_v9.json"))
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","steps")}, "e2e", d.get("e2e"), "cpu", d.get("cpu_baseline"), "single", d.get("single_network"), "clocks", d.get("clocks"))
r=d["roofline"]; print("  roof", {k:r.get(k) for k in ("bound","achieved","peak","frac","kernel","kernel_share_of_step_time")}); [print(k,v) for k,v in r["families"].items()]
PY
for nb in 18 37; do timeout 300 python bench.py --networks $nb --steps 20 --no-cpu-baseline > gpurun_out/bench_n${nb}_v9.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${nb}_v9.json"))
print($nb, d["value"], d["ms_per_step"], d["e2e"]["value"]); r=d["roofline"]; print("  roof", {k:r.get(k) for k in ("bound","achieved","frac","kernel","kernel_us_per_launch")})
PY
done
