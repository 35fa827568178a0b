mkdir -p gpurun_out
for dt in f32 f64; do
timeout 600 python bench.py --dtype $dt --networks 74 --steps 5 --no-cpu-baseline > gpurun_out/bench_${dt}_n74.json 2> gpurun_out/bench_${dt}.err; tail -2 gpurun_out/bench_${dt}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${dt}_n74.json"))
print("$dt", {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d.get("result_check"))
r=d["roofline"]; print("  roof", {k:r.get(k) for k in ("bound","achieved","peak","frac","kernel","kernel_share_of_step_time")}); [print("   ",k,v) for k,v in r["families"].items()]
PY
done
timeout 300 python bench.py --config cfg4 --steps 20 > gpurun_out/cfg_cfg4_v2.json 2> gpurun_out/cfg4.err || tail -3 gpurun_out/cfg4.err
timeout 400 python bench.py --config cfg3 --dtype f64 --steps 2 > gpurun_out/cfg_cfg3_f64_v2.json 2> gpurun_out/cfg3.err || tail -3 gpurun_out/cfg3.err
timeout 600 python bench.py --config cfg5 --dtype f64 --steps 6 > gpurun_out/cfg_cfg5_f64_v2.json 2> gpurun_out/cfg5.err || tail -3 gpurun_out/cfg5.err
for f in gpurun_out/cfg_cfg4_v2.json gpurun_out/cfg_cfg3_f64_v2.json gpurun_out/cfg_cfg5_f64_v2.json; do python - <<PY
import json
d=json.load(open("$f")); print("$f", {k:d.get(k) for k in ("metric","value","ms_per_step")}, "cpu", (d.get("cpu_baseline") or {}).get("value"), {k:d.get(k) for k in ("parity","sizes","site_update_seconds")})
PY
done
