mkdir -p gpurun_out
for dt in bf16 f32 f64; do timeout 300 python bench.py --config flagship --dtype $dt --steps 20 > gpurun_out/cfg_flagship_${dt}_v3.json 2>> gpurun_out/cfgv3.err; done
timeout 300 python bench.py --config tree32 --steps 20 > gpurun_out/cfg_tree32_bf16_v3.json 2>> gpurun_out/cfgv3.err
timeout 600 python bench.py --config cfg3 --dtype f64 --steps 2 > gpurun_out/cfg_cfg3_f64_v3.json 2>> gpurun_out/cfgv3.err
timeout 900 python bench.py --config cfg5 --dtype f64 --steps 6 > gpurun_out/cfg_cfg5_f64_v3.json 2>> gpurun_out/cfgv3.err
tail -3 gpurun_out/cfgv3.err
for f in gpurun_out/cfg_*_v3.json; do python - <<PY
import json
d=json.loads(open("$f").read().strip().splitlines()[-1]); c=d.get("cpu_baseline") or {}
print("$f".split("/")[-1], {k:d.get(k) for k in ("value","ms_per_step")}, "cpu", {k:c.get(k) for k in ("value","cores","seconds","gflops")})
PY
done
