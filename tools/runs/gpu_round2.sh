set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 600 python bench.py --steps 10 > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err ) 2>&1 | tail -4
tail -2 gpurun_out/bench_default_n1.err
( time timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2>&1 | tail -4
tail -3 gpurun_out/bench_n2.err
( time timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err ) 2>&1 | tail -4
timeout 600 $TR --master-port 29513 tools/multigpu_check.py 128 float32 2>&1 | tail -3
timeout 600 $TR --master-port 29514 tools/multigpu_check.py 256 bfloat16 2>&1 | tail -3
for f in gpurun_out/bench_default_n1.json gpurun_out/bench_n2.json gpurun_out/bench_ref_n2.json; do python - <<PY
import json
d=json.load(open("$f"))
print("$f", {k:d.get(k) for k in ("value","n_gpus","ms_per_step","steps","impl")}, "e2e", d.get("e2e"), "cpu", d.get("cpu_baseline"), "single", d.get("single_network"), "clocks", d.get("clocks"))
r=d.get("roofline") or {}
print("  roof", {k:r.get(k) for k in ("bound","achieved","peak","frac","kernel","kernel_share_of_step_time")})
PY
done
