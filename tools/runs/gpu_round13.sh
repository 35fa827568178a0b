mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/BENCH_n$N.json 2> gpurun_out/BENCH_n$N.err ) 2>&1 | tail -4; tail -3 gpurun_out/BENCH_n$N.err | cut -c1-300
python - <<PY
import json
txt=open("gpurun_out/BENCH_n$N.json").read().strip().splitlines()
print("stdout lines:", len(txt))
d=json.loads(txt[-1])
print({k:d.get(k) for k in ("impl","value","unit","n_gpus","ms_per_step","steps","warmup","dtype","gpu_launches")})
print("   e2e", d.get("e2e")); print("   clocks", d.get("clocks"), d.get("result_check"))
PY
free -g | head -2
