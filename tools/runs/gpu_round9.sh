mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/BENCH_n2.json 2> gpurun_out/BENCH_n2.err ) 2>&1 | tail -4; tail -3 gpurun_out/BENCH_n2.err | cut -c1-300
( time timeout 600 $TR --master-port 29522 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/BENCH_ref_n2.json 2>/dev/null ) 2>&1 | tail -4
python - <<PY
import json
for f in ("gpurun_out/BENCH_n2.json","gpurun_out/BENCH_ref_n2.json"):
  txt=open(f).read().strip().splitlines()
  print(f, "stdout lines:", len(txt))
  d=json.loads(txt[-1])
  print({k:d.get(k) for k in ("impl","value","unit","n_gpus","ms_per_step","steps","warmup","dtype","gpu_launches")})
  print("   e2e", d.get("e2e")); print("   clocks", d.get("clocks"), d.get("result_check"))
PY
