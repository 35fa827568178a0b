mkdir -p gpurun_out
for cfg in "9 3" "9 2" "9 1" "4 3" "18 2" "37 1" "74 1" "9 6" "6 4"; do
  set -- $cfg
  TNB200_CHAIN_G=$1 TNB200_CHAIN_ROT=$2 timeout 300 python bench.py --networks 74 --steps 5 --no-cpu-baseline > gpurun_out/sweep.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep.json"))
f=d["roofline"]["families"]
print("G=$1 ROT=$2", "ms_per_step", round(d["ms_per_step"],3), "chain_us", f.get("tcgen05_chain_16",{}).get("us"), "tf", f.get("tcgen05_chain_16",{}).get("tflops"), d["clocks"])
PY
done
