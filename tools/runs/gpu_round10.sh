echo "== default"; timeout 200 python tools/kernel_bench.py --ramp 74 2>&1 | tail -14
for c in 4 16 32; do echo "== THIN_CPS=$c"; TNB200_THIN_CPS=$c timeout 200 python tools/kernel_bench.py --ramp 74 2>&1 | grep simt; done
for c in 2 4 8; do echo "== THIN_MMA_CPS=$c"; TNB200_THIN_MMA_CPS=$c timeout 200 python tools/kernel_bench.py --ramp 74 2>&1 | grep mma; done
