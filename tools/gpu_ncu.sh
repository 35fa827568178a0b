mkdir -p gpurun_out
# (1) launch list of the bench command (durations only; cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_bf16_n74_v10.csv python bench.py --networks 74 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
# (2) full-set capture of the chained zipper kernel (third launch: a warm graph replay)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_chain --launch-skip 2 --launch-count 1 -f -o gpurun_out/ncu_chain_bf16_n74_v10 python bench.py --networks 74 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_chain.log 2>&1
tail -3 gpurun_out/ncu_chain.log | cut -c1-300
ncu -i gpurun_out/ncu_chain_bf16_n74_v10.ncu-rep --page raw --csv > gpurun_out/ncu_chain_bf16_n74_v10_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
