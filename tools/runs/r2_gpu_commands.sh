#!/bin/bash
# Round 2: the gpurun command lines behind the files in profiles/r2_* (each line is one gpurun call; nothing here runs by itself).
# Usage on a GPU box: pick a line and run its quoted part from the repository root.
cat <<'EOF'
# full GPU test suite                      -> profiles/r2_gpu_tests_v1.log
python -m pytest tests -m gpu -x -q
# driver's default line with all sub-records -> profiles/r2_BENCH_default_n1_v{1,2}.json
python bench.py --steps 20 --warmup 5
# reference arm                            -> profiles/r2_BENCH_reference_n1_v1.json
python bench.py --impl reference --steps 3 --warmup 1
# 2 / 4 GPUs with the strong-scaling record -> profiles/r2_BENCH_default_n{2,4}_v1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 4 --steps 10 --warmup 3
# launch list and full capture of the headline kernel -> profiles/r2_launches_bench_bf16_n74.csv, r2_ncu_full_tcgen05_chain_bf16_n74_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_bf16_n74.csv python bench.py --steps 2 --warmup 1 --no-subrecords --no-cpu-baseline
ncu --set full --clock-control none --import-source on -k regex:chain -s 1 -c 1 -o gpurun_out/r2_chain_full python bench.py --steps 1 --warmup 1 --no-subrecords --no-cpu-baseline
# persistent SVD kernel                    -> profiles/r2_ncu_svd_pair_{2048,4096}_summary.txt, r2_ncu_svd_pair_4096_source.txt
ncu --set full --clock-control none --import-source on -k regex:svd_pair -c 1 -o gpurun_out/r2_svd_pair_4096 python tools/svd_bench.py 4096 --nocheck
python tools/svd_bench.py 1024 2048 4096 ; python tools/qr_bench.py 2048x1024 4096x4096
# chained-kernel experiments               -> profiles/r2_bench_bf16_n74_{chain2,chain4_multicast,l2policy_*}.json
TNB200_CHAIN_CL=4 TNB200_CHAIN_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-subrecords --no-cpu-baseline
TNB200_CHAIN_L2=1 python bench.py --steps 10 --warmup 3 --no-subrecords --no-cpu-baseline      # masks 1, 2, 4
# sanitizers                               -> profiles/r2_sanitizer_*.txt
bash tools/gpu_sanitize_r2.sh
EOF
