mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_blocksparse.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/kernel_bench.py --svd 1024 2048 4096 2>&1 | tail -4
