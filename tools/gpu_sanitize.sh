# memcheck of the kernels written this round (thin streaming kernels, chained launch, skinny dot, SVD eig)
export TNB200_CHAIN_FORCE=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_tcgen05.py -x -q -m gpu -k "thin" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_drivers.py -x -q -m gpu -k "chained or compiled" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_tensordot.py tests/test_gpu_split.py -x -q -m gpu -k "skinny or svd_known or golden_decomp" 2>&1 | tail -6
