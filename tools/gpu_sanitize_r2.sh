#!/bin/bash
# Round 2: compute-sanitizer over the kernels written / changed this round, plus racecheck and synccheck of the chained launch
# (inter-CTA red.release / ld.acquire protocol) that the round-1 verdict asked for.  Output -> gpurun_out/r2_sanitizer_*.txt
export TNB200_CHAIN_FORCE=1
out=gpurun_out
run() { # name tool pytest-args...
  name=$1; tool=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest "$@" -x -q -m gpu 2>&1 | tail -12 > $out/r2_sanitizer_${tool}_${name}.txt
  echo "== $tool $name: $(tail -1 $out/r2_sanitizer_${tool}_${name}.txt)"
}
run svd_qr memcheck tests/test_gpu_split.py -k "persistent or qr_properties or golden_decomp"
run blocksparse memcheck tests/test_gpu_blocksparse.py -k "device_built or dense_equivalence or golden_blocksparse"
run splitk memcheck tests/test_gpu_tensordot.py -k "split_k"
run chain memcheck tests/test_gpu_drivers.py -k "chained"
run chain racecheck tests/test_gpu_drivers.py -k "chained"
run chain synccheck tests/test_gpu_drivers.py -k "chained"
run svd synccheck tests/test_gpu_split.py -k "persistent_pair_kernel and float64 and 300"
