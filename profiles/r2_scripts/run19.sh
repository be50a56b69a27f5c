#!/bin/bash
# second half of the profiling pass (template kernels, matched on demangled names) + the state after the last kernel edit
OUT=gpurun_out; TAG=r2; mkdir -p $OUT
NCU="ncu --clock-control none"
cap() {
  timeout 200 $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o $OUT/$1_${TAG} python tools/one_step.py 1 > $OUT/cap_$1.log 2>&1
  ncu -i $OUT/$1_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > $OUT/ncu_$1_${TAG}.csv
  echo "== $1"; cut -d, -f1-9 $OUT/ncu_$1_${TAG}.csv | head -5; tail -2 $OUT/cap_$1.log | cut -c1-200
}
cap igemm256   "igemm_kernel<256"      0 3
cap igemm128   "igemm_kernel<128, 3"   20 3
cap wgrad256   "wgrad_kernel<256"      0 2
cap wgrad128   "wgrad_kernel<128"      20 3
cap bnbwdapply "bn_bwd_kernel<true>|bn_bwd_kernel<\\(bool\\)1>|bn_bwd_kernel<1>" 4 3
cap avgpoolbwd "avgpool_bwd_kernel"    0 1
echo "== elementwise tests"
timeout 600 python -m pytest tests/test_gpu_elementwise.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
echo "== bench"
timeout 400 python bench.py --no-gpu-context 2>&1 | tail -1
