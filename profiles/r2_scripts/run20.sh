#!/bin/bash
OUT=gpurun_out; TAG=r2; mkdir -p $OUT
NCU="ncu --clock-control none"
cap() {
  timeout 200 $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o $OUT/$1_${TAG} python tools/one_step.py 1 > $OUT/cap_$1.log 2>&1
  ncu -i $OUT/$1_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > $OUT/ncu_$1_${TAG}.csv
  echo "== $1"; cut -d, -f1-9 $OUT/ncu_$1_${TAG}.csv | head -5
}
cap igemm256   "igemm_kernel<.int.256"      0 3
cap igemm128   "igemm_kernel<.int.128, .int.3"   20 3
cap wgrad256   "wgrad_kernel<.int.256"      0 2
cap wgrad128   "wgrad_kernel<.int.128"      20 3
