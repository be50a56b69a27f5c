#!/bin/bash
OUT=gpurun_out; TAG=r2; mkdir -p $OUT
NCU="ncu --clock-control none"
cap() {
  timeout 60 $NCU --set full --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o $OUT/$1_${TAG} python tools/one_step.py 1 > $OUT/cap_$1.log 2>&1
  ncu -i $OUT/$1_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > $OUT/ncu_$1_${TAG}.csv
  echo "== $1"; cut -d, -f1-6 $OUT/ncu_$1_${TAG}.csv | head -4
}
cap igemm64    "igemm_kernel<.int.64"          6 3
cap igemm128d  "igemm_kernel<.int.128, .int.6" 4 3
cap bnfinal    "bn_finalize_kernel"            10 2
cap bnbwdred   "bn_bwd_kernel<.bool.0"         2 2
