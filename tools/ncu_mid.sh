#!/bin/bash
# ncu --set full of the latency-bound mid-size launches with a WARM L2 (--cache-control none: in the real step their
# operands were just written by the previous kernel), run under gpurun from the repo root:  bash tools/ncu_mid.sh r2
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none --cache-control none"
STEP="python tools/one_step.py 1"
cap() {  # cap <name> <kernel regex> <skip> <count>
  timeout 300 $NCU --set full --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/$1_${TAG} $STEP > $OUT/$1_${TAG}.log 2>&1
  ncu -i $OUT/$1_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > $OUT/$1_${TAG}_metrics.csv
  echo "== $1"; cat $OUT/$1_${TAG}_metrics.csv | head -8
}
cap igemm_l3   "igemm_kernel"   29 3     # layer3 block 1: 1x1 1024->256, 3x3 (dil 2) 256->256, 1x1 256->1024
cap bnapply_l3 "bn_apply_kernel" 27 2
cap bnbwd_mid  "bn_bwd_kernel"   40 3
