#!/bin/bash
# Profiling pass of one round on a B200 box (run under gpurun from the repo root; ~4 GPU-minutes):
#   bash tools/profile_round.sh r2
# 1. launch list of ONE eager training step (cold-cache per-launch durations; shares must agree with the live bench)
# 2. `ncu --set full` captures of the kernels named in profiles/<round>_summary.md as next targets
# 3. the raw-page metrics the roofline / traffic figures come from, as small CSVs that can be committed to profiles/
# Numbers printed by anything that runs under ncu are never bench values.
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
STEP="python tools/one_step.py 2"

# one step = the second eager run: skip the launches of the first (count them with the library's own counter)
N=$(timeout 120 python tools/one_step.py 1 | sed -n 's/^launches\/step \([0-9]*\).*/\1/p')
echo "launches per step: $N"
timeout 300 $NCU --metrics gpu__time_duration.sum -s ${N:-460} -c ${N:-460} --csv --log-file $OUT/launches_${TAG}.csv $STEP > $OUT/one_step_${TAG}.log 2>&1

cap() {  # cap <name> <kernel regex> <skip> <count>
  timeout 240 $NCU --set full --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/$1_${TAG} $STEP > /dev/null 2>&1
  ncu -i $OUT/$1_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > $OUT/$1_${TAG}_metrics.csv
  echo "== $1"; cat $OUT/$1_${TAG}_metrics.csv | head -8
}
cap weights   weights_batched_kernel 0 2     # re-layout of all conv weights (prep) and of all weight gradients
cap bnapply   bn_apply_kernel        2 2     # conv2 / conv3 of the stem: the largest activations
cap bnbwd     bn_bwd_kernel          2 2
cap igemm_mid "igemm_kernel<128"     20 3    # mid-size convolutions (layer3), where fixed per-CTA cost dominates
cap wgrad_mid "wgrad_kernel<128"     20 3
