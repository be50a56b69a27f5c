#!/bin/bash
# Profiling pass of one round on a B200 box (run under gpurun from the repo root; ~6 GPU-minutes):
#   bash tools/profile_round.sh r2
# 1. launch list of ONE eager training step (cold-cache per-launch durations; shares must agree with the live bench)
# 2. `ncu --set full` captures of every kernel family that takes >= 1 % of the step, reduced to the raw-page metrics the
#    roofline / traffic figures come from (small CSVs that can be committed to profiles/; the .ncu-rep files stay in scratch)
# Numbers printed by anything that runs under ncu are never bench values.
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
STEP="python tools/one_step.py 2"

# one step = the second eager run: skip the launches of the first (count them with the library's own counter)
N=$(timeout 120 python tools/one_step.py 1 | sed -n 's/^launches\/step \([0-9]*\).*/\1/p')
echo "launches per step: $N"
timeout 300 $NCU --metrics gpu__time_duration.sum -s ${N:-400} -c ${N:-400} --csv --log-file $OUT/launches_${TAG}.csv $STEP > $OUT/one_step_${TAG}.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.DictReader(l for l in open("$OUT/launches_${TAG}.csv") if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].replace("void sseg::", "").split("(")[0][:48]
    d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
print("launch list: %d launches, %.3f ms" % (len(rows), tot / 1e3))
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("  %-50s %4d %9.1f us %5.1f%%" % (k, c, us, 100 * us / tot))
PY

cap() {  # cap <name> <kernel regex> <skip> <count>
  # template kernels are told apart by their arguments: match on the demangled name
  timeout 200 $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o $OUT/$1_${TAG} python tools/one_step.py 1 > /dev/null 2>&1
  ncu -i $OUT/$1_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > $OUT/ncu_$1_${TAG}.csv
  echo "== $1"; cut -d, -f1-9 $OUT/ncu_$1_${TAG}.csv | head -5
}
cap igemm256   "igemm_kernel<.int.256"      0 3     # conv_last / deepsup / layer4 3x3 (128 x 256 tiles)
cap igemm128   "igemm_kernel<.int.128, .int.3"   20 3    # mid-size convolutions
cap igemm64    "igemm_kernel<.int.64"       6 3
cap wgrad256   "wgrad_kernel<.int.256"      0 2
cap wgrad128   "wgrad_kernel<.int.128"      20 3
cap bnapply    "bn_apply_kernel"       2 3
cap bnbwdapply "bn_bwd_kernel<.bool.1" 4 3
cap weights    "weights_batched_kernel" 0 1
cap avgpoolbwd "avgpool_bwd_kernel"    0 1
cap stem       "stem_conv"             0 2
