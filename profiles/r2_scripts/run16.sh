#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (full suite)"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30
echo "== timeline (graph replay, no SGD)"
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r50_final.csv 2>&1 | sed -n 3,6p
python - <<'P'
import csv
rows=list(csv.DictReader(open('gpurun_out/timeline_r50_final.csv')))
for r in rows[:14]:
    print("%8.1f %8.1f  st %-4s grid %-12s %s"%(float(r['start_us']),float(r['dur_us']),r['stream'],r['grid'],r['name'][:40]))
P
for v in "SSEG_X=1" "SSEG_SPLIT_PREP=0" "SSEG_PREP_SIDE_BLOCKS=16"; do
echo "== bench $v"
env $v timeout 300 python bench.py --steps 50 --warmup 5 --no-gpu-context 2>&1 | tail -1 | cut -c1-330
done
