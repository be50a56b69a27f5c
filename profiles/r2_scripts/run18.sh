#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (full suite)"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -8
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
echo "== timeline (graph replay, no SGD)"
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r50_final.csv > gpurun_out/timeline_r50_final.txt 2>&1
sed -n 3,40p gpurun_out/timeline_r50_final.txt
python - <<'P'
import csv
rows=list(csv.DictReader(open('gpurun_out/timeline_r50_final.csv')))
for r in rows[8:14]:
    print("%8.1f %8.1f  st %-4s grid %-12s %s"%(float(r['start_us']),float(r['dur_us']),r['stream'],r['grid'],r['name'][:40]))
P
echo "== bench (default flags)"
timeout 600 python bench.py 2>&1 | tail -1
