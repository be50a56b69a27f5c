#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (full suite)"
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3
echo "== timeline (graph replay, no SGD)"
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r50_final.csv 2>&1 | head -24
echo "== bench (default flags)"
timeout 600 python bench.py 2>&1 | tail -1
echo "== bench SSEG_SPLIT_PREP=0"
SSEG_SPLIT_PREP=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-gpu-context 2>&1 | tail -1 | cut -c1-330
echo "== e2e breakdown"
timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -10
