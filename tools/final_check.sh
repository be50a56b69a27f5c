#!/bin/bash
# The state a round is closed on, on ONE B200 box (gpurun -- 'bash tools/final_check.sh'): full GPU test suite, smoke(),
# a CUPTI timeline of one graph replay and the default bench line. Output -> profiles/rN_final_bench.log.
mkdir -p gpurun_out
echo "== gpu tests (full suite)"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -8
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
echo "== timeline (graph replay, no SGD)"
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r50_final.csv > gpurun_out/timeline_r50_final.txt 2>&1
sed -n 3,6p gpurun_out/timeline_r50_final.txt
echo "== bench (default flags)"
timeout 600 python bench.py 2>&1 | tail -1
