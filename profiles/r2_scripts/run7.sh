mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_north_star.py tests/test_gpu_elementwise.py tests/test_gpu_igemm.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
echo "== replay"; timeout 120 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-context 2>&1 | tail -2
echo "== timeline"; timeout 300 python tools/timeline.py --with-sgd --out gpurun_out/timeline_r50_v2.csv 2>&1 | tail -42
