timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_elementwise.py tests/test_gpu_e2e.py -q -m gpu 2>&1 | grep -v Warn | tail -8
timeout 400 python tools/step_breakdown.py --top 10 2>&1 | tail -14
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-600
