timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_elementwise.py -q -m gpu -k "fused_sgd or batched" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-900
