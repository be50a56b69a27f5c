mkdir -p gpurun_out
echo "== trace"; SSEG_LIB=libsseg_b200_trace.so timeout 300 python tools/trace_igemm.py 2>&1 | tail -30
echo "== e2e gpu tests"; timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_north_star.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -3
