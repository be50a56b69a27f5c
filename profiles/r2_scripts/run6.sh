mkdir -p gpurun_out
echo "== kernel tests"; timeout 600 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_elementwise.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "== igemm tests, N=256 forced"; SSEG_IGEMM_N256_KSTEPS=1 SSEG_IGEMM_N256_TILES=1 timeout 300 python -m pytest tests/test_gpu_igemm.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
echo "== trace"; SSEG_LIB=libsseg_b200_trace.so timeout 300 python tools/trace_igemm.py 2>&1 | grep -v "CTA lifetime"
for sw in "SSEG_FUSE_BNFIN=1" "SSEG_FUSE_BNFIN=0" "SSEG_IGEMM_N256=0" "SSEG_NTILE_THRESH=160"; do
  echo "[$sw]"; env $sw timeout 120 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
done
echo "== e2e + north star + syncbn"; timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_north_star.py tests/test_gpu_syncbn_module.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-context 2>&1 | tail -2
echo "== timeline"; timeout 300 python tools/timeline.py --with-sgd --out gpurun_out/timeline_r50_v2.csv 2>&1 | tail -45
