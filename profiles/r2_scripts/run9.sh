mkdir -p gpurun_out
echo "== tests (fused reduce for residual-block outputs is now on by default)"
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_e2e.py tests/test_gpu_north_star.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
for sw in "SSEG_FUSE_BNBWD_RES=1" "SSEG_FUSE_BNBWD_RES=0"; do
  echo "[$sw]"; env $sw timeout 120 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
done
echo "== gemm micro, defaults"; timeout 200 python tools/gemm_micro.py 2>&1 | tail -14
echo "== gemm micro, CTA pairs"; SSEG_IGEMM_2CTA=2 timeout 200 python tools/gemm_micro.py 2>&1 | tail -14
echo "== gemm micro, 128-wide tiles only"; SSEG_IGEMM_N256=0 SSEG_WGRAD_N256=0 timeout 200 python tools/gemm_micro.py 2>&1 | tail -14
echo "== per-launch GEMM list of the step"; timeout 200 python tools/step_breakdown.py --detail --top 12 2>&1 | head -100 > gpurun_out/breakdown_detail.txt; head -20 gpurun_out/breakdown_detail.txt
