mkdir -p gpurun_out
echo "== igemm tests default"; timeout 300 python -m pytest tests/test_gpu_igemm.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
echo "== igemm tests, N=256 forced"; SSEG_IGEMM_N256_KSTEPS=1 SSEG_IGEMM_N256_TILES=1 timeout 300 python -m pytest tests/test_gpu_igemm.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
echo "== trace"; SSEG_LIB=libsseg_b200_trace.so timeout 300 python tools/trace_igemm.py 2>&1 | grep -v "CTA lifetime"
for sw in "SSEG_IGEMM_N256=1" "SSEG_IGEMM_N256=0" "SSEG_IGEMM_N256_KSTEPS=24" "SSEG_IGEMM_DEEP=0"; do
  echo "[$sw]"; env $sw timeout 120 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
done
echo "== syncbn module + e2e + north star"; timeout 900 python -m pytest tests/test_gpu_syncbn_module.py tests/test_gpu_e2e.py tests/test_gpu_north_star.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "== grad parity c1_deepsup bias-shift 2"; timeout 200 python tools/grad_parity.py --dec c1_deepsup --bias-shift 2 --repeat 2 --top 4 2>&1 | tail -9
echo "== grad parity ppm_deepsup bias-shift 2"; timeout 200 python tools/grad_parity.py --bias-shift 2 --repeat 2 --top 4 2>&1 | tail -9
echo "== grad parity ppm_deepsup bias-shift 2 n=8 hw=96"; timeout 200 python tools/grad_parity.py --bias-shift 2 --n 8 --hw 96 --repeat 2 --top 4 2>&1 | tail -9
echo "== grad parity r50 ppm_deepsup bias-shift 2 n=4 hw=128"; timeout 300 python tools/grad_parity.py --enc resnet50dilated --fc 2048 --bias-shift 2 --n 4 --hw 128 --repeat 2 --top 4 2>&1 | tail -9
