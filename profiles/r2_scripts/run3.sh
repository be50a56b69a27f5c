mkdir -p gpurun_out
echo "== igemm tests, persistent forced"; SSEG_IGEMM_PERSISTENT=2 SSEG_IGEMM_PERSISTENT_CTAS=5 timeout 300 python -m pytest tests/test_gpu_igemm.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "== igemm tests, default"; timeout 300 python -m pytest tests/test_gpu_igemm.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for sw in "SSEG_IGEMM_DEEP=0" "SSEG_IGEMM_DEEP=1" "SSEG_IGEMM_PERSISTENT=149" "SSEG_IGEMM_PERSISTENT=300" "SSEG_IGEMM_PERSISTENT=600" "SSEG_IGEMM_PERSISTENT=1"; do
  echo "[$sw]"; env $sw timeout 120 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
done
echo "== north star tests"; timeout 600 python -m pytest tests/test_gpu_north_star.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -15
echo "== grad parity c1_deepsup"; timeout 200 python tools/grad_parity.py --dec c1_deepsup --repeat 3 --top 8 2>&1 | tail -16
echo "== grad parity ppm_deepsup n=8 hw=96"; timeout 200 python tools/grad_parity.py --n 8 --hw 96 --repeat 2 --top 5 2>&1 | tail -12
