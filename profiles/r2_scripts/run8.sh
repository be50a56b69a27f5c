mkdir -p gpurun_out
echo "== wgrad 256-wide tiles forced"; SSEG_WGRAD_N256_TILES=1 timeout 240 python -m pytest tests/test_gpu_igemm.py -m gpu -q -p no:cacheprovider -k wgrad 2>&1 | tail -3
echo "== igemm tests as CTA pairs (cta_group::2), 256-wide tiles forced onto the unit-test shapes"
SSEG_IGEMM_2CTA=1 SSEG_IGEMM_N256_KSTEPS=1 SSEG_IGEMM_N256_TILES=1 timeout 240 python -m pytest tests/test_gpu_igemm.py -m gpu -q -p no:cacheprovider 2>&1 | tail -12
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
echo "== ... and the 128-wide tiles as pairs too"
SSEG_IGEMM_2CTA=2 SSEG_IGEMM_N256_KSTEPS=1 SSEG_IGEMM_N256_TILES=1 timeout 240 python -m pytest tests/test_gpu_igemm.py -m gpu -q -p no:cacheprovider 2>&1 | tail -12
SSEG_IGEMM_2CTA=2 timeout 240 python -m pytest tests/test_gpu_igemm.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
for sw in "SSEG_IGEMM_2CTA=0" "SSEG_IGEMM_2CTA=1" "SSEG_IGEMM_2CTA=2"; do
  echo "[$sw]"; env $sw timeout 120 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
done
echo "== whole-step tests with pairs"; SSEG_IGEMM_2CTA=2 timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider -k "wiring or full_config or graph_replay" 2>&1 | tail -4
for sw in "SSEG_WGRAD_N256=0" "SSEG_WGRAD_N256=1"; do
  echo "[$sw]"; env $sw timeout 120 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
done
echo "== input pipeline"; timeout 300 python tools/input_pipeline_bench.py --steps 30 2>&1 | tail -3
echo "== loss curve on B200"; timeout 400 python tools/loss_curve_b200.py --steps 200 > gpurun_out/loss_curve_b200.txt 2>&1; tail -8 gpurun_out/loss_curve_b200.txt
