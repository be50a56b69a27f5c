#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root, ~6 GPU-minutes):
# everything that was written after this round's GPU budget was spent gets its first run, then each opt-in switch is
# timed against the default schedule so the defaults can be flipped on evidence.
mkdir -p gpurun_out
echo "== gated tests"
SSEG_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_widen_hrnet.py -m gpu -q -p no:cacheprovider -s \
  > gpurun_out/experimental_tests.log 2>&1
grep -E "passed|failed" gpurun_out/experimental_tests.log | tail -2
grep -E "^FAILED|^ERROR|^E  " gpurun_out/experimental_tests.log | head -30
echo "== training step, CUDA-graph replay"
for sw in "" "SSEG_BRANCH_STREAMS=1" "SSEG_OVERLAP_RELAYOUT=1" "SSEG_COOP_BN=1" "SSEG_BRANCH_STREAMS=1 SSEG_OVERLAP_RELAYOUT=1 SSEG_COOP_BN=1"; do
  echo "[$sw]"; env $sw timeout 120 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
done
echo "== HRNetV2+C1 training step"
for sw in "" "SSEG_BRANCH_STREAMS=1"; do
  echo "[$sw]"; env $sw timeout 200 python tools/step_breakdown.py --net hrnet --replay-only 2>&1 | tail -1
done
echo "== inference (configs 2 and 5)"
for sw in "SSEG_FOLD_BN_EVAL=0" "SSEG_FOLD_BN_EVAL=1"; do
  env $sw timeout 120 python tools/infer_bench.py --net r18ppm 2>&1 | tail -1
  env $sw timeout 200 python tools/infer_bench.py --net hrnet --multiscale 2>&1 | tail -1
done
