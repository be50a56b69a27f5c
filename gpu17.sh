#!/bin/bash
# new code first (HRNet path), then the whole suite as a regression check, then smoke
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -p no:cacheprovider -s \
  -k "channel_counts or stride2 or fused_bn_backward or exchange_unit or hrnet" > gpurun_out/r1_hrnet_tests.log 2>&1
echo "new tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r1_hrnet_tests.log | tail -3
grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/r1_hrnet_tests.log | head -20
grep -E "grad_rel_median|agreement" gpurun_out/r1_hrnet_tests.log | cut -c1-700
timeout 150 python -m pytest tests -m gpu -q -p no:cacheprovider \
  -k "not (channel_counts or stride2 or fused_bn_backward or exchange_unit or hrnet)" > gpurun_out/r1_regression_tests.log 2>&1
echo "regression rc=$?"; tail -3 gpurun_out/r1_regression_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
