#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -p no:cacheprovider -s \
  -k "channel_counts or hrnet" > gpurun_out/r1_hrnet_tests2.log 2>&1
echo "rc=$?"; grep -E "passed|failed" gpurun_out/r1_hrnet_tests2.log | tail -2
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r1_hrnet_tests2.log | head -30
grep -E "grad_rel_median|agreement" gpurun_out/r1_hrnet_tests2.log | cut -c1-900
