#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_widen_hrnet.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r1_hrnet_tests3.log 2>&1
echo "rc=$?"; grep -E "passed|failed" gpurun_out/r1_hrnet_tests3.log | tail -2
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r1_hrnet_tests3.log | head -30
grep -E "grad_rel_median|agreement" gpurun_out/r1_hrnet_tests3.log | cut -c1-900
