#!/bin/bash
SSEG_IGEMM_PERSISTENT=2 SSEG_IGEMM_PERSISTENT_CTAS=5 timeout 120 python -m pytest tests/test_gpu_igemm.py -q -x 2>&1 | tail -3
SSEG_IGEMM_PERSISTENT=148 timeout 150 python -m pytest tests/test_gpu_e2e.py -q -x 2>&1 | tail -3
for t in 0 148 300 600; do
  echo "persistent threshold $t"
  SSEG_IGEMM_PERSISTENT=$t timeout 90 python tools/step_breakdown.py --replay-only 2>&1 | tail -1
done
