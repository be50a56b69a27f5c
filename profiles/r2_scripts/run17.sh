#!/bin/bash
for v in "SSEG_PREP_SIDE_BLOCKS=296" "SSEG_PREP_SIDE_BLOCKS=148" "SSEG_PREP_SIDE_BLOCKS=592"; do
echo "== bench $v"
env $v timeout 300 python bench.py --steps 50 --warmup 5 --no-gpu-context 2>&1 | tail -1 | cut -c1-330
done
