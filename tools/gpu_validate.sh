#!/bin/bash
# Round-end validation on a B200 box (run under gpurun from the repo root): GPU tests, smoke, both bench arms.
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warn | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
