timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warn | tail -6
timeout 400 python tools/step_breakdown.py --top 16 2>&1 | tail -22
timeout 600 python bench.py 2>&1 | tail -1 | cut -c1-1000
