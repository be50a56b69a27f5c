timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warn | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -3
SSEG_PDL=1 timeout 300 python tools/step_breakdown.py --top 8 2>&1 | tail -12
