SSEG_PDL=1 timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warn | tail -8
SSEG_PDL=1 timeout 600 python bench.py 2>&1 | tail -1
