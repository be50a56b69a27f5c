timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_e2e.py -q -m gpu -k "fused or wiring or train_mode_bn" 2>&1 | grep -v Warn | tail -4
timeout 400 python tools/step_breakdown.py --top 8 2>&1 | tail -11
SSEG_FUSE_BNBWD=0 timeout 400 python tools/step_breakdown.py --top 3 2>&1 | tail -2
