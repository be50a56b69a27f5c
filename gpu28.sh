timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_e2e.py -q -m gpu -k "finalize_apply or wiring or train_mode or running_stats or graph_replay" 2>&1 | grep -v Warn | tail -4
timeout 400 python tools/step_breakdown.py --top 4 2>&1 | tail -3
SSEG_FUSE_FINALIZE=0 timeout 400 python tools/step_breakdown.py --top 3 2>&1 | tail -1
