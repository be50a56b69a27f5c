#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
echo "== igemm + whole-step tests with the final library"
timeout 600 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
echo "== two-process SyncBN tests"
timeout 600 python -m pytest tests/test_gpu_north_star.py -m gpu -q -s -p no:cacheprovider -k two_processes 2>&1 | grep -E "passed|failed|OK|Error" | tail -5
echo "== bench N=2 (final)"
timeout 400 $RUN 29513 bench.py --gpus 2 2>&1 | tail -1
