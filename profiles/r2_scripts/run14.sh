#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
echo "== dist_check (push protocol)"
timeout 300 $RUN 29517 tools/dist_check.py 2>&1 | tail -5
echo "== two-process SyncBN tests"
timeout 600 python -m pytest tests/test_gpu_north_star.py -m gpu -q -s -p no:cacheprovider -k two_processes 2>&1 | grep -E "passed|failed|OK|Error|error" | tail -6
echo "== shapes"
timeout 300 $RUN 29521 tools/dist_shapes.py 2>&1 | tail -3
for v in "SSEG_PEER_LL=1" "SSEG_PEER_LL=0" "SSEG_PEER_LL=1"; do
echo "== bench N=2 $v"
env $v timeout 400 $RUN 29513 bench.py --gpus 2 --steps 60 --warmup 8 --no-gpu-context 2>&1 | tail -1 | cut -c1-260
done
