#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
echo "== two-process SyncBN tests"
timeout 600 python -m pytest tests/test_gpu_north_star.py -m gpu -q -s -p no:cacheprovider -k two_processes 2>&1 | grep -E "passed|failed|OK|cosine|loss|Error|error" | tail -12
echo "== shapes"
timeout 300 $RUN 29521 tools/dist_shapes.py 2>&1 | tail -3
echo "== bench N=2 fused peer bwd"
timeout 400 $RUN 29513 bench.py --gpus 2 --steps 60 --warmup 8 --no-gpu-context 2>&1 | tail -1 | cut -c1-700
echo "== bench N=2 SSEG_PEER_FUSE_BWD=0"
SSEG_PEER_FUSE_BWD=0 timeout 400 $RUN 29514 bench.py --gpus 2 --steps 60 --warmup 8 --no-gpu-context 2>&1 | tail -1 | cut -c1-700
echo "== bench N=2 fused again"
timeout 400 $RUN 29516 bench.py --gpus 2 --steps 60 --warmup 8 --no-gpu-context 2>&1 | tail -1 | cut -c1-400
