#!/bin/bash
# Eight-GPU check (gpurun --gpus 8): out-of-phase batch shapes through the public API, then the bench at N = 8 (push
# protocol, the default, and the flag protocol for comparison) and at N = 4.
mkdir -p gpurun_out
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
RUN4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port"
echo "== ranks with alternating batch shapes, 8 ranks"
timeout 240 $RUN8 29531 tools/dist_shapes.py 2>&1 | grep -E "DIST_SHAPES_OK|rank 0|rank 7|Error|error|Traceback" | head -8
echo "== bench N=8"
timeout 400 $RUN8 29532 bench.py --gpus 8 --steps 50 --warmup 6 2>&1 | tail -1 | cut -c1-1300
echo "== bench N=8, flag protocol (SSEG_PEER_LL=0)"
SSEG_PEER_LL=0 timeout 400 $RUN8 29533 bench.py --gpus 8 --steps 50 --warmup 6 2>&1 | tail -1 | cut -c1-300
echo "== bench N=4"
timeout 400 $RUN4 29534 bench.py --gpus 4 --steps 50 --warmup 6 2>&1 | tail -1 | cut -c1-300
