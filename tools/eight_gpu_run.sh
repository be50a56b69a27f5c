#!/bin/bash
# Eight-GPU check (gpurun --gpus 8): out-of-phase batch shapes through the public API, then the bench at N = 8.
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
echo "== ranks with alternating batch shapes, 8 ranks"
timeout 240 $RUN 29531 tools/dist_shapes.py 2>&1 | grep -E "DIST_SHAPES_OK|rank 0|rank 7|Error|error|Traceback" | head -8
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
echo "== bench N=8"
timeout 400 $RUN 29532 bench.py --gpus 8 --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-1200
