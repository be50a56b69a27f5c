#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
for v in "SSEG_PEER_APPLY_BLOCKS=2" "SSEG_PEER_APPLY_BLOCKS=1" "SSEG_PEER_FUSE_BWD=0" "SSEG_PEER_APPLY_BLOCKS=2"; do
echo "== bench N=2 $v"
env $v timeout 400 $RUN 29513 bench.py --gpus 2 --steps 60 --warmup 8 --no-gpu-context 2>&1 | tail -1 | cut -c1-260
done
echo "== dist_check with the fused apply"
timeout 300 $RUN 29517 tools/dist_check.py 2>&1 | tail -4
