#!/bin/bash
# Second GPU call of the next round, on TWO GPUs (gpurun --gpus 2 -- 'bash tools/next_round_two_gpu_run.sh'; ~4 box-minutes
# x 2): the data-parallel step with the default schedule and with the fused conv+BN kernels pooling the SyncBN statistics
# over NVLink inside the kernel. Each run is wrapped in `timeout` (cross-GPU spin-waits must not hang the box).
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
echo "== dist_check, default schedule"
timeout 240 $RUN 29511 tools/dist_check.py 2>&1 | grep -E "loss|cosine|running_mean|DIST_CHECK_OK|Error|error" | tail -8
echo "== dist_check, SSEG_COOP_BN=1"
SSEG_COOP_BN=1 timeout 240 $RUN 29512 tools/dist_check.py 2>&1 | grep -E "loss|cosine|running_mean|DIST_CHECK_OK|Error|error" | tail -8
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
echo "== bench N=2"
timeout 300 $RUN 29513 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-400
echo "== bench N=2, SSEG_COOP_BN=1"
SSEG_COOP_BN=1 timeout 300 $RUN 29514 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-400
