#!/bin/bash
# Two-GPU validation + timing (gpurun --gpus 2 -- 'bash tools/two_gpu_run.sh'; every spin-wait in the kernels is bounded,
# every run is wrapped in `timeout`).
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
echo "== two-process SyncBN tests (module level + step program vs the oracle on the concatenated batch)"
timeout 600 python -m pytest tests/test_gpu_north_star.py -m gpu -q -s -p no:cacheprovider -k two_processes 2>&1 | grep -E "passed|failed|OK|cosine|loss|Error|error" | tail -12
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
echo "== ranks with DIFFERENT, alternating batch shapes through the public API (shared peer arena, lazy capture)"
timeout 300 $RUN 29521 tools/dist_shapes.py 2>&1 | tail -6
echo "== bench N=2"
timeout 400 $RUN 29513 bench.py --gpus 2 --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-900
echo "== bench N=2, NCCL SyncBN (SSEG_PEER_SYNC=0)"
SSEG_PEER_SYNC=0 timeout 400 $RUN 29514 bench.py --gpus 2 --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-500
echo "== config 5: HRNetV2-W48 + C1 multi-scale inference, scales sharded over the two GPUs (and on one GPU for comparison)"
timeout 300 python tools/infer_bench.py --net hrnet --multiscale 2>&1 | tail -1
timeout 300 $RUN 29515 tools/infer_bench.py --net hrnet --multiscale 2>&1 | tail -1
