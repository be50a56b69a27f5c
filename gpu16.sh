mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dist_probe.py all > gpurun_out/dist_probe.log 2>&1
echo "rc=$?" >> gpurun_out/dist_probe.log
grep -v Warn gpurun_out/dist_probe.log | tail -40
