mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_c.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_c.log
grep -v -i warn gpurun_out/dist_check_c.log | tail -8
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2_c.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_c.log
grep -v -i warn gpurun_out/bench_n2_c.log | tail -3 | cut -c1-420
