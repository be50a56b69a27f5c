mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_n4.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n4.log
grep -v -i warn gpurun_out/bench_n4.log | tail -3 | cut -c1-420
