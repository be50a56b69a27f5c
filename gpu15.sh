timeout 600 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_elementwise.py -q -m gpu 2>&1 | grep -v Warn | tail -4
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -k "modules_called or running_stats or graph_replay or wiring_bn_eval_r18" 2>&1 | grep -v Warn | tail -6
timeout 400 python tools/step_breakdown.py --top 14 2>&1 | tail -22
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 462 -c 462 --csv --log-file gpurun_out/launches_r1b.csv python tools/one_step.py 2 > gpurun_out/one_step.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 56 -c 6 -o gpurun_out/igemm_convlast_r1 -f python tools/one_step.py 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_kernel -s 8 -c 12 -o gpurun_out/bnbwd_big_r1 -f python tools/one_step.py 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 3 -c 3 -o gpurun_out/wgrad_convlast_r1 -f python tools/one_step.py 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
