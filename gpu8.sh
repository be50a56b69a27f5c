timeout 400 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_igemm.py -q -m gpu 2>&1 | tail -6
for cfg in "--emulate bf16 --gain 0.25 --bn-eval" "--bn-eval" "--emulate bf16 --gain 0.25 --bn-eval --enc resnet18dilated --fc 512 --hw 96 --n 3"; do timeout 300 python tools/debug_parity.py --brief $cfg 2>&1 | grep -v Warn | tail -20; done
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 460 -c 460 --csv --log-file gpurun_out/launches_r1.csv python tools/one_step.py 2 > gpurun_out/one_step.log 2>&1
tail -2 gpurun_out/one_step.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 150 -c 4 -o gpurun_out/igemm_r1 -f python tools/one_step.py 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_kernel -s 4 -c 4 -o gpurun_out/bnbwd_r1 -f python tools/one_step.py 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 0 -c 3 -o gpurun_out/wgrad_r1 -f python tools/one_step.py 1 > /dev/null 2>&1
ls -la gpurun_out
