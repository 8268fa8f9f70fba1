# baseline numbers of both kernel families on the GPU box
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/base_c3_v2.log 2>&1
tail -2 gpurun_out/base_c3_v2.log
NUFFT_B200_RT=1 timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/base_c3_rt.log 2>&1
tail -2 gpurun_out/base_c3_rt.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/base_tests.log
tail -5 gpurun_out/base_tests.log
