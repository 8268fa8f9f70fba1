set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.log 2>&1
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/r1_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spread_sm_kernel|interp_sm_kernel" -c 2 -o gpurun_out/r1_v2_spread_interp python tools/run_c3.py --iters 1 > gpurun_out/r1_ncu.log 2>&1
tail -3 gpurun_out/r1_tests.log; tail -3 gpurun_out/r1_bench.log; tail -4 gpurun_out/r1_c3.log
