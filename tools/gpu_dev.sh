# development round on the GPU box (dev library: NUFFT_DEV_M=4 build)
set -x
timeout 900 python tools/quick_check.py > gpurun_out/dev_quick.log 2>&1
tail -30 gpurun_out/dev_quick.log
timeout 900 python -m pytest tests -m gpu -q -x -k "3d_matrix or 1d_matrix or ntransforms or clustered or callbacks or large or chunk" 2>&1 | tail -15 > gpurun_out/dev_tests.log
cat gpurun_out/dev_tests.log
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/dev_c3.log 2>&1
tail -4 gpurun_out/dev_c3.log
