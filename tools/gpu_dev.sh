# development round on the GPU box (dev library: NUFFT_DEV_M=4 build)
set -x
timeout 900 python -m pytest tests -m gpu -q -x -k "pruned or matrix or callbacks or fftshift or ntransforms or chunk or large or empty" 2>&1 | tail -15 > gpurun_out/dev_tests.log
cat gpurun_out/dev_tests.log
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/dev_c3.log 2>&1
tail -2 gpurun_out/dev_c3.log
NUFFT_B200_RT=1 timeout 300 python tools/run_c3.py --iters 3 2>&1 | tail -1
