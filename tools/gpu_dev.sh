# development round on the GPU box (dev library: NUFFT_DEV_M=4 build)
set -x
export NUFFT_B200_RT=1
timeout 900 python tools/quick_check.py > gpurun_out/dev_quick.log 2>&1
grep -c "^ok" gpurun_out/dev_quick.log; grep -E "FAIL|ALL OK|FAILURES|Error" gpurun_out/dev_quick.log | head
timeout 900 python -m pytest tests -m gpu -q -x -k "3d_matrix or callbacks or ntransforms or large or chunk or pruned" 2>&1 | tail -8 > gpurun_out/dev_tests.log
cat gpurun_out/dev_tests.log
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/dev_c3.log 2>&1
tail -2 gpurun_out/dev_c3.log
