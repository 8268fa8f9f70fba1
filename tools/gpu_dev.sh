# development round on the GPU box (dev library: NUFFT_DEV_M=4 build)
set -x
timeout 900 python -m pytest tests -m gpu -q -x -k "matrix_points or nfft_frontend or pruned or 1d_matrix or 3d_matrix or callbacks" 2>&1 | tail -25 > gpurun_out/dev_tests.log
cat gpurun_out/dev_tests.log
