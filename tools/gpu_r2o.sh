set -x
mkdir -p gpurun_out
export NUFFT_B200_LIB=$PWD/nonuniformffts.jl_b200/libnufft_b200_dev.so
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "multi_gpu" > gpurun_out/r2t_tests.log 2>&1; echo TESTS $?; tail -5 gpurun_out/r2t_tests.log
timeout 200 python tools/mgpu_check.py --gpus 2 --modes 0 --time-modes 512 --time-np 134217728 --iters 5 --skip-single > gpurun_out/r2t_check_p2p.log 2>&1; echo CHECK $?; tail -3 gpurun_out/r2t_check_p2p.log
