# multi-GPU (2 ranks, one process): results against the single-GPU plan, C5 strong-scaling timing; single-GPU suite on the new build
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python tools/mgpu_check.py --gpus 2 2>&1 | tail -12 | tee gpurun_out/r2c_mgpu_check.log
timeout 600 python tools/mgpu_check.py --gpus 2 --modes 0 --time-modes 512 --time-np 134217728 --iters 3 2>&1 | tail -6 | tee gpurun_out/r2c_mgpu_c5.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2c_tests.log
