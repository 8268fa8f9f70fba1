# development round for the column-streaming kernels (cs_spread.cuh / cs_interp.cuh)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q -x -k "fast_path or 3d_matrix or clustered" 2>&1 | tail -8 > gpurun_out/cs_tests.log
cat gpurun_out/cs_tests.log
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/cs_c3.log 2>&1
tail -3 gpurun_out/cs_c3.log
if [ "${1:-ncu}" = "ncu" ]; then
  bash tools/gpu_ncu.sh cs_spread cs_spread 0
  bash tools/gpu_ncu.sh cs_interp cs_interp 0
  head -30 gpurun_out/cs_spread_sass.txt
  head -30 gpurun_out/cs_interp_sass.txt
fi
