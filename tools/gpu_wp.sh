# development round for the warp-private-tile kernels (wp_spread.cuh / wp_interp.cuh)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
export NUFFT_B200_WP=1
timeout 600 python tools/wp_check.py > gpurun_out/wp_check.log 2>&1
tail -12 gpurun_out/wp_check.log
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/wp_c3.log 2>&1
tail -5 gpurun_out/wp_c3.log
NUFFT_B200_WP=0 timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/v2_c3.log 2>&1
tail -3 gpurun_out/v2_c3.log
if [ "${1:-ncu}" = "ncu" ]; then
  bash tools/gpu_ncu.sh wp_spread wp_spread 0
  bash tools/gpu_ncu.sh wp_interp wp_interp 0
fi
