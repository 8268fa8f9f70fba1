# development round for the column-streaming kernels (cs_spread.cuh / cs_interp.cuh)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
export NUFFT_B200_CS=1
timeout 600 python tools/wp_check.py > gpurun_out/cs_check.log 2>&1
tail -12 gpurun_out/cs_check.log
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/cs_c3.log 2>&1
tail -4 gpurun_out/cs_c3.log
timeout 300 python tools/run_c3.py --iters 2 --dist clustered > gpurun_out/cs_c3_clustered.log 2>&1
tail -2 gpurun_out/cs_c3_clustered.log
if [ "${1:-ncu}" = "ncu" ]; then
  bash tools/gpu_ncu.sh cs_spread cs_spread 0
  bash tools/gpu_ncu.sh cs_interp cs_interp 0
fi
