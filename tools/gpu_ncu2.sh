set -x
bash tools/gpu_ncu.sh rt_spread rt_spread 1
bash tools/gpu_ncu.sh rt_interp rt_interp 1
