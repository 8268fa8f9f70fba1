set -x
bash tools/gpu_ncu.sh spread_sm v2_spread 1
bash tools/gpu_ncu.sh interp_sm v2_interp 1
