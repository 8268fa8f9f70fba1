set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "matrix or clustered or column_streaming or fast_path" 2>&1 | tail -3 | tee gpurun_out/quick_tests.log
timeout 300 python tools/run_c3.py --iters 4 2>&1 | tail -2 | cut -c1-400 | tee gpurun_out/quick_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_keys|radix|scan" -s 16 -c 16 --csv python tools/run_c3.py --iters 2 2>/dev/null | grep -E "bin_keys|radix|scan" | awk -F'","' '{print substr($5,1,40), $NF}' | tee gpurun_out/quick_setpoints_kernels.log
