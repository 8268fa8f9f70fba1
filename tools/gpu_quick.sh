# quick regression after a kernel change: full GPU parity suite + C3 stage times (complex, clustered, real)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/quick_tests.log
timeout 300 python tools/run_c3.py --iters 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/quick_c3.log
timeout 300 python tools/run_c3.py --iters 3 --dist clustered 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/quick_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_keys|radix|scan" -s 20 -c 16 --csv python tools/run_c3.py --iters 2 2>/dev/null | grep -E "bin_keys|radix|scan" | awk -F'","' '{print substr($5,1,40), $NF}' | tee gpurun_out/quick_setpoints_kernels.log
