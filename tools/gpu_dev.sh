# development round on the GPU box (dev library: NUFFT_DEV_M=4 build)
set -x
timeout 900 python -m pytest tests -m gpu -q -x -k "pruned or 1d_matrix or callbacks or fftshift" 2>&1 | tail -15 > gpurun_out/dev_tests.log
cat gpurun_out/dev_tests.log
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/dev_c3.log 2>&1
tail -2 gpurun_out/dev_c3.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pfft_pass -s 6 -c 6 python tools/run_c3.py --iters 2 2>&1 | grep -E "pfft_pass|gpu__time|dram__bytes|bank_conf|warps_active" > gpurun_out/dev_pfft_times.log
cat gpurun_out/dev_pfft_times.log
