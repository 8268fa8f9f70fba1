set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 21 python tools/run_c3.py --iters 2 2>&1 | grep -E "^\s+(void |nufft::)|gpu__time" | paste - - | awk '{print $NF, $1, $2, $3}' > gpurun_out/sp_times.log
cat gpurun_out/sp_times.log
