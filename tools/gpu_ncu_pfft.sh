set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pfft_pass -s 6 -c 3 -f -o /tmp/pfft python tools/run_c3.py --iters 2 > gpurun_out/ncu_pfft.log 2>&1
ncu -i /tmp/pfft.ncu-rep --page details > gpurun_out/pfft_details.txt 2>&1
ncu -i /tmp/pfft.ncu-rep --page raw --csv > gpurun_out/pfft_raw.csv 2>&1
