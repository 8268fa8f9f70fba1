# ncu --set full capture of one kernel (regex $1) of the C3 workload; keeps CSV/text pages (the .ncu-rep only if small)
set -x
K=${1:-cs_spread}
TAG=${2:-$K}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${3:-1} -c 1 -f -o /tmp/$TAG python tools/run_c3.py --iters 2 ${4:-} > gpurun_out/ncu_$TAG.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>&1
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_src.csv 2>&1
python tools/ncu_lines.py /tmp/${TAG}_src.csv 0.004 > gpurun_out/${TAG}_lines.txt 2>&1
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source sass > /tmp/${TAG}_sass.csv 2>&1
python tools/ncu_sass.py /tmp/${TAG}_sass.csv 40 > gpurun_out/${TAG}_sass.txt 2>&1
ls -la /tmp/$TAG.ncu-rep
SZ=$(stat -c %s /tmp/$TAG.ncu-rep); if [ "$SZ" -lt 30000000 ]; then cp /tmp/$TAG.ncu-rep gpurun_out/; fi
head -5 gpurun_out/${TAG}_lines.txt
