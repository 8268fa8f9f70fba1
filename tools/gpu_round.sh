# full round on the GPU box: parity tests, smoke, bench, launch list, ncu of the two dominant kernels
set -x
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
timeout 300 python tools/run_c3.py --iters 3 > gpurun_out/${TAG}_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_smoke.log; tail -3 gpurun_out/${TAG}_bench.log; tail -4 gpurun_out/${TAG}_c3.log
