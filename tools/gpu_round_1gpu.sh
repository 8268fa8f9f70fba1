# round 2: full GPU parity suite on the release build, smoke, the N = 1 bench line, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12 | tee gpurun_out/r2_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 2500 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/r2_bench_under_ncu.log 2>&1
