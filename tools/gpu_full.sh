# full GPU round: complete pytest -m gpu, smoke, chunk sweep, bench line, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/full_tests.log
cat gpurun_out/full_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for ch in 128 512 1024 4096; do
  echo "chunk $ch"; NUFFT_B200_CS_CHUNK=$ch timeout 300 python tools/run_c3.py --iters 2 2>&1 | tail -1 | cut -c1-400
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cs.json 2> gpurun_out/bench_cs.err
tail -c 3000 gpurun_out/bench_cs.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cs.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches_cs.csv
