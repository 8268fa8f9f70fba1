# 1 GPU, release library: the whole GPU suite, smoke, the default bench line
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu > gpurun_out/r2v_tests.log 2>&1; echo TESTS $?; tail -4 gpurun_out/r2v_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE ok')" 2>&1 | tail -2
timeout 200 python bench.py > gpurun_out/r2v_bench_1gpu.json 2> gpurun_out/r2v_bench_1gpu.err; echo BENCH $?; tail -c 1500 gpurun_out/r2v_bench_1gpu.json | head -c 900
