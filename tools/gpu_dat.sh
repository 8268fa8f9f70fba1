set -x
mkdir -p gpurun_out profiles/bench_dat
timeout 1500 python tools/run_benchmarks.py --types ComplexF64 Float64 --samples 7 --out gpurun_out/bench_dat 2>&1 | grep -v "^\[" | tail -30
timeout 900 python tools/run_benchmarks.py --types ComplexF32 --fast --samples 7 --out gpurun_out/bench_dat 2>&1 | tail -14
timeout 900 python tools/run_benchmarks.py --types ComplexF32 --fast --sigma 2 --samples 7 --out gpurun_out/bench_dat 2>&1 | tail -14
