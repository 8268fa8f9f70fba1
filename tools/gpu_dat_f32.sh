# reference benchmark protocol (tools/run_benchmarks.py) for the Float32 types with the final build: .dat curves
set -x
mkdir -p gpurun_out/bench_dat
timeout 1200 python tools/run_benchmarks.py --types Float32 ComplexF32 --sigma 2 --fast --samples 7 --out gpurun_out/bench_dat 2>&1 | grep -E "16777216|167772160|1678 |wrote"
timeout 1200 python tools/run_benchmarks.py --types Float32 ComplexF32 --sigma 1.5 --samples 7 --out gpurun_out/bench_dat 2>&1 | grep -E "16777216|167772160|1678 |wrote"
