# the reference's benchmark protocol on B200 (tools/run_benchmarks.py): .dat curves under profiles/bench_dat
set -x
mkdir -p gpurun_out/bench_dat
timeout 900 python -m pytest tests -m gpu -q -x -k "kernels_and_evalmodes or half_supports or matrix or ntransforms or fast_path" 2>&1 | tail -3
timeout 1500 python tools/run_benchmarks.py --types ComplexF64 Float64 ComplexF32 --samples 7 --out gpurun_out/bench_dat 2>&1 | grep -v "^\[" | grep -E "16777216|167772160|1678 |wrote"
timeout 900 python tools/run_benchmarks.py --types ComplexF32 --sigma 2 --samples 7 --out gpurun_out/bench_dat 2>&1 | grep -E "16777216|167772160|1678 |wrote"
