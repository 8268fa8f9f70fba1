set -x
timeout 900 python -m pytest tests -m gpu -q -x -k "kernels_and_evalmodes or half_supports or 3d_matrix or 2d_matrix or 1d_matrix or ntransforms" 2>&1 | tail -4
timeout 1500 python tools/run_benchmarks.py --types ComplexF64 Float64 --samples 7 --out gpurun_out/bench_dat 2>&1 | grep -v "^\[" | tail -26
