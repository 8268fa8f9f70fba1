# real Float32 data on the column-streaming kernels: parity sweep + C3-sized timings (real vs complex)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "column_streaming or fast_path or callbacks or ntransforms or real_data" 2>&1 | tail -6 > gpurun_out/real_tests.log
cat gpurun_out/real_tests.log
timeout 300 python tools/run_c3.py --iters 3 --real 2>&1 | tail -2 | cut -c1-500 | tee gpurun_out/real_c3.log
timeout 300 python tools/run_c3.py --iters 3 2>&1 | tail -1 | cut -c1-500 | tee gpurun_out/cplx_c3.log
