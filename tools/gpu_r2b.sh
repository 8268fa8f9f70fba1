# kernel variants (development builds, HalfSupport(4) only): C3 stage times per variant
set -x
mkdir -p gpurun_out
: > gpurun_out/r2b_c3.log
for t in a b c e; do
  echo "== variant $t" | tee -a gpurun_out/r2b_c3.log
  NUFFT_B200_LIB=$PWD/nonuniformffts.jl_b200/libnufft_b200_$t.so timeout 300 python tools/run_c3.py --iters 4 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/r2b_c3.log
done
for ch in 256 1024; do
  echo "== variant a chunk $ch" | tee -a gpurun_out/r2b_c3.log
  NUFFT_B200_CS_CHUNK=$ch NUFFT_B200_LIB=$PWD/nonuniformffts.jl_b200/libnufft_b200_a.so timeout 300 python tools/run_c3.py --iters 4 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/r2b_c3.log
done
NUFFT_B200_LIB=$PWD/nonuniformffts.jl_b200/libnufft_b200_a.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast_path or column_streaming" 2>&1 | tail -3 | tee gpurun_out/r2b_tests.log
export NUFFT_B200_LIB=$PWD/nonuniformffts.jl_b200/libnufft_b200_a.so
bash tools/gpu_ncu.sh ring_spread r2b_ring_spread 0
