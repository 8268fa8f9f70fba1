set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/full_tests.log
cat gpurun_out/full_tests.log
for np in 1048576 4194304 8388608 16777216; do
  echo "np $np"; timeout 300 python tools/run_c3.py --iters 2 --np $np 2>&1 | tail -1 | cut -c1-400
done
