set -x
for np in 262144 1048576 4194304 16777216 67108864; do
  for cs in 1 0; do
    echo "np $np cs $cs"; NUFFT_B200_CS=$cs timeout 300 python tools/run_c3.py --iters 2 --np $np 2>&1 | tail -1 | cut -c1-400
  done
done
