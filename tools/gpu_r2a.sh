# round 2, first GPU pass: parity suite with the ring-window kernels + single-sweep sort, C3 stage times against the
# round-1 kernels (NUFFT_B200_RING=0 / NUFFT_B200_ONESWEEP=0), ncu captures of the two new kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2a_tests.log
for env in "NUFFT_B200_RING=1" "NUFFT_B200_RING=0" "NUFFT_B200_ONESWEEP=0"; do
  echo "== $env" | tee -a gpurun_out/r2a_c3.log
  env $env timeout 300 python tools/run_c3.py --iters 4 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/r2a_c3.log
done
echo "== clustered" | tee -a gpurun_out/r2a_c3.log
timeout 300 python tools/run_c3.py --iters 4 --dist clustered 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/r2a_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_keys|radix|scan|onesweep" -s 12 -c 12 --csv python tools/run_c3.py --iters 2 2>/dev/null | grep -E "bin_keys|radix|scan|onesweep" | awk -F'","' '{print substr($5,1,40), $NF}' | tee gpurun_out/r2a_setpoints_kernels.log
bash tools/gpu_ncu.sh ring_spread r2_ring_spread 0
bash tools/gpu_ncu.sh ring_interp r2_ring_interp 0
