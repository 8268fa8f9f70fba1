# column-streaming work-item size (NUFFT_B200_CS_CHUNK): stage times and DRAM traffic per launch of cs_spread / cs_interp at C3
set -x
mkdir -p gpurun_out
for c in 64 128 256 512 1024; do
  echo "chunk $c" | tee -a gpurun_out/chunk_sweep.log
  NUFFT_B200_CS_CHUNK=$c timeout 300 python tools/run_c3.py --iters 3 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/chunk_sweep.log
  NUFFT_B200_CS_CHUNK=$c timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:cs_ -s 2 -c 2 --csv python tools/run_c3.py --iters 2 2>/dev/null | grep -E "cs_(spread|interp)" | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-200 | tee -a gpurun_out/chunk_sweep.log
done
