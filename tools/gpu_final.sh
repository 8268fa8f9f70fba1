# final round: full parity suite, smoke, bench line, launch list, ncu captures of the two dominant kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/run_c3.py --iters 3 2>&1 | tail -1 | cut -c1-420
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','type1_ms','type2_ms','gpu_launches')}, d['e2e']['value'], d['cpu_baseline'] and d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
bash tools/gpu_ncu.sh cs_spread cs_spread 0
bash tools/gpu_ncu.sh cs_interp cs_interp 0
