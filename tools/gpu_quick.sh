# quick GPU round: fast-path parity subset, C3 stage times, bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fast_path or 3d_matrix or clustered or callbacks or ntransforms" 2>&1 | tail -4 > gpurun_out/quick_tests.log
cat gpurun_out/quick_tests.log
timeout 300 python tools/run_c3.py --iters 3 2>&1 | tail -2 | cut -c1-420
timeout 300 python tools/run_c3.py --iters 2 --dist clustered 2>&1 | tail -1 | cut -c1-420
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','type1_ms','type2_ms','gpu_launches')}, d['e2e']['value'], d['roofline']['stage_ms'])
PY
