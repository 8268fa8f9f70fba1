# 8 ranks: bench.py under torchrun (C5 strong scaling, peer-window exchanges)
set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo EXIT $?; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_8gpu.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "type1_ms", "type2_ms", "check", "speedup_vs_single_gpu")})
print(d["stage_ms_rank0"]); print(d["e2e"]["value"], d["weak_c3"]["value"], d["config"]["exchange"])
PY
grep -v "^\[W\|Warning\|^\*\|OMP_NUM" gpurun_out/r2_bench_8gpu.err | tail -8
