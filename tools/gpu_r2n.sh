# 8 ranks: bench.py under torchrun (C5 strong scaling)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r2n_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2n_bench_8gpu.json 2> gpurun_out/r2n_bench_8gpu.err
echo EXIT $?; tail -c 1500 gpurun_out/r2n_bench_8gpu.json; grep -v "^\[W\|Warning\|^\*\|OMP_NUM" gpurun_out/r2n_bench_8gpu.err | tail -8
