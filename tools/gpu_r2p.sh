# 8 ranks: bench.py under torchrun (C5 strong scaling, peer-window exchanges), then the exchange-mode test on two of the GPUs
set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2p_bench_8gpu.json 2> gpurun_out/r2p_bench_8gpu.err
echo EXIT $?; tail -c 3000 gpurun_out/r2p_bench_8gpu.json; grep -v "^\[W\|Warning\|^\*\|OMP_NUM" gpurun_out/r2p_bench_8gpu.err | tail -8
timeout 200 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "exchange_modes or against_single" > gpurun_out/r2p_tests.log 2>&1; echo TESTS $?; tail -5 gpurun_out/r2p_tests.log
