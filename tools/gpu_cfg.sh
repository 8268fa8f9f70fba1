set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/fullsize_tests.log
cat gpurun_out/fullsize_tests.log
timeout 600 python tools/run_cfg.py --cfg c4 --np 16777216 2>&1 | tail -2 | cut -c1-500
timeout 600 python tools/run_cfg.py --cfg c5 --np 16777216 2>&1 | tail -1 | cut -c1-500
timeout 600 python tools/run_cfg.py --cfg c5 --np 134217728 2>&1 | tail -1 | cut -c1-500
