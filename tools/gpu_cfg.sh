# stage timings of the other BASELINE.json configurations (C4, C5) on one GPU
set -x
timeout 600 python tools/run_cfg.py --cfg c4 --np 16777216 2>&1 | tail -1 | cut -c1-500
timeout 600 python tools/run_cfg.py --cfg c4 --np 16777216 --fast 2>&1 | tail -1 | cut -c1-500
timeout 600 python tools/run_cfg.py --cfg c5 --np 134217728 2>&1 | tail -1 | cut -c1-500
