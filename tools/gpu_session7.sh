#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s7_pytest.log 2>&1
tail -5 gpurun_out/s7_pytest.log
timeout 900 python tools/bench_configs.py --configs c2,c4l 2>&1 | tee gpurun_out/s7_configs.log
timeout 600 python tools/bench_configs.py --configs c2,c4f,c4l --frames 32768 --points 0.5 2>&1 | tee -a gpurun_out/s7_configs.log
timeout 600 python tools/bench_configs.py --configs c3w --reps 1 2>&1 | tee -a gpurun_out/s7_configs.log
