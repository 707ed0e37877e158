#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/s11_pytest.log 2>&1
tail -5 gpurun_out/s11_pytest.log
timeout 600 python tools/latency_probe.py 2>&1 | tee gpurun_out/s11_latency.log
timeout 300 python tools/quick_bench.py --tiles 1184 --iters 10 --mean 2.24 --std 2.12 --signs 1 --reps 3 2>&1 | cut -c1-150,230-560 | tee gpurun_out/s11_quick.log
