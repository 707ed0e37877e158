#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s5_pytest.log 2>&1
tail -5 gpurun_out/s5_pytest.log
timeout 900 python tools/bench_configs.py --configs c1,c2,c4f,c4l 2>&1 | tee gpurun_out/s5_configs.log
LDPC_B200_LAYERED=tile timeout 600 python tools/bench_configs.py --configs c2,c4l --points 0.5 2>&1 | tee gpurun_out/s5_configs_tile.log
