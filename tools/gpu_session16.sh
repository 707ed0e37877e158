#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in 256 192 128; do
  echo "== K3Q_THREADS=$t"
  LDPC_B200_K3Q_THREADS=$t timeout 300 python tools/bench_configs.py --configs c2,c4l --points=-0.5,0.5 2>&1 | cut -c1-200
  LDPC_B200_K3Q_THREADS=$t timeout 300 python tools/bench_configs.py --configs c2 --frames 32768 --points 0.5 2>&1 | cut -c1-200
done | tee gpurun_out/s18_threads.log
