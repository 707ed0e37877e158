#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:flood_i8 -s 1 -c 1 -f -o gpurun_out/r01_e_flood_i8 \
    python tools/quick_bench.py --tiles 1184 --iters 25 --mean 2.24 --std 2.12 --signs 1 --reps 1 > gpurun_out/chk_ncu_k1.log 2>&1
tail -2 gpurun_out/chk_ncu_k1.log | cut -c1-300
ls -la gpurun_out/r01_e_flood_i8.ncu-rep
