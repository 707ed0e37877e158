#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1
tail -3 gpurun_out/s4_pytest.log
QB="--tiles 1184 --iters 10 --mean 2.24 --std 2.12 --signs 1 --reps 4"
V=$PWD/ldpc_toolbox_b200/_build/variants
L=gpurun_out/s4_variants.log
: > $L
run() { echo "== $1" | tee -a $L; LDPC_B200_LIB=$V/$1/libldpc_toolbox.so timeout 300 python tools/quick_bench.py $QB 2>&1 | cut -c1-150,230-560 | tee -a $L; }
for v in tma tma3 tma12; do run $v; done
echo "== main" | tee -a $L; timeout 300 python tools/quick_bench.py $QB 2>&1 | cut -c1-150,230-560 | tee -a $L
timeout 900 python tools/bench_configs.py --configs c1,c2,c4f,c4l 2>&1 | tee gpurun_out/s4_configs.log
