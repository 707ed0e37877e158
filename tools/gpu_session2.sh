#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
QB="--tiles 1184 --iters 10 --mean 2.24 --std 2.12 --signs 1 --reps 3"
V=$PWD/ldpc_toolbox_b200/_build/variants
L=gpurun_out/s2_variants.log
: > $L
run() { echo "== $1 dephase_us=${2:-0}" | tee -a $L; LDPC_I8_DEPHASE_US=${2:-0} LDPC_B200_LIB=$V/$1/libldpc_toolbox.so timeout 300 python tools/quick_bench.py $QB 2>&1 | cut -c1-420 | tee -a $L; }
run old
run prof
for d in 0 5000 8000 11000 14000 18000 -1; do run dephase $d; done
run old
