#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s3_pytest.log 2>&1
tail -3 gpurun_out/s3_pytest.log
QB="--tiles 1184 --iters 10 --mean 2.24 --std 2.12 --signs 1 --reps 4"
V=$PWD/ldpc_toolbox_b200/_build/variants
L=gpurun_out/s3_variants.log
: > $L
run() { echo "== $1" | tee -a $L; LDPC_B200_LIB=$V/$1/libldpc_toolbox.so timeout 300 python tools/quick_bench.py $QB 2>&1 | cut -c1-150,230-560 | tee -a $L; }
for r in 1 2; do for v in oldfold new tma; do run $v; done; done
run tmaprof
