#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s13_pytest.log 2>&1
tail -4 gpurun_out/s13_pytest.log
A=$(python -c "from ldpc_toolbox_b200 import codes; print(codes.cached_alist_path('dvbs2:R1_2'))")
for ts in 1 0; do
echo "== LDPC_B200_TWO_STAGE=$ts"
LDPC_B200_TWO_STAGE=$ts timeout 600 python -m ldpc_toolbox_b200.cli ber $A --decoder Minstarapproxi8 --min-ebn0 1.2 --max-ebn0 1.6 --step-ebn0 0.2 --max-iter 25 \
   --frame-errors 100000 --max-frames 1000000 2>&1 | tail -5
done | tee gpurun_out/s13_two_stage.log
timeout 300 python tools/quick_bench.py --tiles 1184 --iters 10 --mean 2.24 --std 2.12 --signs 1 --reps 3 2>&1 | cut -c1-150,230-560 | tee gpurun_out/s13_quick.log
