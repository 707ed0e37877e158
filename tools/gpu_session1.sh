#!/bin/bash
# GPU session: parity tests, bench line, K1 variant probes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s1_pytest.log 2>&1
tail -3 gpurun_out/s1_pytest.log
QB="--tiles 1184 --iters 10 --mean 2.24 --std 2.12 --signs 1 --reps 3"
for v in prof vpref u4 vpref_u4 u34 vpref_u34; do
  echo "== $v" | tee -a gpurun_out/s1_variants.log
  LDPC_B200_LIB=$PWD/ldpc_toolbox_b200/_build/variants/$v/libldpc_toolbox.so timeout 300 python tools/quick_bench.py $QB 2>&1 | cut -c1-400 | tee -a gpurun_out/s1_variants.log
done
echo "== main" | tee -a gpurun_out/s1_variants.log
timeout 300 python tools/quick_bench.py $QB 2>&1 | cut -c1-400 | tee -a gpurun_out/s1_variants.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -c 3000 gpurun_out/s1_bench.json
