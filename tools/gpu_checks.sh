#!/bin/bash
# One GPU session: parity tests, the other BASELINE configurations, the bench line, and the ncu evidence
# kept under profiles/ (launch list of the bench command + one --set full capture of K1 and of K3q).
# Run on a B200 through gpurun:  gpurun --timeout 3000 -- bash tools/gpu_checks.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/chk_pytest.log 2>&1
tail -5 gpurun_out/chk_pytest.log
timeout 900 python tools/bench_configs.py --configs c2,c4f,c4l 2>&1 | tee gpurun_out/chk_configs.log
timeout 600 python tools/bench_configs.py --configs c2,c4f,c4l --frames 32768 --points 0.5 2>&1 | tee -a gpurun_out/chk_configs.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/chk_bench.json 2> gpurun_out/chk_bench.err
tail -c 2500 gpurun_out/chk_bench.json
# launch list of the bench command (per-launch durations; shares of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_e_launches_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/chk_bench_under_ncu.log 2>&1
[ -n "$SKIP_FULL_NCU" ] && exit 0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layered_smem -s 1 -c 1 -f -o gpurun_out/r01_e_layered_smem \
    python tools/bench_configs.py --configs c2 --points=-0.5 --reps 1 > gpurun_out/chk_ncu_k3q.log 2>&1
[ -n "$SKIP_K1_FULL_NCU" ] && exit 0
# full capture of the dominant kernel (one launch, 25 iterations, bench-sized batch; ~7 minutes of replays)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flood_i8 -s 1 -c 1 -f -o gpurun_out/r01_e_flood_i8 \
    python tools/quick_bench.py --tiles 1184 --iters 25 --mean 2.24 --std 2.12 --signs 1 --reps 1 > gpurun_out/chk_ncu_k1.log 2>&1
ls -la gpurun_out/*.ncu-rep
