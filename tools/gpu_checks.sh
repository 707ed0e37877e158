#!/bin/bash
# One GPU session: parity tests, the other BASELINE configurations, the bench line, and the ncu evidence kept under
# profiles/ (launch list of the bench command + one --set full capture of K1, K2, K3q and the BER kernels; summaries are
# produced on the box because gpurun merges at most 64 MiB back).
# Run on a B200 through gpurun:  gpurun --timeout 3000 -- bash tools/gpu_checks.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${ROUND_TAG:-r02}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/chk_pytest.log 2>&1
tail -5 gpurun_out/chk_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python tools/bench_configs.py --configs c2,c4f,c4l 2>&1 | tee gpurun_out/chk_configs.jsonl
timeout 1500 python tests/config5_dvbs2_all.py --frames 18944 --out gpurun_out/chk_config5.jsonl > gpurun_out/chk_config5.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/chk_bench.json 2> gpurun_out/chk_bench.err
tail -c 2500 gpurun_out/chk_bench.json
[ -n "$SKIP_NCU" ] && exit 0
# launch list of the bench command (per-launch durations; shares of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/chk_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${R}_launches_raw.csv gpurun_out/${R}_launches_summary.csv | head -8
cap() {   # name, kernel regex, command...
    local name=$1 rx=$2; shift 2
    timeout 900 ncu --set full --clock-control none -k regex:$rx -s 1 -c 1 -f -o gpurun_out/${R}_$name "$@" > gpurun_out/chk_ncu_$name.log 2>&1
    python tools/ncu_summary.py gpurun_out/${R}_$name.ncu-rep gpurun_out/${R}_${name}_ncu_full.json > /dev/null
    rm -f gpurun_out/${R}_$name.ncu-rep
}
cap flood_i8 flood_i8 python tools/quick_bench.py --tiles 1184 --iters 25 --mean 2.24 --std 2.12 --signs 1 --reps 1
cap flood_float_c4 flood_float python tools/bench_configs.py --configs c4f --points=0.0 --reps 1
cap layered_smem_c2 layered_smem python tools/bench_configs.py --configs c2 --points=-0.5 --reps 1
ls -la gpurun_out | tail -12
