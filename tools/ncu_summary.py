#!/usr/bin/env python3
"""Condenses an .ncu-rep into the handful of metrics quoted in profiles/ and DESIGN.md.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.json]"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
out = []
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'sm__cycles_elapsed.avg',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'smsp__cycles_active.avg']
for vals in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
            d[h] = f"{v} {u}".strip()
    d['Kernel Name'] = vals[hdr.index('Kernel Name')]
    out.append(d)
txt = json.dumps(out if len(out) > 1 else out[0], indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(txt + "\n")
print(txt)
