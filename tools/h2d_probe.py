#!/usr/bin/env python3
"""Host->device ceiling of the box for the end-to-end path: every rank (one per GPU, torchrun) copies a pinned host
buffer the size of one half-launch of f32 LLRs (75 776 frames x 64 800 x 4 B = 19.6 GB, or --gb) to its GPU over and
over, all ranks at once, optionally with the D2H stream of the decoded words running beside it.  Prints one JSON line:
per-rank and aggregate GB/s.  bench.py's e2e needs 259.2 KB in + 32.4 KB out per frame, i.e. 42 GB/s per GPU at 5.3 Gbit/s.

  python tools/h2d_probe.py                                  # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--gb", type=float, default=19.6)
ap.add_argument("--seconds", type=float, default=6.0)
ap.add_argument("--d2h", action="store_true", help="run a D2H stream of 1/8 of the bytes beside the H2D stream")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

n = int(a.gb * 1e9) // 4
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
h.fill_(1.0)                                   # touch every page
d = torch.empty(n, dtype=torch.float32, device=dev)
h2 = torch.empty(n // 8, dtype=torch.float32, pin_memory=True)
d2 = torch.zeros(n // 8, dtype=torch.float32, device=dev)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
with torch.cuda.stream(s_in):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize(dev)
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
copies = 0
while time.perf_counter() - t0 < a.seconds:
    with torch.cuda.stream(s_in):
        d.copy_(h, non_blocking=True)
    if a.d2h:
        with torch.cuda.stream(s_out):
            h2.copy_(d2, non_blocking=True)
    s_in.synchronize()
    copies += 1
torch.cuda.synchronize(dev)
dt = time.perf_counter() - t0
gbs = copies * n * 4 / dt / 1e9
t = torch.tensor([gbs], device=dev, dtype=torch.float64)
if world > 1:
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    per = [float(x.item()) for x in g]
else:
    per = [gbs]
if rank == 0:
    print(json.dumps({"gpus": world, "buffer_gb": a.gb, "with_d2h": a.d2h, "h2d_gbs_per_rank": [round(x, 2) for x in per],
                      "h2d_gbs_aggregate": round(sum(per), 1), "h2d_gbs_min": round(min(per), 2),
                      "needed_gbs_per_gpu_at_5.3_gbit_s": 42.4,
                      "host_cores": os.cpu_count()}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
