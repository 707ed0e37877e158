#!/usr/bin/env python3
"""Quick kernel-time probe for experiments (not the judged benchmark): decode `tiles`*128 synthetic
frames for `iters` iterations and print the BP kernel's device time and frame-iterations/s."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ldpc_toolbox_b200 import Decoder, codes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--code", default="dvbs2:R1_2")
ap.add_argument("--impl", default="Minstarapproxi8")
ap.add_argument("--tiles", default="148,296,592")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--nw", type=int, default=0)
a = ap.parse_args()

if a.nw:
    os.environ["LDPC_B200_NW"] = str(a.nw)
alist = codes.cached_alist_path(a.code)
dev = torch.device("cuda", 0)
for tiles in [int(t) for t in a.tiles.split(",")]:
    dec = Decoder(alist, a.impl, device=0, max_tiles=tiles)
    n, E = dec.n, dec.num_edges
    frames = tiles * 128
    g = torch.Generator(device=dev); g.manual_seed(1)
    llrs = torch.randn((frames, n), generator=g, device=dev, dtype=torch.float32) * 2.0 + 1.0   # never converges
    out = torch.empty((frames, 8), dtype=torch.uint8, device=dev)
    its = torch.empty((frames,), dtype=torch.int32, device=dev)
    ms = []
    for r in range(a.reps + 1):
        dec.decode_batch_ptr(llrs.data_ptr(), False, n, frames, a.iters, out.data_ptr(), 8, 8, its.data_ptr(), device=True,
                             stream=torch.cuda.current_stream().cuda_stream)
        t = dec.last_timing()
        if r:
            ms.append(t["decode_ms"])
    ms = float(np.median(ms))
    fi = frames * a.iters / (ms * 1e-3)
    print(f"{a.code} {a.impl} tiles={tiles} iters={a.iters} kernel_ms={ms:.2f} frame_iter/s={fi/1e6:.3f}M "
          f"alg_GB/s={fi*4*E/1e9:.0f} frac={fi*4*E/1e9/6553.6:.3f} conv={(its>=0).float().mean().item():.3f}", flush=True)
    del dec, llrs
    torch.cuda.empty_cache()
