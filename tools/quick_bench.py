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
ap.add_argument("--mean", type=float, default=1.0)
ap.add_argument("--std", type=float, default=2.0)
ap.add_argument("--signs", type=int, default=0, help="1: flip the LLR sign of a random half of the positions (random-codeword-like data)")
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
    llrs = torch.randn((frames, n), generator=g, device=dev, dtype=torch.float32)
    llrs.mul_(a.std).add_(a.mean)            # in place: no 39 GB temporaries left in torch's cache
    if a.signs:
        sgn = (torch.randint(0, 2, (n,), generator=g, device=dev, dtype=torch.int32) * 2 - 1).to(torch.float32)
        llrs *= sgn
    out = torch.empty((frames, 8), dtype=torch.uint8, device=dev)
    its = torch.empty((frames,), dtype=torch.int32, device=dev)
    torch.cuda.empty_cache()
    ms = []
    for r in range(a.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dec.decode_batch_ptr(llrs.data_ptr(), False, n, frames, a.iters, out.data_ptr(), 8, 8, its.data_ptr(), device=True,
                             stream=torch.cuda.current_stream().cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        t = dec.last_timing()
        if r:
            ms.append(e0.elapsed_time(e1))      # whole call (all chunks, ingest + BP + emit)
    all_ms = ",".join(f"{m:.0f}" for m in ms)
    ms = float(np.median(ms))
    fi = frames * a.iters / (ms * 1e-3)
    prof = ""
    try:
        import ctypes
        from ldpc_toolbox_b200 import capi
        fn = capi.load().ldpc_toolbox_debug_i8_profile
        buf = (ctypes.c_uint64 * 8)()
        fn(buf)
        tot = sum(buf[:4]) or 1
        prof = " prof[init,check,stop,var]=" + ",".join(f"{buf[i] / tot:.3f}" for i in range(4)) + f" cyc/cta-iter={tot / max(buf[4], 1):.0f}"
    except AttributeError:
        pass
    print(f"{a.code} {a.impl} tiles={tiles} iters={a.iters} kernel_ms={ms:.2f} frame_iter/s={fi/1e6:.3f}M "
          f"alg_GB/s={fi*4*E/1e9:.0f} frac={fi*4*E/1e9/6553.6:.3f} conv={(its>=0).float().mean().item():.3f} "
          f"its[min,max]={its.min().item()},{its.max().item()} reps_ms=[{all_ms}] stages={t}{prof}", flush=True)
    del dec, llrs
    torch.cuda.empty_cache()
