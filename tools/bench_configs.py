#!/usr/bin/env python3
"""Throughput of the other BASELINE.json configurations through the on-device BER engine
(random message -> encode -> BPSK/AWGN -> decode -> error counters, everything on the GPU).
Not the judged benchmark (bench.py is); prints one JSON line per (config, Eb/N0) point.

  python tools/bench_configs.py [--configs c1,c2,c4f,c4l] [--frames N] [--reps R]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ldpc_toolbox_b200 import codes  # noqa: E402
from ldpc_toolbox_b200.ber import COUNTER_NAMES, BerEngine  # noqa: E402

CONFIGS = {
    # name: (code spec, implementation, puncturing, max_iterations, frames, Eb/N0 points)
    "c1": ("ar4ja:1/2:1024", "Phif64", "1,1,1,1,0", 100, 16384, [1.0, 1.5, 2.0]),
    "c2": ("nr5g:2:384", "HLMinstarapproxf32", "", 50, 4096, [-0.5, 0.0, 0.5, 1.0]),
    "c4f": ("nr5g:1:384", "Aminstarf32", "", 50, 4096, [0.0, 0.5, 1.0, 1.5]),
    "c4l": ("nr5g:1:384", "HLAminstarf32", "", 50, 4096, [0.0, 0.5, 1.0, 1.5]),
    "c3w": ("dvbs2:R1_2", "Minstarapproxi8", "", 25, 151552, [0.9, 1.0, 1.1]),
}

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="c1,c2,c4f,c4l")
ap.add_argument("--frames", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--points", default="")
a = ap.parse_args()

for name in a.configs.split(","):
    spec, impl, punct, max_iter, frames, points = CONFIGS[name]
    if a.frames:
        frames = a.frames
    if a.points:
        points = [float(x) for x in a.points.split(",")]
    path = codes.cached_alist_path(spec)
    head = open(path).read().split("\n", 4)
    n_cw = int(head[0].split()[0])
    edges = sum(int(x) for x in head[2].split())                     # column weights
    layered = impl.startswith("HL")
    s_msg = 8 if impl.endswith("f64") else (4 if impl.endswith("f32") else 1)
    # 4 E s bytes per frame-iteration (flooding), 2 E s (layered), + f32 LLRs in + packed bits out
    bytes_per_frame_it = (2 if layered else 4) * edges * s_msg
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    eng = BerEngine(path, impl, punct, device=0)
    for ebn0 in points:
        eng.run(ebn0, max_iter, 0, frames)            # warm-up (allocations, first launch)
        best, c = None, None
        for r in range(a.reps):
            t0 = time.perf_counter()
            c = eng.run(ebn0, max_iter, (r + 1) * frames, frames)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        d = dict(zip(COUNTER_NAMES, [int(x) for x in c]))
        print(json.dumps({"config": name, "code": spec, "impl": impl, "max_iter": max_iter, "ebn0_db": ebn0, "frames": frames,
                          "seconds": round(best, 4), "info_gbps": round(eng.k * frames / best / 1e9, 5),
                          "frames_per_s": round(frames / best, 1), "avg_iterations": round(d["total_iterations"] / frames, 3),
                          "fer": d["frame_errors"] / frames, "ber": d["bit_errors"] / (frames * eng.k),
                          "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame_iteration": bytes_per_frame_it,
                                       "achieved_gbs": round((d["total_iterations"] * bytes_per_frame_it + frames * (n_cw * 4 + n_cw / 8)) / best / 1e9, 1),
                                       "peak_gbs": peak,
                                       "frac": round((d["total_iterations"] * bytes_per_frame_it + frames * (n_cw * 4 + n_cw / 8)) / best / 1e9 / peak, 4),
                                       "note": "whole on-device BER pipeline (front-end + decode + back-end) in the denominator"}}), flush=True)
    eng.close()
