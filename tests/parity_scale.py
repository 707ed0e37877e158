#!/usr/bin/env python3
"""Frame-level parity at scale for the float decoders (BASELINE.json: decoded words must match the
reference's on >= 99.99 % of frames): decodes the same AWGN frames with the GPU path (C-ABI) and with
the CPU checker (oracle/, all host threads) and counts frames whose hard word or iteration count differ.
Test infrastructure (lives under tests/ because it loads the oracle): the oracle is only the checker here.

  python tests/parity_scale.py --code nr5g:2:384 --impl HLMinstarapproxf32 --frames 8192 --ebn0 0.25 --max-iter 50
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402
import oraclelib  # noqa: E402
from ldpc_toolbox_b200 import Decoder, codes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--code", default="nr5g:2:384")
ap.add_argument("--impl", default="HLMinstarapproxf32")
ap.add_argument("--frames", type=int, default=8192)
ap.add_argument("--ebn0", type=float, default=0.25)
ap.add_argument("--max-iter", type=int, default=50)
ap.add_argument("--cpu-only", action="store_true", help="time the checker only (no GPU)")
a = ap.parse_args()

o = oraclelib.load()
alist = codes.alist_for(a.code)
first = alist.split("\n", 1)[0].split()
n, m = int(first[0]), int(first[1])
k = n - m
rng = np.random.default_rng(2026)
enc = o.encoder(alist)
msgs, cws = helpers.encoded_frames(enc, rng, k, n, 32)
llrs = helpers.awgn_llrs(rng, cws[np.arange(a.frames) % 32], helpers.sigma_for(a.ebn0, k / n),
                         np.float64 if a.impl.endswith("f64") else np.float32)
t0 = time.perf_counter()
rout, rits = o.decoder(alist, a.impl).decode_batch(llrs, a.max_iter, out_len=k, nthreads=os.cpu_count())
t_cpu = time.perf_counter() - t0
res = {"code": a.code, "impl": a.impl, "frames": a.frames, "ebn0_db": a.ebn0, "max_iter": a.max_iter,
       "cpu_seconds": round(t_cpu, 2), "cpu_threads": os.cpu_count(),
       "avg_iterations": float(np.where(rits < 0, a.max_iter, rits).mean()), "cpu_failures": int((rits < 0).sum())}
if not a.cpu_only:
    dec = Decoder(alist, a.impl, device=0)
    out, its = dec.decode_batch(llrs, a.max_iter, output_len=k)
    t0 = time.perf_counter()
    out, its = dec.decode_batch(llrs, a.max_iter, output_len=k)
    res["gpu_seconds_host_buffers"] = round(time.perf_counter() - t0, 3)
    word_diff = (out != rout).any(axis=1)
    res["frames_word_differs"] = int(word_diff.sum())
    res["frames_iterations_differ"] = int((its != rits).sum())
    res["frame_match_rate"] = float(1.0 - (word_diff | (its != rits)).mean())
    res["word_match_rate"] = float(1.0 - word_diff.mean())
print(json.dumps(res), flush=True)
