#!/usr/bin/env python3
"""BASELINE.json configs[4]: ALL 21 DVB-S2 codes (11 normal + 10 short FECFRAMEs, reference
src/codes/dvbs2.rs:133-189) through the int8 flooding decoder.

For every code:
  1. parity — a small sample of AWGN frames around the code's waterfall is decoded by the GPU (C-ABI) and by the CPU
     checker (oracle/); words AND iteration counts must be identical.  This is what exercises the generic-degree
     check path (row degrees 10...30 of the rates >= 3/5) and the staircase fusion on every real matrix.
  2. waterfall — a short BER sweep through the on-device BER engine (Eb/N0 around the standard's threshold), one
     JSON line per point with FER / BER / average iterations / info Gbit/s.

Test infrastructure (it loads the oracle, hence lives under tests/).  Run on a GPU box:
  python tests/config5_dvbs2_all.py [--codes R1_2,R9_10] [--parity-frames 96] [--frames 37888] [--max-iter 50] [--gpus 1]
"""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402
import oraclelib  # noqa: E402
from ldpc_toolbox_b200 import Decoder, codes  # noqa: E402
from ldpc_toolbox_b200.ber import COUNTER_NAMES, BerEngine, BerTest  # noqa: E402

# Eb/N0 (dB, BPSK) near which the code's waterfall sits for ~50 iterations: the standard's QPSK Es/N0 at quasi
# error-free (ETSI EN 302 307-1 table 13) minus 10 log10(2 R); short frames lose a few tenths of a dB.
QEF_ESN0_QPSK = {"1_4": -2.35, "1_3": -1.24, "2_5": -0.30, "1_2": 1.00, "3_5": 2.23, "2_3": 3.10, "3_4": 4.03, "4_5": 4.68,
                 "5_6": 5.18, "8_9": 6.20, "9_10": 6.42}


def waterfall_ebn0(name: str, k: int, n: int) -> float:
    key = name[1:].replace("short", "")
    return QEF_ESN0_QPSK[key] - 10 * np.log10(2 * k / n) + (0.35 if name.endswith("short") else 0.0)


ap = argparse.ArgumentParser()
ap.add_argument("--codes", default="")
ap.add_argument("--impl", default="Minstarapproxi8")
ap.add_argument("--max-iter", type=int, default=50)
ap.add_argument("--parity-frames", type=int, default=96)
ap.add_argument("--frames", type=int, default=37888, help="frames per BER point (0: skip the sweep)")
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--out", default="")
a = ap.parse_args()

names = [c for c in a.codes.split(",") if c] or codes.dvbs2_names()
o = oraclelib.load()
out = open(a.out, "a") if a.out else None


def emit(d):
    line = json.dumps(d)
    print(line, flush=True)
    if out:
        out.write(line + "\n")
        out.flush()


for name in names:
    spec = "dvbs2:" + name
    alist = codes.alist_for(spec)
    first = alist.split("\n", 2)
    n, m = (int(x) for x in first[0].split())
    k = n - m
    row_w = np.array([int(x) for x in alist.split("\n")[3].split()])
    eb0 = waterfall_ebn0(name, k, n)
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    # ---- 1. parity sample: three operating points around the waterfall
    enc = o.encoder(alist)
    msgs, cws = helpers.encoded_frames(enc, rng, k, n, 8)
    per = a.parity_frames // 3
    llrs = np.concatenate([helpers.awgn_llrs(rng, cws[np.arange(per) % 8], helpers.sigma_for(e, k / n)) for e in (eb0 - 0.35, eb0, eb0 + 0.4)])
    t0 = time.perf_counter()
    rout, rits = o.decoder(alist, a.impl).decode_batch(llrs, a.max_iter, out_len=k, nthreads=os.cpu_count())
    t_cpu = time.perf_counter() - t0
    dec = Decoder(alist, a.impl, device=0)
    gout, gits = dec.decode_batch(llrs, a.max_iter, output_len=k)
    dec.close()
    rec = {"code": name, "n": n, "k": k, "row_degree_max": int(row_w.max()), "impl": a.impl, "max_iter": a.max_iter,
           "parity_frames": int(llrs.shape[0]), "word_mismatches": int((gout != rout).any(axis=1).sum()),
           "iteration_mismatches": int((gits != rits).sum()), "converged": int((rits >= 0).sum()),
           "cpu_frames_per_s": round(llrs.shape[0] / t_cpu, 1), "ebn0_center_db": round(float(eb0), 2)}
    emit(rec)
    if rec["word_mismatches"] or rec["iteration_mismatches"]:
        sys.exit(f"parity FAILED on {name}")
    if a.frames <= 0:
        continue
    # ---- 2. waterfall through the BER engine
    path = codes.cached_alist_path(spec)
    engines = [BerEngine(path, a.impl, device=g) for g in range(a.gpus)]
    batch = max(512, a.frames // a.gpus // 2)
    for e in (eb0 - 0.2, eb0, eb0 + 0.2, eb0 + 0.4):
        e = round(float(e), 2)
        t = BerTest(engines, engines[0].k, [e], max_iterations=a.max_iter, max_frame_errors=10**9, batch=batch, max_frames=a.frames)
        t0 = time.perf_counter()
        st = t.run()[0]
        dt = time.perf_counter() - t0
        emit({"code": name, "ebn0_db": e, "frames": st.num_frames, "fer": st.ldpc.fer, "ber": st.ldpc.ber,
              "avg_iterations": round(st.average_iterations, 2), "false_decodes": st.false_decodes,
              "info_gbps": round(engines[0].k * st.num_frames / dt / 1e9, 4), "gpus": a.gpus, "seconds": round(dt, 3)})
    for eng in engines:
        eng.close()
