#!/usr/bin/env python3
"""Small-batch latency of the int8 flooding decoder through the reference-shaped calls:
decode() of ONE DVB-S2 normal frame (f64 LLRs, host buffers) and decode_batch of a few thousand
short frames, with the default cluster size and with LDPC_B200_CLUSTER=1 (one CTA per tile)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402
from ldpc_toolbox_b200 import Decoder, Encoder, codes  # noqa: E402

rng = np.random.default_rng(3)
for spec, n, k, ebn0, nframes in (("dvbs2:R1_2", 64800, 32400, 1.3, 1), ("dvbs2:R1_2", 64800, 32400, 1.3, 128),
                                  ("dvbs2:R1_2short", 16200, 7200, 1.4, 4096)):
    alist = codes.cached_alist_path(spec)
    enc = Encoder(alist)
    cws = np.stack([enc.encode(rng.integers(0, 2, k, dtype=np.uint8), n) for _ in range(min(nframes, 16))])
    llrs = helpers.awgn_llrs(rng, cws[np.arange(nframes) % len(cws)], helpers.sigma_for(ebn0, k / n), np.float64)
    for cl in ("", "1"):
        if cl:
            os.environ["LDPC_B200_CLUSTER"] = cl
        else:
            os.environ.pop("LDPC_B200_CLUSTER", None)
        dec = Decoder(alist, "Minstarapproxi8", device=0)
        ts = []
        for r in range(6):
            t0 = time.perf_counter()
            if nframes == 1:
                out, it = dec.decode(llrs[0], 25)
                its = np.array([it])
            else:
                out, its = dec.decode_batch(llrs, 25, output_len=k)
            ts.append(time.perf_counter() - t0)
        print(f"{spec} frames={nframes} cluster={'auto' if not cl else cl}: best {min(ts[1:]) * 1e3:.2f} ms per call, "
              f"iterations mean {np.where(its < 0, 25, its).mean():.1f}, converged {(its >= 0).mean():.2f}", flush=True)
        dec.close()
