#!/usr/bin/env python3
"""Diagnostic (test infrastructure): posterior-LLR agreement between the GPU decoders and the CPU checker, per
implementation: quantiles of the relative error on frames whose word and iteration count match."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402
import oraclelib  # noqa: E402
from ldpc_toolbox_b200 import Decoder, codes  # noqa: E402

o = oraclelib.load()
CASES = [("ar4ja:1/2:1024", "1,1,1,1,0", 1.8, 50, ["Phif64", "Phif32", "Tanhf64", "Tanhf32", "Minstarapproxf64", "Minstarapproxf32", "Aminstarf64", "Aminstarf32"]),
         ("nr5g:2:96", "", 1.2, 30, ["HLPhif64", "HLPhif32", "HLTanhf64", "HLTanhf32", "HLMinstarapproxf64", "HLMinstarapproxf32", "HLAminstarf64", "HLAminstarf32"])]
for code, punct, ebn0, max_iter, impls in CASES:
    alist = codes.alist_for(code)
    n, m = (int(x) for x in alist.split("\n")[0].split())
    k = n - m
    rng = np.random.default_rng(5)
    enc = o.encoder(alist, punct)
    n_tx = n * 4 // 5 if punct else n
    msgs = rng.integers(0, 2, size=(8, k), dtype=np.uint8)
    tx = np.stack([enc.encode(mm, n_tx) for mm in msgs])
    for impl in impls:
        dtype = np.float64 if impl.endswith("f64") else np.float32
        llrs = helpers.awgn_llrs(rng, tx[np.arange(256) % 8], helpers.sigma_for(ebn0, k / n_tx), dtype)
        out, its, post = Decoder(alist, impl, punct).decode_batch_posteriors(llrs, max_iter, output_len=k)
        ref = o.decoder(alist, impl, punct)
        rel, used = [], 0
        for f in range(llrs.shape[0]):
            rout, rit = ref.decode(llrs[f], max_iter, out_len=k)
            if rit != its[f] or (rout != out[f]).any() or rit == 0:
                continue
            rp = ref.posteriors()
            rel.append(np.abs(post[f] - rp) / np.maximum(np.abs(rp), 1.0))
            used += 1
        rel = np.concatenate(rel) if rel else np.zeros(1)
        print(f"{code} {impl}: frames compared {used}/256, rel err median {np.median(rel):.3g} p99 {np.quantile(rel, 0.99):.3g} "
              f"p99.99 {np.quantile(rel, 0.9999):.3g} max {rel.max():.3g}", flush=True)
