#!/usr/bin/env python3
"""Generates tests/golden/golden_small.npz: seeded LLR inputs and the decoded words / iteration
counts produced by the CPU checker (oracle/) for all 36 implementation names on two small codes.

Provenance: the reference (Rust) cannot be executed in this environment, so these vectors are the
OUTPUT OF THE C++ RESTATEMENT, not of the reference itself.  They pin the checker against
accidental change (compiler flags, refactors) and give the GPU tests a fixture that does not
depend on building the checker.  The reference's own known-answer tests are embedded separately in
tests/test_oracle_kat.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import helpers  # noqa: E402
import oraclelib  # noqa: E402

JOHNSON = "6 4\n2 3\n2 2 2 2 2 2\n3 3 3 3\n1 3\n1 2\n2 4\n1 4\n2 3\n3 4\n1 2 4\n2 3 5\n1 5 6\n3 4 6\n"


def main():
    o = oraclelib.load()
    rng = np.random.default_rng(20261017)
    random_alist = helpers.random_code_alist(rng, 96, 48, col_w=[1, 2, 3, 4, 9], extra_heavy_rows=1)
    out = {"alist_johnson": np.array(JOHNSON), "alist_random": np.array(random_alist), "max_iter": np.array(12)}
    for tag, alist, n in (("johnson", JOHNSON, 6), ("random", random_alist, 96)):
        cw = np.zeros((48, n), dtype=np.uint8)
        if tag == "johnson":
            cw[:] = np.array([0, 0, 1, 0, 1, 1], dtype=np.uint8)
        llrs = np.concatenate([helpers.awgn_llrs(rng, cw[:16], s) for s in (0.4, 0.8, 1.2)])
        llrs[0] = 0.0
        out[f"llrs_{tag}"] = llrs
        for impl in o.implementations():
            dec = o.decoder(alist, impl)
            bits, its = dec.decode_batch(llrs, 12, nthreads=1)
            out[f"bits_{tag}_{impl}"] = np.packbits(bits, axis=1)
            out[f"its_{tag}_{impl}"] = its.astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)
    print("wrote golden_small.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
