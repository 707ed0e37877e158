#!/usr/bin/env python3
"""Prints, per float implementation, how many frames of the small-code stimuli of tests/test_gpu_parity_generic.py differ
from the CPU checker (the evidence behind that test's bounds)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oraclelib
import test_gpu_parity_generic as T
o = oraclelib.load()
for seed in (22, 23):
    for impl in T.FLOAT_FLOOD + T.HL_FLOAT:
        total = bad = 0
        if seed == 23:
            os.environ["LDPC_B200_LAYERED"] = "tile"
        for alist, llrs in T.small_code_stimulus(seed, np.float64 if impl.endswith("f64") else np.float32):
            nbad, its, rits = T.run_pair(o, alist, impl, llrs, 12)
            bad += nbad; total += len(its)
        print(seed, impl, bad, total, flush=True)
