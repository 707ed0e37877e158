#!/usr/bin/env python3
"""Experiment builds of K1: compiles flood_i8.cu with extra -D flags (only the north-star
instantiation) and links it with the regular objects into
ldpc_toolbox_b200/_build/variants/<name>/libldpc_toolbox.so.  Use with LDPC_B200_LIB=<that path>.

  python tools/build_variant.py name -DLDPC_I8_U3=4 ...
"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from ldpc_toolbox_b200 import build as B  # noqa: E402

name, defines = sys.argv[1], sys.argv[2:]
full = "--full" in defines            # all 8 instantiations (parity tests), not only the north-star one
src = next((d[6:] for d in defines if d.startswith("--src=")), "flood_i8.cu")      # another revision of the kernel source, in csrc/
defines = [d for d in defines if d != "--full" and not d.startswith("--src=")]
B.build()
out = os.path.join(B.OUT_DIR, "variants", name)
os.makedirs(out, exist_ok=True)
ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
obj = os.path.join(out, "flood_i8.o")
cmd = [B._nvcc()] + ccbin + B.NVCC_FLAGS + ([] if full else ["-DLDPC_I8_BENCH_ONLY"]) + defines + ["-x", "cu", "-c", os.path.join(B.CSRC, src), "-o", obj]
r = subprocess.run(cmd, capture_output=True, text=True)
open(os.path.join(out, "build.log"), "w").write(r.stdout + r.stderr)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
for l in (r.stdout + r.stderr).splitlines():
    if "registers" in l or "spill" in l:
        print(name, l.strip())
objs = [obj] + [os.path.join(B.OUT_DIR, s.rsplit(".", 1)[0] + ".o") for s in B.CU_SOURCES + B.CPP_SOURCES if s != "flood_i8.cu"]
lib = os.path.join(out, "libldpc_toolbox.so")
subprocess.check_call([B._nvcc()] + ccbin + ["-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
print(lib)
