"""Builds the CUDA extension in-tree: ldpc_toolbox_b200/_build/libldpc_toolbox.so (sm_100a only).

nvcc cross-compiles without a GPU; the resulting .so travels with the repo snapshot to the GPU
box.  Nothing here falls back to a CPU implementation: if the build or the load fails, importing
the C-ABI raises.
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libldpc_toolbox.so")
STATIC_LIB = os.path.join(OUT_DIR, "libldpc_toolbox.a")     # the reference builds cdylib + staticlib (Cargo.toml:17-19)

# heavy kernels first: one nvcc per translation unit runs in parallel (see build())
CU_SOURCES = ["flood_float_f64.cu", "flood_float_f32.cu", "layered_smem_f64.cu", "layered_smem_f32.cu", "layered_tile_f64.cu",
              "layered_tile_f32.cu", "flood_i8.cu", "flood_i8_w16.cu", "flood_i8_w32.cu", "layered_tile_i8.cu", "layered_smem_i8.cu", "ber.cu", "capi.cu", "decoder.cu",
              "generic_bp.cu", "ingest.cu", "layered_smem.cu"]
CPP_SOURCES = ["host.cpp"]

# experiment knobs for flood_i8.cu (warps per CTA, min CTAs per SM); empty = the defaults in the source
EXTRA_DEFINES = [d for d in os.environ.get("LDPC_B200_DEFINES", "").split() if d]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + EXTRA_DEFINES


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found")
    return cand


def _sources():
    names = CU_SOURCES + CPP_SOURCES
    extra = [f for f in sorted(os.listdir(CSRC)) if f.endswith((".hpp", ".cuh", ".h"))]
    return names, extra


def _digest() -> str:
    h = hashlib.sha256()
    names, extra = _sources()
    for f in names + extra:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "ldpc_toolbox.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "stamp.txt")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STATIC_LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    log = []
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); nvcc must use the system host compiler
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    names, headers = _sources()
    hh = hashlib.sha256()
    for f in headers:
        hh.update(open(os.path.join(CSRC, f), "rb").read())
    hh.update(open(os.path.join(HERE, "..", "include", "ldpc_toolbox.h"), "rb").read())
    hh.update(" ".join(NVCC_FLAGS).encode())
    header_digest = hh.hexdigest()

    def compile_one(src):
        # per translation unit: recompiled only when its own source, a header or the flags changed
        obj = os.path.join(OUT_DIR, src.rsplit(".", 1)[0] + ".o")
        body = open(os.path.join(CSRC, src), "rb").read()
        for inc in re.findall(rb'#include "([^"]+\.cu)"', body):            # a unit that includes another .cu depends on it
            body += open(os.path.join(CSRC, inc.decode()), "rb").read()
        tu = hashlib.sha256(body + header_digest.encode()).hexdigest()
        tu_stamp = obj + ".stamp"
        if not force and os.path.exists(obj) and os.path.exists(tu_stamp) and open(tu_stamp).read() == tu:
            return obj, f"# {src}: up to date\n", 0
        cmd = [nvcc] + ccbin + NVCC_FLAGS + ["-x", "cu" if src.endswith(".cu") else "c++", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            open(tu_stamp, "w").write(tu)
        return obj, "$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr, r.returncode

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:      # one nvcc per translation unit
        results = list(ex.map(compile_one, CU_SOURCES + CPP_SOURCES))
    for (obj, text, rc), src in zip(results, CU_SOURCES + CPP_SOURCES):
        log.append(text)
        if rc != 0:
            sys.stderr.write(text)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [nvcc] + ccbin + ["-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    # static archive of the same objects (link with: g++ app.o libldpc_toolbox.a -lcudart)
    if os.path.exists(STATIC_LIB):
        os.remove(STATIC_LIB)
    cmd = ["ar", "rcs", STATIC_LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("ar failed")
    open(os.path.join(OUT_DIR, "build.log"), "w").write("\n".join(log))
    open(stamp, "w").write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
