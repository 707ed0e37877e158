"""Standard-code parity-check matrices as alist text.

The hot path only ever consumes alist text (reference: src/sparse.rs:352-389),
but the benchmark configurations need the DVB-S2 / 5G-NR / CCSDS matrices and
there is no Rust binary in this environment to print them.  This module builds
them from the standards tables in ``data/`` (see tools/extract_standard_tables.py
for provenance) using the construction rules of the standards, and writes alist
text byte-identical to the reference writer (src/sparse.rs:250-299: 1-based,
sorted, zero padded).

Construction rules restated from:
  * DVB-S2  ETSI EN 302 307-1 §5.3.2.1          (reference: src/codes/dvbs2.rs:79-98)
  * 5G NR   3GPP TS 38.212 §5.3.2               (reference: src/codes/nr5g.rs:40-53)
  * AR4JA   CCSDS 131.0-B-5 §7.4.2              (reference: src/codes/ccsds.rs:51-187)
"""
from __future__ import annotations

import os
from functools import lru_cache

import numpy as np

_DATA = os.path.join(os.path.dirname(__file__), "data")

__all__ = [
    "Edges", "alist_text", "dvbs2_names", "dvbs2", "nr5g", "ar4ja", "alist_for", "cached_alist_path",
]


class Edges:
    """A sparse binary matrix as (rows, cols) index arrays plus its shape."""

    def __init__(self, nrows: int, ncols: int, r: np.ndarray, c: np.ndarray):
        self.nrows, self.ncols = int(nrows), int(ncols)
        # toggle semantics of the AR4JA construction: entries appearing an even
        # number of times cancel; others collapse to a single one.
        key = r.astype(np.int64) * self.ncols + c.astype(np.int64)
        uniq, counts = np.unique(key, return_counts=True)
        self._toggle_key = uniq[counts % 2 == 1]
        self._insert_key = uniq
        self.r, self.c = r, c

    def finalize(self, toggle: bool = False) -> "Edges":
        key = self._toggle_key if toggle else self._insert_key
        self.r = (key // self.ncols).astype(np.int64)
        self.c = (key % self.ncols).astype(np.int64)
        return self

    @property
    def nnz(self) -> int:
        return int(self.r.size)


def alist_text(e: Edges, padding: bool = True) -> str:
    """alist writer, same text as src/sparse.rs:250-299."""
    order_c = np.lexsort((e.r, e.c))
    order_r = np.lexsort((e.c, e.r))
    cw = np.bincount(e.c, minlength=e.ncols)
    rw = np.bincount(e.r, minlength=e.nrows)
    maxc = int(cw.max()) if cw.size else 0
    maxr = int(rw.max()) if rw.size else 0
    out = [f"{e.ncols} {e.nrows}", f"{maxc} {maxr}", " ".join(map(str, cw.tolist())), " ".join(map(str, rw.tolist()))]

    def section(sorted_vals, weights, maxw):
        pos = 0
        vals = (sorted_vals + 1).tolist()
        for w in weights.tolist():
            items = vals[pos:pos + w]
            pos += w
            s = " ".join(map(str, items))
            if padding:
                if w == 0:
                    s = "0"
                s += " 0" * (maxw - max(w, 1))
            out.append(s)

    section(e.r[order_c], cw, maxc)
    section(e.c[order_r], rw, maxr)
    return "\n".join(out) + "\n"


# ----------------------------------------------------------------------------- DVB-S2
@lru_cache(maxsize=None)
def _dvbs2_tables():
    tabs = {}
    cur = None
    for line in open(os.path.join(_DATA, "dvbs2_addresses.txt")):
        t = line.split()
        if not t:
            continue
        if t[0] == "code":
            cur = {"n": int(t[3]), "q": int(t[5]), "rows": []}
            tabs[t[1]] = cur
        else:
            cur["rows"].append([int(x) for x in t])
    return tabs


def dvbs2_names():
    return list(_dvbs2_tables().keys())


def dvbs2(name: str, reference_bug: bool = False) -> Edges:
    """DVB-S2 parity-check matrix, e.g. ``dvbs2("R1_2")`` / ``dvbs2("R1_2short")``.

    m = 360*q as in the standard.  ``reference_bug=True`` reproduces the
    reference's wrong m for R3_4short (src/codes/dvbs2.rs:152) – never useful
    for decoding, kept only so the discrepancy can be demonstrated.
    """
    t = _dvbs2_tables()[name]
    n, q = t["n"], t["q"]
    m = 360 * q
    if reference_bug and name == "R3_4short":
        m = n * 14 // 15
    k = n - m
    assert k == 360 * len(t["rows"]) or reference_bug
    rr, cc = [], []
    w = np.arange(360, dtype=np.int64)
    for g, addrs in enumerate(t["rows"]):
        for x in addrs:
            rr.append((x + w * q) % m)
            cc.append(360 * g + w)
    j = np.arange(1, m, dtype=np.int64)
    rr += [np.array([0]), j, j]
    cc += [np.array([k]), j + k, j + k - 1]
    return Edges(m, n, np.concatenate(rr), np.concatenate(cc)).finalize()


# ----------------------------------------------------------------------------- 5G NR
_NR_SETS = [
    [2, 4, 8, 16, 32, 64, 128, 256], [3, 6, 12, 24, 48, 96, 192, 384], [5, 10, 20, 40, 80, 160, 320],
    [7, 14, 28, 56, 112, 224], [9, 18, 36, 72, 144, 288], [11, 22, 44, 88, 176, 352],
    [13, 26, 52, 104, 208], [15, 30, 60, 120, 240],
]


@lru_cache(maxsize=None)
def _nr_tables():
    bgs, cur, row = {}, None, None
    for line in open(os.path.join(_DATA, "nr5g_basegraphs.txt")):
        t = line.split()
        if not t:
            continue
        if t[0] == "basegraph":
            cur = []
            bgs[int(t[1])] = cur
        elif t[0] == "row":
            row = []
            cur.append(row)
        else:
            row.append([int(x) for x in t])
    return bgs


def nr5g(base_graph: int, z: int) -> Edges:
    """5G NR base graph 1/2 lifted by Z: row Z*j+r <-> col Z*k+((r+V) mod Z)."""
    bg = _nr_tables()[base_graph]
    ils = next(i for i, s in enumerate(_NR_SETS) if z in s)
    ncolb = 68 if base_graph == 1 else 52
    r = np.arange(z, dtype=np.int64)
    rr, cc = [], []
    for j, rows in enumerate(bg):
        for ent in rows:
            kcol, v = ent[0], ent[1 + ils]
            rr.append(z * j + r)
            cc.append(z * kcol + ((r + v) % z))
    return Edges(len(bg) * z, ncolb * z, np.concatenate(rr), np.concatenate(cc)).finalize()


# ----------------------------------------------------------------------------- CCSDS AR4JA
@lru_cache(maxsize=None)
def _ar4ja_tables():
    theta, phi = None, {}
    for line in open(os.path.join(_DATA, "ccsds_ar4ja.txt")):
        t = line.split()
        if not t:
            continue
        if t[0] == "theta":
            theta = [int(x) for x in t[1:]]
        else:
            phi[(int(t[1]), int(t[2]))] = [int(x) for x in t[3:]]
    return theta, phi


_AR4JA_M = {("1/2", 1024): 512, ("2/3", 1024): 256, ("4/5", 1024): 128,
            ("1/2", 4096): 2048, ("2/3", 4096): 1024, ("4/5", 4096): 512,
            ("1/2", 16384): 8192, ("2/3", 16384): 4096, ("4/5", 16384): 2048}


def ar4ja(rate: str, k: int) -> Edges:
    """CCSDS AR4JA code, rate in {"1/2","2/3","4/5"}, k in {1024,4096,16384}.
    The last M columns are punctured on transmission (puncturing "1,1,1,1,0"
    for rate 1/2)."""
    theta, phi = _ar4ja_tables()
    m = _AR4JA_M[(rate, k)]
    mlog = m.bit_length() - 1
    midx = mlog - 7
    i = np.arange(m, dtype=np.int64)

    def pi(kk):
        j = 4 * i // m
        a = (theta[kk - 1] + j) & 3
        ph = np.array([phi[(jj, kk)][midx] for jj in range(4)], dtype=np.int64)[j]
        b = (ph + i) & (m // 4 - 1)
        return (a << (mlog - 2)) + b

    extra_blocks = {"1/2": 0, "2/3": 2, "4/5": 6}[rate]
    ex = m * extra_blocks
    rr, cc = [], []

    def add(rbase, cbase, col):
        rr.append(rbase + i)
        cc.append(cbase + col)

    add(0, ex + 2 * m, i)
    add(0, ex + 4 * m, i); add(0, ex + 4 * m, pi(1))
    add(m, ex, i); add(m, ex + m, i); add(m, ex + 3 * m, i)
    add(m, ex + 4 * m, pi(2)); add(m, ex + 4 * m, pi(3)); add(m, ex + 4 * m, pi(4))
    add(2 * m, ex, i)
    add(2 * m, ex + m, pi(5)); add(2 * m, ex + m, pi(6))
    add(2 * m, ex + 3 * m, pi(7)); add(2 * m, ex + 3 * m, pi(8))
    add(2 * m, ex + 4 * m, i)
    if rate != "1/2":
        e2 = 0 if rate == "2/3" else 4 * m
        add(m, e2, pi(9)); add(m, e2, pi(10)); add(m, e2, pi(11))
        add(m, e2 + m, i)
        add(2 * m, e2, i)
        add(2 * m, e2 + m, pi(12)); add(2 * m, e2 + m, pi(13)); add(2 * m, e2 + m, pi(14))
    if rate == "4/5":
        add(m, 0, pi(21)); add(m, 0, pi(22)); add(m, 0, pi(23))
        add(m, m, i)
        add(m, 2 * m, pi(15)); add(m, 2 * m, pi(16)); add(m, 2 * m, pi(17))
        add(m, 3 * m, i)
        add(2 * m, 0, i)
        add(2 * m, m, pi(24)); add(2 * m, m, pi(25)); add(2 * m, m, pi(26))
        add(2 * m, 2 * m, i)
        add(2 * m, 3 * m, pi(18)); add(2 * m, 3 * m, pi(19)); add(2 * m, 3 * m, pi(20))
    # the reference mixes insert (first term of a block) and toggle (others);
    # within one block the permutations never collide with the first term in a
    # way that insert-vs-toggle would differ except by cancellation, which the
    # toggle semantics below reproduce.
    return Edges(3 * m, ex + 5 * m, np.concatenate(rr), np.concatenate(cc)).finalize(toggle=True)


# ----------------------------------------------------------------------------- convenience
def alist_for(spec: str) -> str:
    """``spec`` examples: ``dvbs2:R1_2``, ``nr5g:2:384``, ``ar4ja:1/2:1024``."""
    p = spec.split(":")
    if p[0] == "dvbs2":
        return alist_text(dvbs2(p[1]))
    if p[0] == "nr5g":
        return alist_text(nr5g(int(p[1]), int(p[2])))
    if p[0] == "ar4ja":
        return alist_text(ar4ja(p[1], int(p[2])))
    raise ValueError(f"unknown code spec {spec!r}")


def cached_alist_path(spec: str, cache_dir: str | None = None) -> str:
    """Write (once) the alist of ``spec`` under a cache directory and return its path."""
    cache_dir = cache_dir or os.environ.get(
        "LDPC_B200_CACHE", os.path.join(os.path.dirname(__file__), "..", "..", "build", "alist"))
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, spec.replace(":", "_").replace("/", "") + ".alist")
    if not os.path.exists(path):
        tmp = path + f".tmp{os.getpid()}"
        with open(tmp, "w") as f:
            f.write(alist_for(spec))
        os.replace(tmp, path)
    return os.path.abspath(path)
