"""Shared stimulus generation for the parity tests (test infrastructure)."""
import numpy as np


def random_code_alist(rng, n, m, col_w=3, extra_heavy_rows=0, heavy_deg=14):
    """Random sparse H as alist text: every row degree >= 2, column lists in random order (so
    cols[v] order differs from sorted order), optionally a few heavy rows (generic-degree path)."""
    while True:
        cols = [rng.choice(m, size=min(int(col_w if np.isscalar(col_w) else rng.choice(col_w)), m), replace=False).tolist()
                for _ in range(n)]
        for r in range(extra_heavy_rows):
            for c in rng.choice(n, size=heavy_deg, replace=False):
                if r not in cols[c]:
                    cols[c].append(r)
        rw = np.bincount(np.concatenate(cols), minlength=m)
        if rw.min() >= 2:
            break
    maxc = max(len(c) for c in cols)
    lines = [f"{n} {m}", f"{maxc} {int(rw.max())}", " ".join(str(len(c)) for c in cols), " ".join(map(str, rw.tolist()))]
    lines += [" ".join(str(r + 1) for r in c) for c in cols]
    lines += ["0"] * m
    return "\n".join(lines) + "\n"


def random_ira_alist(rng, k, m, info_row_w=(1, 2, 3, 5, 7, 9), heavy_rows=()):
    """Random IRA (staircase) code as alist text: H = [H0 | S], S the m x m dual-diagonal staircase of
    DVB-S2 (reference src/codes/dvbs2.rs:92-96: parity column j touches rows j and j+1), H0 random with
    row weights drawn from info_row_w (so row degrees mix the register path, the fused path and, for
    `heavy_rows`, the generic-degree path).  Column lists of the info part are in random order."""
    n = k + m
    cols = [[] for _ in range(n)]
    for r in range(m):
        w = int(rng.choice(info_row_w)) if r not in heavy_rows else 13
        for c in rng.choice(k, size=min(w, k), replace=False):
            cols[int(c)].append(r)
    for c in range(k):
        if not cols[c]:
            cols[c].append(int(rng.integers(m)))
        rng.shuffle(cols[c])
    for j in range(m):
        cols[k + j] = [j, j + 1] if j + 1 < m else [j]
    rw = np.bincount(np.concatenate([np.array(c) for c in cols]), minlength=m)
    maxc = max(len(c) for c in cols)
    lines = [f"{n} {m}", f"{maxc} {int(rw.max())}", " ".join(str(len(c)) for c in cols), " ".join(map(str, rw.tolist()))]
    lines += [" ".join(str(r + 1) for r in c) for c in cols]
    lines += ["0"] * m
    return "\n".join(lines) + "\n"


def awgn_llrs(rng, codewords, sigma, dtype=np.float32):
    """BPSK (bit 0 -> -1, bit 1 -> +1; reference modulation.rs:87-95) + AWGN, LLR = -2 y / sigma^2."""
    sym = np.where(np.asarray(codewords) == 1, 1.0, -1.0)
    y = sym + sigma * rng.standard_normal(sym.shape)
    return (-2.0 / sigma**2 * y).astype(dtype)


def encoded_frames(oracle_encoder, rng, k, n, nframes):
    msgs = rng.integers(0, 2, size=(nframes, k), dtype=np.uint8)
    cws = np.stack([oracle_encoder.encode(m, n) for m in msgs])
    return msgs, cws


def sigma_for(ebn0_db, rate):
    return float(np.sqrt(0.5 / (rate * 10 ** (np.float32(ebn0_db) / 10))))
