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
