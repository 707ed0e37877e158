"""Command line: `ber` is a drop-in for `ldpc-toolbox ber` (reference src/cli/ber.rs:39-341); `encode`,
`systematic`, `dvbs2`, `nr5g` and `ccsds` mirror the reference's subcommands that sit either side of the
hot path (src/cli/encode.rs, systematic.rs, dvbs2.rs, nr5g.rs, ccsds.rs; `--girth` is out of scope).

    python -m ldpc_toolbox_b200.cli ber <alist> --min-ebn0 0.5 --max-ebn0 1.5 --step-ebn0 0.25 \
        --decoder Minstarapproxi8 --max-iter 25 [--gpus 8] [--batch 75776] [--seed 24301]

Same flags, banner and result table as the reference; additions: --gpus, --batch,
--seed, --max-frames.  `--num-threads` is accepted and reported but the work runs on GPUs.
Modulation: BPSK or 8PSK, with the DVB-S2 bit interleaver (--interleaving n, negative = rows backwards).
"""
from __future__ import annotations

import argparse
import math
import re
import sys

from .ber import BerEngine, BerTest, Statistics


def parse_duration(s: str) -> float:
    """humantime-style durations: "90s", "1m 30s", "2h", "500ms"."""
    units = {"ms": 1e-3, "s": 1.0, "sec": 1.0, "m": 60.0, "min": 60.0, "h": 3600.0, "d": 86400.0}
    total, pos = 0.0, 0
    for m in re.finditer(r"\s*(\d+(?:\.\d+)?)\s*([a-z]+)", s):
        if m.start() != pos or m.group(2) not in units:
            raise argparse.ArgumentTypeError(f"invalid duration {s!r}")
        total += float(m.group(1)) * units[m.group(2)]
        pos = m.end()
    if pos != len(s.rstrip()) or pos == 0:
        raise argparse.ArgumentTypeError(f"invalid duration {s!r}")
    return total


def format_duration(seconds: float) -> str:
    """humantime::format_duration of whole seconds."""
    s = int(seconds)
    if s == 0:
        return "0s"
    parts = []
    for name, span in (("days", 86400), ("h", 3600), ("m", 60), ("s", 1)):
        q, s = divmod(s, span)
        if q:
            parts.append(f"{q}{name}")
    return " ".join(parts)


def rust_lower_exp(x: float, prec: int = 2, width: int = 7) -> str:
    """Rust's `{:7.2e}`: mantissa with `prec` decimals, exponent without padding or plus sign."""
    if x != x:
        s = "NaN"
    elif math.isinf(x):
        s = "inf" if x > 0 else "-inf"
    elif x == 0:
        s = f"{0:.{prec}f}e0"
    else:
        m, e = f"{x:.{prec}e}".split("e")
        s = f"{m}e{int(e)}"
    return s.rjust(width)


HEADER = ("  Eb/N0 |   Frames | Bit errs | Frame er | False de |     BER |     FER | Avg iter | Avg corr | Throughp | Elapsed\n"
          "--------|----------|----------|----------|----------|---------|---------|----------|----------|----------|----------")


def _fix1(x: float, width: int = 8) -> str:
    return ("NaN" if x != x else f"{x:.1f}").rjust(width)


def format_progress(st: Statistics, force_ldpc: bool = False) -> str:      # cli/ber.rs:320-339
    code = st.ldpc if (force_ldpc or st.bch is None) else st.bch
    return (f"{st.ebn0_db:7.2f} | {st.num_frames:8} | {code.bit_errors:8} | {code.frame_errors:8} | {st.false_decodes:8} | "
            f"{rust_lower_exp(code.ber)} | {rust_lower_exp(code.fer)} | {_fix1(st.average_iterations)} | "
            f"{_fix1(code.average_iterations_correct)} | {st.throughput_mbps:8.3f} | {format_duration(st.elapsed)}")


def write_details(f, a, k: int, n_cw: int, n: int, rate: float) -> None:   # cli/ber.rs:161-211
    w = lambda s="": f.write(s + "\n")
    w("BER TEST PARAMETERS")
    w("-------------------")
    w("Simulation:")
    w(f" - Minimum Eb/N0: {a.min_ebn0:.2f} dB")
    w(f" - Maximum Eb/N0: {a.max_ebn0:.2f} dB")
    w(f" - Eb/N0 step: {a.step_ebn0:.2f} dB")
    w(f" - Number of frame errors: {a.frame_errors}")
    if a.min_time is not None:
        w(f" - Minimum run time per Eb/N0: {format_duration(a.min_time)}")
    if a.max_time is not None:
        w(f" - Maximum run time per Eb/N0: {format_duration(a.max_time)}")
    w(f" - Number of worker threads: {a.num_threads}")
    if a.gpus:
        w(f" - Number of GPUs: {a.gpus} (frames per GPU per launch: {a.batch})")
    w("Channel:")
    w(f" - Modulation: {a.modulation}")
    w("LDPC code:")
    w(f" - alist: {a.alist}")
    if a.puncturing is not None:
        w(f" - Puncturing pattern: {a.puncturing}")
    if a.interleaving is not None:
        w(f" - Interleaving columns: {a.interleaving}")
    w(f" - Information bits (k): {k}")
    w(f" - Codeword size (N_cw): {n_cw}")
    w(f" - Frame size (N): {n}")
    w(f" - Code rate: {rate:.3f}")
    w("LDPC decoder:")
    w(f" - Implementation: {a.decoder}")
    w(f" - Maximum iterations: {a.max_iter}")
    if a.bch_max_errors > 0:
        w("BCH decoder:")
        w(f" - Maximum bit errors correctable: {a.bch_max_errors}")
    w()


def ebn0_list(min_ebn0: float, max_ebn0: float, step: float) -> list[float]:   # cli/ber.rs:106-109
    num = int(math.floor((max_ebn0 - min_ebn0) / step)) + 1
    return [min_ebn0 + k * step for k in range(max(num, 0))]


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="ldpc-toolbox", description="B200-native ldpc-toolbox hot path")
    sub = ap.add_subparsers(dest="command", required=True)
    b = sub.add_parser("ber", help="Performs a BER simulation")
    b.add_argument("alist", help="alist file for the code")
    b.add_argument("--output-file")
    b.add_argument("--output-file-ldpc")
    b.add_argument("--decoder", default="Phif64")
    b.add_argument("--modulation", default="BPSK", choices=["BPSK", "8PSK"])
    b.add_argument("--puncturing")
    b.add_argument("--interleaving", type=int)
    b.add_argument("--min-ebn0", type=float, required=True)
    b.add_argument("--max-ebn0", type=float, required=True)
    b.add_argument("--step-ebn0", type=float, required=True)
    b.add_argument("--max-iter", type=int, default=100)
    b.add_argument("--frame-errors", type=int, default=100)
    b.add_argument("--min-time", type=parse_duration)
    b.add_argument("--max-time", type=parse_duration)
    b.add_argument("--bch-max-errors", type=int, default=0)
    b.add_argument("--num-threads", type=int, default=1)
    b.add_argument("--gpus", type=int, default=1, help="GPUs of this node to shard the frames over")
    b.add_argument("--batch", type=int, default=0, help="frames per GPU per launch (default: by code size)")
    b.add_argument("--seed", type=int, default=0x5EED)
    b.add_argument("--max-frames", type=int, help="stop an Eb/N0 point after this many frames")
    e = sub.add_parser("encode", help="Performs LDPC encoding")                       # reference src/cli/encode.rs:19-32
    e.add_argument("alist", help="alist file for the code")
    e.add_argument("input", help="input file (information words as unpacked bits)")
    e.add_argument("output", help="output file (punctured words as unpacked bits)")
    e.add_argument("--puncturing")
    sy = sub.add_parser("systematic", help="Converts a parity check matrix into systematic form")   # src/cli/systematic.rs
    sy.add_argument("alist", help="alist file for the code")
    d = sub.add_parser("dvbs2", help="Generates the alist of DVB-S2 LDPCs")           # reference src/cli/dvbs2.rs:12-24
    d.add_argument("-r", "--rate", required=True)
    d.add_argument("--short", action="store_true")
    d.add_argument("--girth", action="store_true")
    n = sub.add_parser("nr5g", help="Generates the alist of 5G NR LDPCs")             # reference src/cli/nr5g.rs:8-19
    n.add_argument("--base-graph", type=int, required=True, choices=[1, 2])
    n.add_argument("--lifting-size", type=int, required=True)
    n.add_argument("--girth", action="store_true")
    c = sub.add_parser("ccsds", help="Generates the alist of CCSDS LDPCs")            # reference src/cli/ccsds.rs:10-22
    c.add_argument("-r", "--rate", required=True)
    c.add_argument("--block-size", type=int, required=True)
    c.add_argument("--girth", action="store_true")
    return ap


def run_encode(a) -> None:
    """reference src/cli/encode.rs:34-70: k unpacked bits in, n bytes out per word (the punctured word, then the
    rest of the n-byte buffer, which the reference never writes and so stays zero); a trailing partial word is dropped."""
    import numpy as np

    from .decoder import Encoder
    text = open(a.alist).read()
    first = text.split("\n", 1)[0].split()
    ncols, nrows = int(first[0]), int(first[1])
    k = ncols - nrows
    enc = Encoder(a.alist, a.puncturing or "")
    out_len = ncols
    if a.puncturing:
        pat = [int(x) for x in a.puncturing.split(",")]
        out_len = ncols // len(pat) * sum(pat)
    with open(a.input, "rb") as fi, open(a.output, "wb") as fo:
        while True:
            word = fi.read(k)
            if len(word) < k or k == 0:
                break
            cw = enc.encode(np.frombuffer(word, dtype=np.uint8), out_len)
            buf = np.zeros(ncols, dtype=np.uint8)
            buf[:out_len] = cw
            fo.write(buf.tobytes())


def parity_to_systematic(alist_text: str) -> str:
    """reference src/systematic.rs:40-93: permute the columns so that the last n-k columns (the first
    linearly independent ones, in order) are invertible; the others keep their order in front."""
    import numpy as np

    from . import codes
    lines = alist_text.split("\n")
    ncols, nrows = (int(x) for x in lines[0].split())
    if nrows > ncols:
        raise SystemExit("the parity check matrix has more rows than columns")
    cols = [[int(x) - 1 for x in lines[4 + c].split() if int(x) > 0] for c in range(ncols)]      # sparse.rs:372-386
    words = (ncols + 63) // 64
    a = np.zeros((nrows, words), dtype=np.uint64)
    for c, rows in enumerate(cols):
        for r in rows:
            a[r, c >> 6] |= np.uint64(1) << np.uint64(c & 63)       # insert(): a repeated entry stays a single one (sparse.rs:114-119)
    pivots, krow = [], 0
    for j in range(ncols):                               # linalg.rs:68-104 (row echelon form over GF(2))
        if krow >= nrows:
            break
        bit = (a[krow:, j >> 6] >> np.uint64(j & 63)) & np.uint64(1)
        nz = np.flatnonzero(bit)
        if nz.size == 0:
            continue
        s = krow + int(nz[0])
        if s != krow:
            a[[s, krow]] = a[[krow, s]]
        below = krow + 1 + np.flatnonzero((a[krow + 1:, j >> 6] >> np.uint64(j & 63)) & np.uint64(1))
        a[below] ^= a[krow]
        pivots.append(j)
        krow += 1
    if len(pivots) < nrows:
        raise SystemExit("the parity check matrix does not have full rank")
    piv = set(pivots)
    order = [c for c in range(ncols) if c not in piv] + pivots
    r = np.array([row for c in order for row in cols[c]], dtype=np.int64)
    cc = np.array([i for i, c in enumerate(order) for _ in cols[c]], dtype=np.int64)
    return codes.alist_text(codes.Edges(nrows, ncols, r, cc).finalize())


DVBS2_RATES = ("1/4", "1/3", "2/5", "1/2", "3/5", "2/3", "3/4", "4/5", "5/6", "8/9", "9/10")


def run_codes(a, out=None) -> None:
    from . import codes
    out = out or sys.stdout
    if a.girth:
        raise SystemExit("--girth belongs to the code-design tools, which are outside this framework's scope (DESIGN.md section 1)")
    if a.command == "dvbs2":
        if a.rate not in DVBS2_RATES or (a.short and a.rate == "9/10"):                  # cli/dvbs2.rs:27-58
            raise SystemExit(f"Invalid rate {a.rate} for {'short' if a.short else 'normal'} FECFRAME")
        out.write(codes.alist_for("dvbs2:R" + a.rate.replace("/", "_") + ("short" if a.short else "")))
    elif a.command == "nr5g":
        out.write(codes.alist_for(f"nr5g:{a.base_graph}:{a.lifting_size}") + "\n")       # cli/nr5g.rs:33 uses println!
    else:
        if a.rate not in ("1/2", "2/3", "4/5"):
            raise SystemExit(f"Invalid code rate {a.rate}")
        if a.block_size not in (1024, 4096, 16384):
            raise SystemExit(f"Invalid information block size k = {a.block_size}")
        out.write(codes.alist_for(f"ar4ja:{a.rate}:{a.block_size}"))


def run_ber(a, out=sys.stdout) -> list[Statistics]:
    engines = [BerEngine(a.alist, a.decoder, a.puncturing or "", device=g, modulation=a.modulation, interleaving=a.interleaving)
               for g in range(a.gpus)]
    e0 = engines[0]
    if a.batch <= 0:
        # one 512-frame tile per SM per batch for big codes — two batches are in flight per GPU, which fills the
        # SMs with two tiles each — fewer frames for small codes
        a.batch = 75776 if e0.n_cw >= 32768 else 65536
    write_details(out, a, e0.k, e0.n_cw, e0.n, e0.rate)
    files = []
    if a.output_file:
        f = open(a.output_file, "w")
        write_details(f, a, e0.k, e0.n_cw, e0.n, e0.rate)
        if a.bch_max_errors > 0:
            f.write("\nLDPC+BCH results\n\n")
        files.append((f, False))
    if a.output_file_ldpc and a.bch_max_errors > 0:
        f = open(a.output_file_ldpc, "w")
        write_details(f, a, e0.k, e0.n_cw, e0.n, e0.rate)
        f.write("\nLDPC-only results\n\n")
        files.append((f, True))
    out.write(HEADER + "\n")
    for f, _ in files:
        f.write(HEADER + "\n")
    tty = hasattr(out, "isatty") and out.isatty()
    state = {"live": False}

    def reporter(st: Statistics, final: bool):
        line = format_progress(st)
        if tty:
            out.write(("\x1b[1A\x1b[2K" if state["live"] else "") + line + "\n")
            state["live"] = not final
        elif final:
            out.write(line + "\n")
        out.flush()
        if final:
            for f, ldpc_only in files:
                f.write(format_progress(st, ldpc_only) + "\n")
                f.flush()

    test = BerTest(engines, e0.k, ebn0_list(a.min_ebn0, a.max_ebn0, a.step_ebn0), max_iterations=a.max_iter,
                   max_frame_errors=a.frame_errors, min_time=a.min_time or 0.0,
                   max_time=a.max_time if a.max_time is not None else float("inf"), bch_max_errors=a.bch_max_errors,
                   batch=a.batch, seed=a.seed, reporter=reporter, max_frames=a.max_frames)
    if test.pipeline_depth > 1:
        # the reference's workers also run ahead of the controller's stop rule (ber.rs:312-343); here by whole batches
        out.write(f"(batches of {a.batch} frames x {a.gpus} GPU(s), {test.pipeline_depth} in flight per GPU: a point may "
                  f"run up to {test.overshoot_bound()} frames past the stop rule)\n")
    stats = test.run()
    for f, _ in files:
        f.close()
    return stats


def main(argv=None) -> int:
    a = build_parser().parse_args(argv)
    if a.command == "ber":
        run_ber(a)
    elif a.command == "encode":
        run_encode(a)
    elif a.command == "systematic":
        sys.stdout.write(parity_to_systematic(open(a.alist).read()) + "\n")             # cli/systematic.rs:23 uses println!
    else:
        run_codes(a)
    return 0


if __name__ == "__main__":
    sys.exit(main())
