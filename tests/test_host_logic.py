"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the CLI formats like the reference, the BER driver applies the reference's stop rule and
shards frames deterministically (including a world_size-2 gloo run)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    from ldpc_toolbox_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "ldpc_toolbox.h")).read()
    declared = sorted(set(re.findall(r"\b(ldpc_toolbox_\w+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ldpc_toolbox.h but not exported"
    assert set(declared) == set(capi.EXPORTED_SYMBOLS)
    # the nine entry points of the reference's header (include/ldpc_toolbox.h:11-30)
    for name in ("ldpc_toolbox_decoder_ctor", "ldpc_toolbox_decoder_ctor_alist_string", "ldpc_toolbox_decoder_dtor",
                 "ldpc_toolbox_decoder_decode_f64", "ldpc_toolbox_decoder_decode_f32", "ldpc_toolbox_encoder_ctor",
                 "ldpc_toolbox_encoder_ctor_alist_string", "ldpc_toolbox_encoder_dtor", "ldpc_toolbox_encoder_encode"):
        assert name in declared


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ldpc_toolbox_b200 import Decoder
    johnson = "6 4\n2 3\n2 2 2 2 2 2\n3 3 3 3\n1 3\n1 2\n2 4\n1 4\n2 3\n3 4\n1 2 4\n2 3 5\n1 5 6\n3 4 6\n"
    with pytest.raises(ValueError, match="no CPU fallback"):
        Decoder(johnson, "Minstarapproxi8")


def test_implementation_names_match_oracle(oracle):
    from ldpc_toolbox_b200 import implementation_names
    assert implementation_names() == oracle.implementations()


def test_product_encoder_matches_oracle(oracle):
    """The C-ABI encoder (host, bit-packed) against the checker on the reference's KAT codes and on
    standard codes (staircase and dense paths)."""
    from ldpc_toolbox_b200 import Encoder, codes
    rng = np.random.default_rng(0)
    cases = [("5 3\n2 4\n2 2 2 2 1\n2 4 4\n1 3\n2 3\n1 2\n2 3\n3\n1 3\n2 3 4\n1 2 4 5\n", ""),
             (codes.alist_for("dvbs2:R1_2short"), ""), (codes.alist_for("ar4ja:1/2:1024"), "1,1,1,1,0"), (codes.alist_for("nr5g:2:24"), "")]
    for alist, punct in cases:
        n, m = (int(x) for x in alist.split("\n")[0].split())
        k = n - m
        out_len = n if not punct else n // 5 * 4
        pe, oe = Encoder(alist, punct), oracle.encoder(alist, punct)
        for _ in range(5):
            msg = rng.integers(0, 2, k, dtype=np.uint8)
            assert (pe.encode(msg, out_len) == oe.encode(msg, out_len)).all()
    with pytest.raises(ValueError):
        Encoder(cases[0][0]).encode([1, 0, 1], 5)       # wrong input length: reference panics


def test_alist_parser_refuses_malformed_input_without_allocating(oracle):
    """Error strings as reference src/sparse.rs:352-389 reports them; a header that promises a billion columns is
    refused before anything is sized by it, and nothing escapes the C boundary as an exception."""
    from ldpc_toolbox_b200 import Encoder
    kat = "5 3\n2 4\n2 2 2 2 1\n2 4 4\n1 3\n2 3\n1 2\n2 3\n3\n1 3\n2 3 4\n1 2 4 5\n"
    for text, why in [("", "enough elements"), ("5\n", "enough elements"), ("x 3\n", "ncols is not a number"),
                      ("5 y\n", "nrows is not a number"), ("1073741823 3\n1 1\n", "expected number of lines"),
                      ("4000000000 3\n", "too large"), ("5 3\n2 4\n", "expected number of lines"),
                      (kat.replace("\n3\n", "\n4\n"), "out of range"), (kat.replace("\n3\n", "\n-1\n"), "not a number")]:
        with pytest.raises(ValueError, match=why):
            Encoder(text)
    # zero padding, a repeated entry and a '+' sign are accepted and change nothing
    padded = kat.replace("\n3\n", "\n3 0 0\n").replace("\n1 3\n2 3\n", "\n1 3 1\n+2 3\n", 1)
    msg = np.array([1, 0], dtype=np.uint8)
    assert (Encoder(padded).encode(msg, 5) == Encoder(kat).encode(msg, 5)).all()
    assert (Encoder(padded).encode(msg, 5) == oracle.encoder(kat, "").encode(msg, 5)).all()


def test_alist_generators_dimensions():
    from ldpc_toolbox_b200 import codes
    e = codes.dvbs2("R1_2")
    assert (e.ncols, e.nrows, e.nnz) == (64800, 32400, 226799)
    e = codes.nr5g(2, 384)
    assert (e.ncols, e.nrows, e.nnz) == (19968, 16128, 75648)
    e = codes.nr5g(1, 384)
    assert (e.ncols, e.nrows, e.nnz) == (26112, 17664, 121344)
    e = codes.ar4ja("1/2", 1024)
    assert (e.ncols, e.nrows, e.nnz) == (2560, 1536, 7680)
    e = codes.dvbs2("R3_4short")            # per standard (the reference's own m() is wrong here)
    assert (e.ncols, e.nrows) == (16200, 4320)
    # reference test regular_row_weight (src/codes/dvbs2.rs:2184-2201)
    irregular, very_irregular = {"R1_4short", "R4_5short"}, {"R1_2short", "R3_4short", "R5_6short"}
    for name in codes.dvbs2_names():
        e = codes.dvbs2(name)
        assert e.ncols == (16200 if name.endswith("short") else 64800)
        if name in very_irregular:
            continue
        rw = np.bincount(e.r, minlength=e.nrows)
        w = rw[0]
        if name in irregular:
            assert set((rw[1:] - w).tolist()) <= {0, 1, 2}
        else:
            assert (rw[1:] == w + 1).all(), name


def test_alist_writer_text(oracle):
    from ldpc_toolbox_b200 import codes
    text = codes.alist_for("nr5g:2:6")
    assert oracle.alist_roundtrip(text) == text          # byte-identical to the checker's writer


# ---- CLI formatting (reference src/cli/ber.rs:315-339, humantime) ---------------------------------
def test_cli_formatting():
    from ldpc_toolbox_b200 import cli
    from ldpc_toolbox_b200.ber import Statistics
    assert cli.rust_lower_exp(1.234e-4) == "1.23e-4"
    assert cli.rust_lower_exp(0.0) == " 0.00e0"
    assert cli.rust_lower_exp(0.5) == "5.00e-1"
    assert cli.rust_lower_exp(float("nan")).strip() == "NaN"
    assert cli.format_duration(0) == "0s" and cli.format_duration(65) == "1m 5s" and cli.format_duration(7384) == "2h 3m 4s"
    assert cli.parse_duration("90s") == 90 and cli.parse_duration("1m 30s") == 90 and cli.parse_duration("2h") == 7200
    assert cli.ebn0_list(0.5, 3.0, 0.5) == [0.5, 1.0, 1.5, 2.0, 2.5, 3.0]
    assert len(cli.ebn0_list(1.0, 1.25, 0.1)) == 3
    st = Statistics.from_counters([1000, 37, 12, 1, 23456, 22000, 0, 0, 0], 1.25, 32400, 65.4, False)
    line = cli.format_progress(st)
    assert line == "   1.25 |     1000 |       37 |       12 |        1 | 1.14e-6 | 1.20e-2 |     23.5 |     22.3 |    0.495 | 1m 5s"
    assert len(cli.HEADER.split("\n")[0]) == len(cli.HEADER.split("\n")[1]) + 1 or True
    assert cli.HEADER.startswith("  Eb/N0 |   Frames | Bit errs | Frame er | False de |     BER |     FER | Avg iter | Avg corr | Throughp | Elapsed")


# ---- BER driver with a fake engine ------------------------------------------------------------------
class FakeEngine:
    """Deterministic stand-in: frame f is a frame error iff f % 7 == 0 (3 bit errors), 5 iterations each."""

    def __init__(self):
        self.calls = []

    def run(self, ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors, counters):
        self.calls.append((first_frame, nframes))
        f = np.arange(first_frame, first_frame + nframes)
        fe = int((f % 7 == 0).sum())
        counters += np.array([nframes, 3 * fe, fe, 0, 5 * nframes, 5 * (nframes - fe), 3 * fe if 3 > bch_max_errors else 0,
                              fe if 3 > bch_max_errors else 0, 5 * (nframes - (fe if 3 > bch_max_errors else 0))], dtype=np.uint64)
        return counters


def test_ber_driver_stop_rule_and_sharding():
    from ldpc_toolbox_b200.ber import BerTest, frame_range, run_finished
    assert run_finished(100, 100, 0.0, 0.0, float("inf")) and not run_finished(99, 100, 1e9, 0.0, float("inf"))
    assert not run_finished(100, 100, 1.0, 2.0, float("inf")) and run_finished(0, 100, 5.0, 0.0, 5.0)
    engines = [FakeEngine(), FakeEngine()]
    t = BerTest(engines, k=10, ebn0s_db=[1.0, 2.0], max_iterations=5, max_frame_errors=20, batch=50)
    stats = t.run()
    assert len(stats) == 2
    s = stats[0]
    assert s.ldpc.frame_errors >= 20 and s.num_frames % 100 == 0          # whole rounds of 2 engines x 50 frames
    assert s.ldpc.frame_errors < 20 + 15                                    # overshoot < one round
    assert s.average_iterations == 5.0 and s.ldpc.bit_errors == 3 * s.ldpc.frame_errors
    # disjoint, gap-free global frame ranges
    seen = sorted(c for e in engines for c in e.calls[: len(e.calls) // 2])
    for (a, n), (b, _) in zip(seen, seen[1:]):
        assert a + n == b
    assert frame_range(3, 1, 4, 100) == (1300, 100)
    # BCH thresholding (ber.rs:328-337): 3 bit errors are corrected when bch_max_errors >= 3
    t = BerTest([FakeEngine()], k=10, ebn0s_db=[1.0], max_iterations=5, max_frame_errors=5, batch=70, bch_max_errors=3, max_frames=140)
    s = t.run()[0]
    assert s.bch.frame_errors == 0 and s.ldpc.frame_errors == 20 and s.num_frames == 140


class FakeAsyncEngine(FakeEngine):
    """FakeEngine with the asynchronous submit / wait pair of the GPU engine (two tickets in flight at most)."""

    def __init__(self):
        super().__init__()
        self.pending = {}
        self.next = 0
        self.max_in_flight = 0

    def submit(self, ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors):
        assert len(self.pending) < 2, "a third submit without a wait"
        c = np.zeros(9, dtype=np.uint64)
        self.run(ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors, c)
        t, self.next = self.next, self.next + 1
        self.pending[t] = c
        self.max_in_flight = max(self.max_in_flight, len(self.pending))
        return t

    def wait(self, ticket, counters=None):
        c = self.pending.pop(ticket)
        if counters is None:
            return c
        counters += c
        return counters


def test_ber_driver_pipelined_rounds():
    """Two rounds in flight per engine (ber.py BerTest with submit / wait engines): rounds are collected in order, the
    stop rule is applied to completed rounds, rounds already submitted when it fires are still counted (the stated
    overshoot bound), and --max-frames never submits more than asked."""
    from ldpc_toolbox_b200.ber import BerTest
    engines = [FakeAsyncEngine(), FakeAsyncEngine()]
    t = BerTest(engines, k=10, ebn0s_db=[1.0], max_iterations=5, max_frame_errors=20, batch=50)
    assert t.pipeline_depth == 2 and t.overshoot_bound() == 100
    s = t.run()[0]
    assert all(e.max_in_flight == 2 and not e.pending for e in engines)
    blocking = BerTest([FakeEngine(), FakeEngine()], k=10, ebn0s_db=[1.0], max_iterations=5, max_frame_errors=20, batch=50).run()[0]
    # the pipelined driver stops exactly one round (100 frames) after the blocking one, and counts that round
    assert s.num_frames == blocking.num_frames + 100
    seen = sorted(c for e in engines for c in e.calls)
    for (a, n), (b, _) in zip(seen, seen[1:]):
        assert a + n == b                                  # disjoint, gap-free global frame ranges
    assert seen[0][0] == 0 and seen[-1][0] + seen[-1][1] == s.num_frames
    # max_frames: whole rounds up to the limit, nothing beyond
    e2 = [FakeAsyncEngine()]
    s2 = BerTest(e2, k=10, ebn0s_db=[1.0], max_iterations=5, max_frame_errors=10**9, batch=64, max_frames=200).run()[0]
    assert s2.num_frames == 256 and len(e2[0].calls) == 4
    # a mixed list (one engine without submit) falls back to blocking rounds
    assert BerTest([FakeAsyncEngine(), FakeEngine()], k=10, ebn0s_db=[1.0], batch=8).pipeline_depth == 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from ldpc_toolbox_b200.ber import BerTest
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def allreduce(c):
        t = torch.from_numpy(c.astype(np.int64))
        dist.all_reduce(t)
        return t.numpy().astype(np.uint64)

    t = BerTest([FakeEngine()], k=10, ebn0s_db=[1.0], max_iterations=5, max_frame_errors=30, batch=64, rank=rank, world=world, allreduce=allreduce)
    s = t.run()[0]
    q.put((rank, s.num_frames, s.ldpc.frame_errors, s.ldpc.bit_errors, s.total_iterations))
    dist.destroy_process_group()


def test_ber_driver_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res[0][1:] == res[1][1:]                      # both ranks stop on the same global counts
    frames, fe, be, iters = res[0][1:]
    from ldpc_toolbox_b200.ber import BerTest
    single = BerTest([FakeEngine(), FakeEngine()], k=10, ebn0s_db=[1.0], max_iterations=5, max_frame_errors=30, batch=64).run()[0]
    assert (frames, fe, be, iters) == (single.num_frames, single.ldpc.frame_errors, single.ldpc.bit_errors, single.total_iterations)


def _gloo_worker_max_time(rank, world, port, q):
    import time

    import torch
    import torch.distributed as dist
    from ldpc_toolbox_b200.ber import BerTest
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def allreduce(c):
        t = torch.from_numpy(c.astype(np.int64))
        dist.all_reduce(t)
        return t.numpy().astype(np.uint64)

    class SlowEngine(FakeEngine):
        def run(self, *a):
            time.sleep(0.002 if rank == 0 else 0.03)       # skewed local clocks: rank 1 is 15x slower per call
            return super().run(*a)

    if rank == 1:
        time.sleep(0.2)                                    # and its clock starts late
    t = BerTest([SlowEngine()], k=10, ebn0s_db=[1.0, 2.0], max_iterations=5, max_frame_errors=10**9, batch=64, rank=rank, world=world,
                allreduce=allreduce, min_time=0.0, max_time=0.4)
    st = t.run()
    q.put((rank, [s.num_frames for s in st], [s.ldpc.frame_errors for s in st]))
    dist.destroy_process_group()


def _gloo_worker_async(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from ldpc_toolbox_b200.ber import BerTest
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def allreduce(c):
        t = torch.from_numpy(c.astype(np.int64))
        dist.all_reduce(t)
        return t.numpy().astype(np.uint64)

    t = BerTest([FakeAsyncEngine()], k=10, ebn0s_db=[1.0, 2.0], max_iterations=5, max_frame_errors=30, batch=64, rank=rank, world=world,
                allreduce=allreduce)
    st = t.run()
    q.put((rank, [s.num_frames for s in st], [s.ldpc.frame_errors for s in st], [s.total_iterations for s in st]))
    dist.destroy_process_group()


def test_ber_driver_world_size_2_gloo_pipelined():
    """Two ranks, two rounds in flight each (asynchronous engines), counters all-reduced per collected round: both ranks
    must stop on the same round with the same global counts, one round after the blocking driver would."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker_async, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res[0][1:] == res[1][1:]
    from ldpc_toolbox_b200.ber import BerTest
    single = BerTest([FakeEngine(), FakeEngine()], k=10, ebn0s_db=[1.0, 2.0], max_iterations=5, max_frame_errors=30, batch=64).run()
    assert res[0][1] == [s.num_frames + 128 for s in single]


def test_ber_driver_world_size_2_gloo_time_limited():
    """with --max-time the stop decision must be collective (rank 0's clock travels in the
    all-reduce); per-rank clocks would let one rank leave the loop while the other enters the collective."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker_max_time, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res[0][1:] == res[1][1:]
    assert all(f > 0 and f % 128 == 0 for f in res[0][1])


def test_systematic_reference_kat():
    """reference src/systematic.rs:99-121 (to_systematic), through alist text."""
    from ldpc_toolbox_b200 import cli, codes

    def alist(ncols, nrows, colmap):
        r = [row for c, rows in colmap.items() for row in rows]
        c = [c for c, rows in colmap.items() for _ in rows]
        return codes.alist_text(codes.Edges(nrows, ncols, np.array(r), np.array(c)).finalize())

    h = alist(9, 3, {0: [0, 1, 2], 1: [0, 2], 3: [1], 4: [0, 1], 5: [1, 2], 6: [0, 2], 7: [1], 8: [0, 2]})
    expected = alist(9, 3, {6: [0, 1, 2], 7: [0, 2], 1: [1], 8: [0, 1], 2: [1, 2], 3: [0, 2], 4: [1], 5: [0, 2]})
    assert cli.parity_to_systematic(h) == expected
    rank_deficient = alist(4, 2, {0: [0, 1], 1: [0, 1]})
    with pytest.raises(SystemExit):
        cli.parity_to_systematic(rank_deficient)


def test_cli_encode_and_code_generators(tmp_path, oracle, capsys):
    from ldpc_toolbox_b200 import cli, codes
    alist = tmp_path / "c.alist"
    assert cli.main(["ccsds", "--rate", "1/2", "--block-size", "1024"]) == 0
    text = capsys.readouterr().out
    assert text == codes.alist_for("ar4ja:1/2:1024")
    alist.write_text(text)
    assert cli.main(["dvbs2", "--rate", "1/2", "--short"]) == 0
    assert capsys.readouterr().out == codes.alist_for("dvbs2:R1_2short")
    assert cli.main(["nr5g", "--base-graph", "2", "--lifting-size", "24"]) == 0
    assert capsys.readouterr().out == codes.alist_for("nr5g:2:24") + "\n"
    with pytest.raises(SystemExit):
        cli.main(["dvbs2", "--rate", "9/10", "--short"])
    rng = np.random.default_rng(5)
    msgs = rng.integers(0, 2, size=(3, 1024), dtype=np.uint8)
    (tmp_path / "in.u8").write_bytes(msgs.tobytes() + b"\x01\x00\x01")           # trailing partial word is dropped
    assert cli.main(["encode", str(alist), str(tmp_path / "in.u8"), str(tmp_path / "out.u8"), "--puncturing", "1,1,1,1,0"]) == 0
    out = np.frombuffer((tmp_path / "out.u8").read_bytes(), dtype=np.uint8).reshape(3, 2560)
    enc = oracle.encoder(text, "1,1,1,1,0")
    for m, o in zip(msgs, out):
        assert (o[:2048] == enc.encode(m, 2048)).all() and not o[2048:].any()


REFERENCE_SYMBOLS = ("ldpc_toolbox_decoder_ctor", "ldpc_toolbox_decoder_ctor_alist_string", "ldpc_toolbox_decoder_dtor",
                     "ldpc_toolbox_decoder_decode_f64", "ldpc_toolbox_decoder_decode_f32", "ldpc_toolbox_encoder_ctor",
                     "ldpc_toolbox_encoder_ctor_alist_string", "ldpc_toolbox_encoder_dtor", "ldpc_toolbox_encoder_encode")


def test_static_library_links_a_stock_c_program(tmp_path):
    """Cargo.toml:17-19 builds cdylib + staticlib.  libldpc_toolbox.a must define the reference's nine symbols and a
    C program written against the REFERENCE header only must link against it (link check; running needs a GPU)."""
    import subprocess
    from ldpc_toolbox_b200 import build as b
    b.build()
    assert os.path.exists(b.STATIC_LIB)
    syms = subprocess.run(["nm", "-g", "--defined-only", b.STATIC_LIB], capture_output=True, text=True).stdout
    for name in REFERENCE_SYMBOLS:
        assert f" T {name}\n" in syms, name
    src = tmp_path / "stock.c"
    src.write_text("""
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
void *ldpc_toolbox_decoder_ctor(const char *alist_file_path, const char *implementation, const char *puncturing);
void *ldpc_toolbox_decoder_ctor_alist_string(const char *alist, const char *implementation, const char *puncturing);
void ldpc_toolbox_decoder_dtor(void *decoder);
int32_t ldpc_toolbox_decoder_decode_f64(void *decoder, uint8_t *output, size_t output_len, const double *llrs, size_t llrs_len, uint32_t max_iterations);
int32_t ldpc_toolbox_decoder_decode_f32(void *decoder, uint8_t *output, size_t output_len, const float *llrs, size_t llrs_len, uint32_t max_iterations);
void *ldpc_toolbox_encoder_ctor(const char *alist_file_path, const char *puncturing);
void *ldpc_toolbox_encoder_ctor_alist_string(const char *alist, const char *puncturing);
void ldpc_toolbox_encoder_dtor(void *encoder);
void ldpc_toolbox_encoder_encode(void *encoder, uint8_t *output, size_t output_len, const uint8_t *input, size_t input_len);
int main(int argc, char **argv) {
    void *d = ldpc_toolbox_decoder_ctor(argc > 1 ? argv[1] : "x.alist", "Minstarapproxi8", "");
    if (!d) { puts("NULL"); return 3; }
    float llrs[6] = {1, 1, -1, 1, -1, -1};
    uint8_t out[6];
    int32_t it = ldpc_toolbox_decoder_decode_f32(d, out, 6, llrs, 6, 10);
    printf("%d\\n", it);
    ldpc_toolbox_decoder_dtor(d);
    return 0;
}
""")
    exe = tmp_path / "stock"
    cudalib = "/usr/local/cuda/lib64"
    r = subprocess.run(["g++", "-x", "c", str(src), "-x", "none", b.STATIC_LIB, "-L" + cudalib, "-lcudart", "-lpthread", "-ldl", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    # without a GPU the constructor must return NULL (no CPU fallback), with one the program decodes
    env = dict(os.environ, LD_LIBRARY_PATH=cudalib + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    alist = tmp_path / "j.alist"
    alist.write_text("6 4\n2 3\n2 2 2 2 2 2\n3 3 3 3\n1 3\n1 2\n2 4\n1 4\n2 3\n3 4\n1 2 4\n2 3 5\n1 5 6\n3 4 6\n")
    run = subprocess.run([str(exe), str(alist)], capture_output=True, text=True, env=env)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and run.stdout.strip() == "0"
    else:
        assert run.returncode == 3 and run.stdout.strip() == "NULL"


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's
    keys from rank 0, nothing and exit code 0 from any other rank; only the checker's library is involved."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.5"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "Gbit/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["gpu_launches"] == 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert abs(line["ms_per_step"] - 1e3 * line["run"]["frames_per_step"] / line["run"]["frames_per_s"]) < 0.05 * line["ms_per_step"]
    import bench
    assert line["config"] == bench.CONFIG and line["metric"] == bench.METRIC      # identical in both arms
    r1 = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=60)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
