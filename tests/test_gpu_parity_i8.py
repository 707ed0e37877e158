"""GPU parity: the CUDA flooding i8 decoders (K1) against the CPU checker, through the C-ABI,
bit-exact on decoded words AND iteration counts (BASELINE.json north_star).  Needs a B200."""
import os
import zlib

import numpy as np
import pytest

import helpers
from ldpc_toolbox_b200 import Decoder, codes, implementation_names

pytestmark = pytest.mark.gpu

JOHNSON = "6 4\n2 3\n2 2 2 2 2 2\n3 3 3 3\n1 3\n1 2\n2 4\n1 4\n2 3\n3 4\n1 2 4\n2 3 5\n1 5 6\n3 4 6\n"
I8_FLOOD = [n for n in implementation_names() if "i8" in n and not n.startswith("HL")]


def compare(oracle, alist, impl, llrs, max_iter, puncturing="", out_len=None, label=""):
    dec = Decoder(alist, impl, puncturing)
    ref = oracle.decoder(alist, impl, puncturing)
    out, its = dec.decode_batch(llrs, max_iter, output_len=out_len)
    rout, rits = ref.decode_batch(llrs, max_iter, out_len=out_len)
    bad_it = np.nonzero(its != rits)[0]
    assert bad_it.size == 0, f"{label}{impl}: iteration mismatch at frames {bad_it[:8]}: gpu {its[bad_it[:8]]} ref {rits[bad_it[:8]]}"
    bad = np.nonzero((out != rout).any(axis=1))[0]
    assert bad.size == 0, f"{label}{impl}: word mismatch at frames {bad[:8]}"
    return its


def test_a10_vector_and_single_frame_api(oracle):
    cw = [0, 0, 1, 0, 1, 1]
    llr = np.array([-1.3863, 1.3863, -1.3863, 1.3863, -1.3863, -1.3863])   # bit 0 flipped
    for impl, iters in (("Minstarapproxi8", 2), ("Aminstari8", 2)):
        dec = Decoder(JOHNSON, impl)
        out, it = dec.decode(llr, 100)
        assert it == iters and out.tolist() == cw
        out, it = dec.decode(llr.astype(np.float32), 100, output_len=3)
        assert it == iters and out.tolist() == cw[:3]
        out, it = dec.decode(llr, 1)
        assert it == -1
        clean = np.array([1.3863 if b == 0 else -1.3863 for b in cw])
        out, it = dec.decode(clean, 100)
        assert it == 0 and out.tolist() == cw


@pytest.mark.parametrize("impl", I8_FLOOD)
def test_johnson_all_variants(oracle, impl):
    rng = np.random.default_rng(11)
    cws = np.tile(np.array([0, 0, 1, 0, 1, 1], dtype=np.uint8), (300, 1))
    llrs = np.concatenate([helpers.awgn_llrs(rng, cws[:100], s, np.float64) for s in (0.4, 0.8, 1.3)])
    llrs[5] = 0.0            # all-erasure frame: zero LLR decides 1
    llrs[6, 2] = np.nan
    llrs[7] *= 1e6           # saturates the quantiser
    compare(oracle, JOHNSON, impl, llrs, 20)


@pytest.mark.parametrize("impl", I8_FLOOD)
def test_random_irregular_codes(oracle, impl):
    rng = np.random.default_rng(hash(impl) % 2**32)
    for trial in range(3):
        n, m = [(96, 48), (200, 80), (64, 40)][trial]
        alist = helpers.random_code_alist(rng, n, m, col_w=[1, 2, 3, 4, 9], extra_heavy_rows=2 if trial else 0)
        enc = oracle.encoder(alist) if trial == 99 else None
        cws = np.zeros((520, n), dtype=np.uint8)   # all-zero word is in every code; i8 rules are still asymmetric in LLR=0
        llrs = np.concatenate([helpers.awgn_llrs(rng, cws[:130], s) for s in (0.3, 0.6, 0.9, 1.4)])
        its = compare(oracle, alist, impl, llrs, 12, label=f"trial {trial} ")
        assert (its == -1).any() or (its > 0).any()


@pytest.mark.parametrize("impl", ["Minstarapproxi8", "Aminstari8JonesPartialHardLimitDeg1Clip", "Minstarapproxi8JonesPartialHardLimit", "Aminstari8"])
def test_512_frame_tiles(oracle, impl, monkeypatch):
    """The 512-frame-tile kernel variant (4 words per lane), forced through LDPC_B200_NW."""
    monkeypatch.setenv("LDPC_B200_NW", "4")
    rng = np.random.default_rng(77)
    for trial, (n, m) in enumerate([(96, 48), (150, 70)]):
        alist = helpers.random_code_alist(rng, n, m, col_w=[1, 2, 3, 4, 9], extra_heavy_rows=trial * 2)
        nframes = [700, 513][trial]
        llrs = np.concatenate([helpers.awgn_llrs(rng, np.zeros((nframes // 4 + 1, n), dtype=np.uint8), s) for s in (0.3, 0.6, 0.9, 1.4)])[:nframes]
        compare(oracle, alist, impl, llrs, 12, label=f"nw4 trial {trial} ")
    alist = codes.alist_for("dvbs2:R1_2short")
    enc = oracle.encoder(alist)
    k = 16200 - 9000
    msgs, cws = helpers.encoded_frames(enc, rng, k, 16200, 64)
    llrs = helpers.awgn_llrs(rng, cws, helpers.sigma_for(1.0, k / 16200))
    compare(oracle, alist, impl, llrs, 20, out_len=k, label="nw4 dvbs2 short ")


@pytest.mark.parametrize("nw", ["1", "4"])
@pytest.mark.parametrize("impl", I8_FLOOD)
def test_staircase_fusion(oracle, impl, nw, monkeypatch):
    """IRA codes (H = [H0 | staircase]): the degree-2 parity variables are updated inside the check pass
    (decoder_impl.hpp RowMeta).  Row degrees 2..10 and 14 mix fused rows, rows on the register path that
    cannot fuse (degree > 8), generic-degree rows and chunk boundaries (m not a multiple of 32); LDPC_B200_FUSE=0
    must give the same answer through the ordinary variable pass."""
    monkeypatch.setenv("LDPC_B200_NW", nw)
    rng = np.random.default_rng(zlib.crc32(impl.encode()) + int(nw))
    for trial, (k, m, heavy) in enumerate([(90, 70, ()), (150, 101, (5, 40)), (40, 33, ())]):
        alist = helpers.random_ira_alist(rng, k, m, heavy_rows=heavy)
        enc = oracle.encoder(alist)
        nframes = [300, 600, 130][trial]
        msgs, cws = helpers.encoded_frames(enc, rng, k, k + m, 16)
        llrs = np.concatenate([helpers.awgn_llrs(rng, cws[np.arange(nframes // 3) % 16], s) for s in (0.45, 0.7, 1.0)])
        its = compare(oracle, alist, impl, llrs, 14, label=f"ira trial {trial} nw {nw} ")
        assert (its > 0).any()
        if trial == 1:
            monkeypatch.setenv("LDPC_B200_FUSE", "0")
            compare(oracle, alist, impl, llrs, 14, label=f"ira unfused nw {nw} ")
            monkeypatch.delenv("LDPC_B200_FUSE")


def test_two_lane_pipeline_and_submit_wait(oracle):
    """Host-buffer batches larger than half a GPU-filling launch are cut into half-launch chunks that alternate
    between two compute streams / workspaces (decoder.cu submit_batch); consecutive submits share the pipeline.
    Every frame must come out as from the checker, in place, for ragged tails and for interleaved tickets."""
    import torch
    rng = np.random.default_rng(123)
    alist = helpers.random_ira_alist(rng, 60, 36)
    n = 96
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    nframes = sm * 512 * 2 + 300                      # two half-launch chunks + a ragged third
    enc = oracle.encoder(alist)
    msgs, cws = helpers.encoded_frames(enc, rng, 60, n, 32)
    llrs = np.concatenate([helpers.awgn_llrs(rng, cws[np.arange(nframes // 2 + 1) % 32], s) for s in (0.6, 0.9)])[:nframes]
    rout, rits = oracle.decoder(alist, "Minstarapproxi8").decode_batch(llrs, 12, nthreads=os.cpu_count())
    dec = Decoder(alist, "Minstarapproxi8")
    out, its = dec.decode_batch(llrs, 12)
    assert (its == rits).all() and (out == rout).all()
    # three tickets in flight over pinned buffers, waited out of order
    h_llrs = torch.from_numpy(llrs).pin_memory()
    h_out = torch.zeros((nframes, n), dtype=torch.uint8).pin_memory()
    h_it = torch.zeros(nframes, dtype=torch.int32).pin_memory()
    cuts = [0, sm * 512 + 77, sm * 512 + 78, nframes]
    tickets = []
    for a, b in zip(cuts, cuts[1:]):
        tickets.append(dec.submit_batch_ptr(h_llrs[a:].data_ptr(), False, n, b - a, 12, h_out[a:].data_ptr(), n, n, h_it[a:].data_ptr()))
    dec.wait(tickets[-1])
    dec.wait(tickets[0])
    assert (h_it.numpy() == rits).all() and (h_out.numpy() == rout).all()


@pytest.mark.parametrize("nw", ["1", "4"])
@pytest.mark.parametrize("code,ebn0", [("dvbs2:R3_5short", 1.8), ("dvbs2:R8_9short", 4.0)])
@pytest.mark.parametrize("impl", ["Minstarapproxi8", "Minstarapproxi8JonesPartialHardLimitDeg1Clip", "Aminstari8", "Aminstari8JonesPartialHardLimit"])
def test_wide_row_kernels_on_dvbs2_high_rates(oracle, impl, code, ebn0, nw, monkeypatch):
    """Check degrees 11 (one 16-line stage per warp) and 27 (32-line stage): rows folded from shared memory by
    check_wide, for both rule families and the clipping variants, on 128- and 512-frame tiles."""
    monkeypatch.setenv("LDPC_B200_NW", nw)
    alist = codes.alist_for(code)
    first = alist.split("\n", 1)[0].split()
    n, k = int(first[0]), int(first[0]) - int(first[1])
    rng = np.random.default_rng(zlib.crc32(f"{impl}{code}{nw}".encode()))
    enc = oracle.encoder(alist)
    msgs, cws = helpers.encoded_frames(enc, rng, k, n, 8)
    llrs = np.concatenate([helpers.awgn_llrs(rng, cws[np.arange(40) % 8], helpers.sigma_for(e, k / n)) for e in (ebn0 - 0.5, ebn0 + 0.3)])
    its = compare(oracle, alist, impl, llrs, 20, out_len=k, label=f"{code} nw {nw} ")
    assert (its > 0).any()


def test_async_api_edge_cases(oracle):
    """Empty batches, tickets waited twice / out of range, posterior hook on an unsupported decoder."""
    rng = np.random.default_rng(3)
    alist = helpers.random_code_alist(rng, 60, 30)
    dec = Decoder(alist, "Minstarapproxi8")
    llrs = helpers.awgn_llrs(rng, np.zeros((5, 60), dtype=np.uint8), 0.8)
    out = np.zeros((5, 60), dtype=np.uint8)
    it = np.zeros(5, dtype=np.int32)
    t0 = dec.submit_batch_ptr(llrs.ctypes.data, False, 60, 0, 10, out.ctypes.data, 60, 60, it.ctypes.data)      # nothing to do
    t1 = dec.submit_batch_ptr(llrs.ctypes.data, False, 60, 5, 10, out.ctypes.data, 60, 60, it.ctypes.data)      # pageable buffers work too
    dec.wait(t1)
    dec.wait(t0)                                   # already retired: a no-op
    dec.wait(t1)
    rout, rits = oracle.decoder(alist, "Minstarapproxi8").decode_batch(llrs, 10)
    assert (it == rits).all() and (out == rout).all()
    with pytest.raises(ValueError):
        dec.submit_batch_ptr(llrs.ctypes.data, False, 59, 5, 10, out.ctypes.data, 60, 60, it.ctypes.data)       # wrong llrs_len
    with pytest.raises(ValueError):
        dec.decode_batch_posteriors(llrs, 10)                                                                       # int8 decoder


def test_ragged_batch_and_strides(oracle):
    rng = np.random.default_rng(5)
    alist = helpers.random_code_alist(rng, 120, 60)
    for nframes in (1, 127, 128, 129, 300):
        llrs = helpers.awgn_llrs(rng, np.zeros((nframes, 120), dtype=np.uint8), 0.9)
        compare(oracle, alist, "Minstarapproxi8", llrs, 10, out_len=60, label=f"nframes {nframes} ")
    dec = Decoder(alist, "Minstarapproxi8")
    llrs = helpers.awgn_llrs(rng, np.zeros((10, 120), dtype=np.uint8), 0.9)
    big = np.full((10, 200), 7, dtype=np.uint8)
    out, its = dec.decode_batch(llrs, 10, output_len=50, out=big[:, :50])
    assert (big[:, 50:] == 7).all()
    ref, rits = oracle.decoder(alist, "Minstarapproxi8").decode_batch(llrs, 10, out_len=50)
    assert (big[:, :50] == ref).all() and (its == rits).all()


def test_max_iterations_zero(oracle):
    rng = np.random.default_rng(6)
    alist = helpers.random_code_alist(rng, 60, 30)
    dec = Decoder(alist, "Minstarapproxi8")
    llrs = helpers.awgn_llrs(rng, np.zeros((40, 60), dtype=np.uint8), 0.7)
    out, its = dec.decode_batch(llrs, 0)
    ref = oracle.decoder(alist, "Minstarapproxi8")
    _, rits = ref.decode_batch(llrs, 0)
    assert (its == rits).all()                       # 0 for clean frames, -1 otherwise
    ok = its == 0
    assert (out[ok] == (llrs[ok] <= 0)).all()
    # documented deviation: failing frames return the raw-sign hard decision (the reference returns stale state)
    assert (out[~ok] == (llrs[~ok] <= 0)).all()


def test_argument_errors():
    dec = Decoder(JOHNSON, "Minstarapproxi8")
    with pytest.raises(ValueError):
        dec.decode(np.zeros(5), 10)
    with pytest.raises(ValueError):
        dec.decode(np.zeros(6), 10, output_len=7)
    with pytest.raises(ValueError):
        Decoder(JOHNSON, "minstarapproxi8")
    with pytest.raises(ValueError):
        Decoder(JOHNSON, "Minstarapproxi8", "1,x")
    with pytest.raises(ValueError):
        Decoder("garbage", "Minstarapproxi8")


def test_puncturing_ar4ja(oracle):
    alist = codes.alist_for("ar4ja:1/2:1024")
    rng = np.random.default_rng(8)
    enc = oracle.encoder(alist, "1,1,1,1,0")
    msgs = rng.integers(0, 2, size=(200, 1024), dtype=np.uint8)
    tx = np.stack([enc.encode(m, 2048) for m in msgs])
    sigma = helpers.sigma_for(2.0, 0.5)
    llrs = helpers.awgn_llrs(rng, tx, sigma)
    for impl in ("Minstarapproxi8", "Aminstari8Jones", "Minstarapproxi8Deg1Clip"):
        its = compare(oracle, alist, impl, llrs, 30, puncturing="1,1,1,1,0", out_len=1024)
    assert (its > 0).any()


def test_dvbs2_short(oracle):
    alist = codes.alist_for("dvbs2:R1_2short")
    rng = np.random.default_rng(9)
    enc = oracle.encoder(alist)
    k = 16200 - 9000
    msgs, cws = helpers.encoded_frames(enc, rng, k, 16200, 96)
    for ebn0 in (0.6, 1.4):
        llrs = helpers.awgn_llrs(rng, cws, helpers.sigma_for(ebn0, k / 16200))
        compare(oracle, alist, "Minstarapproxi8", llrs, 25, out_len=k, label=f"ebn0 {ebn0} ")


def test_dvbs2_normal_r12_north_star(oracle):
    """BASELINE.json config 3: DVB-S2 n=64800 r=1/2, Minstarapproxi8, 25 iterations."""
    alist = codes.alist_for("dvbs2:R1_2")
    rng = np.random.default_rng(10)
    enc = oracle.encoder(alist)
    msgs, cws = helpers.encoded_frames(enc, rng, 32400, 64800, 160)
    llrs = np.concatenate([helpers.awgn_llrs(rng, cws[i * 40:(i + 1) * 40], helpers.sigma_for(e, 0.5))
                           for i, e in enumerate((0.5, 1.1, 1.25, 2.5))])
    its = compare(oracle, alist, "Minstarapproxi8", llrs, 25, out_len=32400)
    assert (its == -1).any() and (its > 0).any()


@pytest.mark.parametrize("cluster", ["1", "2", "8", "16"])
@pytest.mark.parametrize("impl", ["Minstarapproxi8", "Aminstari8JonesPartialHardLimitDeg1Clip"])
def test_cluster_sizes(oracle, impl, cluster, monkeypatch):
    """Small batches run one thread-block cluster per tile (checks / variables split over its CTAs, syndrome
    words OR-ed through distributed shared memory); every cluster size must give the single-CTA answer."""
    monkeypatch.setenv("LDPC_B200_CLUSTER", cluster)
    rng = np.random.default_rng(77)
    alist = helpers.random_code_alist(rng, 200, 80, col_w=[1, 2, 3, 4, 9], extra_heavy_rows=2)
    cws = np.zeros((300, 200), dtype=np.uint8)
    llrs = np.concatenate([helpers.awgn_llrs(rng, cws[:100], s) for s in (0.4, 0.8, 1.3)])
    its = compare(oracle, alist, impl, llrs, 15, label=f"cluster {cluster} ")
    assert (its > 0).any()
    # and on a real code with early termination spread over many iterations
    alist = codes.alist_for("dvbs2:R1_2short")
    n, k = 16200, 7200
    enc = oracle.encoder(alist)
    msgs, cw = helpers.encoded_frames(enc, rng, k, n, 150)
    compare(oracle, alist, impl, helpers.awgn_llrs(rng, cw, helpers.sigma_for(1.3, k / n)), 25, label=f"cluster {cluster} short ")


def test_straggler_redecode_is_exact(oracle, monkeypatch):
    """Two-stage chunks (decoder.cu): after the first call has filled the iteration histogram, frames that
    do not converge within m1 iterations are gathered and decoded again from their LLRs.  Words and iteration
    counts must be those of a single pass (and of the CPU checker)."""
    alist = codes.alist_for("dvbs2:R1_2short")
    n, k = 16200, 7200
    rng = np.random.default_rng(91)
    enc = oracle.encoder(alist)
    msgs, cw = helpers.encoded_frames(enc, rng, k, n, 64)
    llrs = helpers.awgn_llrs(rng, cw[np.arange(2048) % 64], helpers.sigma_for(1.55, k / n))
    monkeypatch.setenv("LDPC_B200_TWO_STAGE", "1")                  # opt-in
    dec = Decoder(alist, "Minstarapproxi8")
    out1, it1 = dec.decode_batch(llrs, 30, output_len=k)            # no histogram yet: one pass
    l1 = dec.last_timing()["kernel_launches"]
    out2, it2 = dec.decode_batch(llrs, 30, output_len=k)            # two stages
    l2 = dec.last_timing()["kernel_launches"] - l1
    assert l2 > l1, "the second call was expected to run in two stages"
    assert (it1 == it2).all() and (out1 == out2).all()
    assert (it1 == -1).any() and (it1 > 0).any()
    rout, rits = oracle.decoder(alist, "Minstarapproxi8").decode_batch(llrs[:512], 30, out_len=k)
    assert (rits == it2[:512]).all() and (rout == out2[:512]).all()
    monkeypatch.setenv("LDPC_B200_TWO_STAGE", "0")
    ref = Decoder(alist, "Minstarapproxi8")
    ref.decode_batch(llrs, 30, output_len=k)
    out3, it3 = ref.decode_batch(llrs, 30, output_len=k)
    assert (it3 == it2).all() and (out3 == out2).all()
