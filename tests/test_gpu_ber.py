"""GPU tests of the on-device BER engine (K4/K5) through the C-ABI: the frames it generates are
valid noisy codewords of the code, its counters equal a recount from dumped data with the CPU
checker decoding the very same LLRs, results are invariant to sharding, the noise is Gaussian with
the reference's sigma, and BER/FER agree statistically with the checker's own BER loop."""
import math

import numpy as np
import pytest

from ldpc_toolbox_b200 import codes
from ldpc_toolbox_b200.ber import BerEngine, BerTest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("spec,impl,punct,ebn0", [
    ("dvbs2:R1_2short", "Minstarapproxi8", "", 1.0),          # staircase encoder
    ("ar4ja:1/2:1024", "Phif64", "1,1,1,1,0", 1.5),            # dense encoder + puncturing
    ("nr5g:2:24", "HLAminstari8", "", 2.0),                    # dense encoder, layered decoder
])
def test_dump_consistency(oracle, spec, impl, punct, ebn0):
    alist = codes.alist_for(spec)
    eng = BerEngine(alist, impl, punct)
    nframes, max_iter = 300, 20
    counters, llrs, dec, its, msg = eng.run_dump(ebn0, max_iter, first_frame=1000, nframes=nframes)
    # (1) the transmitted word is the systematic encoding of the message: noiseless part of the LLRs
    enc = oracle.encoder(alist, punct)
    sigma = eng.noise_sigma(ebn0)
    tx = np.stack([enc.encode(m, eng.n) for m in msg])
    y = llrs / (-2.0 / sigma**2)
    noise = y - np.where(tx == 1, 1.0, -1.0)
    assert abs(noise.mean()) < 5 * sigma / math.sqrt(noise.size)
    assert abs(noise.std() / sigma - 1.0) < 0.01
    assert abs((noise**4).mean() / sigma**4 - 3.0) < 0.1           # Gaussian kurtosis
    assert abs(sigma - oracle.lib.ldpc_oracle_noise_sigma(eng.rate, 1.0, ebn0)) < 1e-12
    # (2) the engine's decode of those LLRs = the checker's decode of the same LLRs
    ref = oracle.decoder(alist, impl, punct)
    rout, rits = ref.decode_batch(llrs, max_iter, out_len=eng.k)
    if "i8" in impl:
        assert (rits == its).all() and (rout == dec).all()
    else:
        assert ((rits != its) | (rout != dec).any(axis=1)).sum() <= 1
    # (3) counters = recount (ber.rs:313-337)
    be = (dec != msg).sum(axis=1)
    iters = np.where(its < 0, max_iter, its)
    assert counters["frames"] == nframes
    assert counters["bit_errors"] == be.sum() and counters["frame_errors"] == (be > 0).sum()
    assert counters["false_decodes"] == ((be > 0) & (its >= 0)).sum()
    assert counters["total_iterations"] == iters.sum() and counters["correct_iterations"] == iters[be == 0].sum()
    assert counters["bch_frame_errors"] == (be > 0).sum()            # bch_max_errors = 0
    # messages look random
    assert abs(msg.mean() - 0.5) < 5 * 0.5 / math.sqrt(msg.size)


@pytest.mark.parametrize("bch_max_errors", [1, 3])
def test_bch_threshold_counters(bch_max_errors):
    """reference src/simulation/ber.rs:328-337: a frame with at most bch_max_errors info-bit errors counts as
    corrected by the outer BCH code (its iterations go to bch.correct_iterations), the others add their bit
    errors and one frame error.  DVB-S2 short r=1/2 near the waterfall leaves many frames with 1-3 bit errors."""
    eng = BerEngine(codes.alist_for("dvbs2:R1_2short"), "Minstarapproxi8")
    nframes, max_iter = 3000, 12
    counters, llrs, dec, its, msg = eng.run_dump(1.45, max_iter, first_frame=500, nframes=nframes, bch_max_errors=bch_max_errors)
    be = (dec != msg).sum(axis=1)
    iters = np.where(its < 0, max_iter, its)
    hard = be > bch_max_errors
    assert ((be > 0) & ~hard).sum() >= 5, "the point must produce frames the BCH threshold corrects"
    assert hard.sum() >= 5
    assert counters["bch_bit_errors"] == be[hard].sum()
    assert counters["bch_frame_errors"] == hard.sum()
    assert counters["bch_correct_iterations"] == iters[~hard].sum()
    assert counters["frame_errors"] == (be > 0).sum() and counters["bit_errors"] == be.sum()
    assert counters["correct_iterations"] == iters[be == 0].sum()
    # the driver stops on the BCH frame errors when the threshold is active (ber.rs:514-520)
    t = BerTest([eng], eng.k, [1.45], max_iterations=max_iter, max_frame_errors=40, bch_max_errors=bch_max_errors, batch=1024)
    st = t.run()[0]
    assert st.bch is not None and st.bch.frame_errors >= 40 and st.ldpc.frame_errors > st.bch.frame_errors


def test_sharding_invariance_and_determinism():
    alist = codes.alist_for("dvbs2:R1_2short")
    eng = BerEngine(alist, "Minstarapproxi8")
    a = eng.run(1.0, 20, 0, 1000)
    b = eng.run(1.0, 20, 0, 400)
    eng.run(1.0, 20, 400, 600, counters=b)
    assert (a == b).all()
    c = BerEngine(alist, "Minstarapproxi8").run(1.0, 20, 0, 1000)
    assert (a == c).all()
    d = eng.run(1.0, 20, 0, 1000, seed=1)
    assert (a != d).any()
    e = eng.run(1.1, 20, 0, 1000)
    assert (a != e).any()                                           # a different stream per Eb/N0
    assert a[2] > 0 and a[2] < 1000


def test_submit_wait_two_lanes():
    """Asynchronous form (ber.cu lanes): two batches in flight, waited in either order, add up to what the blocking
    call gives for the same global frames; a third submit without a wait is refused."""
    alist = codes.alist_for("dvbs2:R1_2short")
    eng = BerEngine(alist, "Minstarapproxi8")
    ref = eng.run(1.0, 20, 0, 1500)
    t0 = eng.submit(1.0, 20, 0, 700)
    t1 = eng.submit(1.0, 20, 700, 800)
    with pytest.raises(RuntimeError):
        eng.submit(1.0, 20, 1500, 10)
    c = eng.wait(t1)
    eng.wait(t0, c)
    assert (c == ref).all()
    with pytest.raises(RuntimeError):
        eng.wait(t0)
    # the pipelined driver (two rounds in flight) counts whole rounds and reproduces the blocking totals
    t = BerTest([eng], eng.k, [1.0], max_iterations=20, max_frame_errors=10**9, batch=500, max_frames=1500)
    assert t.pipeline_depth == 2 and t.overshoot_bound() == 500
    st = t.run()[0]
    assert (st.num_frames, st.ldpc.bit_errors, st.ldpc.frame_errors, st.total_iterations) == (int(ref[0]), int(ref[1]), int(ref[2]), int(ref[4]))


def test_ber_statistical_parity_with_checker(oracle):
    """FER of the GPU engine vs the CPU checker's BER loop on the same code/decoder/Eb/N0: each inside
    the other's 3-sigma binomial band (the reference's RNG is OS-seeded, so only statistics compare)."""
    alist = codes.alist_for("ar4ja:1/2:1024")
    eng = BerEngine(alist, "Minstarapproxi8", "1,1,1,1,0")
    for ebn0 in (1.0, 1.75):
        g = eng.run(ebn0, 50, 0, 20000)
        c = oracle.ber_run(alist, "Minstarapproxi8", "1,1,1,1,0", ebn0, 50, frames=4000, nthreads=0 or 8, seed=3)
        pg, pc = g[2] / g[0], c["frame_errors"] / c["frames"]
        s = math.sqrt(pc * (1 - pc) / c["frames"] + pg * (1 - pg) / g[0]) + 1e-9
        assert abs(pg - pc) < 4 * s, (ebn0, pg, pc)
        ig, ic = g[4] / g[0], c["total_iterations"] / c["frames"]
        assert abs(ig - ic) < 0.05 * ic + 0.5, (ebn0, ig, ic)


def test_ber_sweep_driver():
    alist = codes.alist_for("dvbs2:R1_2short")
    eng = BerEngine(alist, "Minstarapproxi8")
    seen = []
    t = BerTest([eng], eng.k, [0.6, 1.6], max_iterations=25, max_frame_errors=50, batch=2048, reporter=lambda st, fin: seen.append((st.ebn0_db, fin)))
    stats = t.run()
    assert [round(s.ebn0_db, 2) for s in stats] == [0.6, 1.6]
    assert stats[0].ldpc.fer > 0.5 and stats[0].ldpc.frame_errors >= 50
    assert stats[0].ldpc.ber > stats[0].ldpc.fer / eng.k
    assert any(fin for _, fin in seen)
    assert stats[0].throughput_mbps > 0


def test_cli_ber_output(tmp_path, capsys):
    from ldpc_toolbox_b200 import cli
    p = tmp_path / "code.alist"
    p.write_text(codes.alist_for("dvbs2:R1_2short"))
    out = tmp_path / "out.txt"
    cli.main(["ber", str(p), "--decoder", "Minstarapproxi8", "--min-ebn0", "0.5", "--max-ebn0", "0.75", "--step-ebn0", "0.25",
              "--max-iter", "20", "--frame-errors", "20", "--batch", "1024", "--output-file", str(out)])
    text = capsys.readouterr().out
    assert "BER TEST PARAMETERS" in text and " - Information bits (k): 7200" in text and " - Code rate: 0.444" in text
    assert " - Implementation: Minstarapproxi8" in text
    rows = [l for l in out.read_text().split("\n") if l.startswith("   0.")]
    assert len(rows) == 2 and rows[0].startswith("   0.50 |") and rows[1].startswith("   0.75 |")
    assert rows[0].count("|") == 10


@pytest.mark.parametrize("interleaving", [None, 3, -3])
def test_psk8_link_noiseless_limit(oracle, interleaving):
    """8PSK + DVB-S2 bit interleaver on the device (reference modulation.rs:144-288, interleaving.rs:40-85):
    at a very high Eb/N0 the dumped LLRs must be the max* demapping of the noiseless symbols of the
    interleaved codeword, put back in codeword order."""
    import pyref
    alist = codes.alist_for("dvbs2:R1_2short")
    eng = BerEngine(alist, "Minstarapproxi8", modulation="8PSK", interleaving=interleaving)
    ebn0 = 50.0
    counters, llrs, dec, its, msg = eng.run_dump(ebn0, 5, first_frame=7, nframes=6)
    sigma = eng.noise_sigma(ebn0)
    # Es/N0 uses 3 bits per symbol (ber.rs:300-302, modulation.rs:151)
    assert abs(sigma - oracle.lib.ldpc_oracle_noise_sigma(eng.rate, 3.0, ebn0)) < 1e-12
    enc = oracle.encoder(alist)
    cols, back = (abs(interleaving), interleaving < 0) if interleaving else (0, False)
    for f in range(msg.shape[0]):
        cw = enc.encode(msg[f], eng.n)
        tx = pyref.interleave(cw, cols, back) if cols else cw
        ref = np.array(pyref.psk8_demodulate(pyref.psk8_modulate(tx), sigma))
        ref = pyref.deinterleave(ref, cols, back) if cols else ref
        assert ((llrs[f] <= 0) == (cw == 1)).all()
        assert np.allclose(llrs[f], ref, rtol=0.05, atol=0.0)          # noise of sigma = 0.002 on unit symbols
    assert counters["bit_errors"] == 0 and (dec == msg).all()


def test_psk8_waterfall_sits_above_bpsk():
    """Same code and decoder: 8PSK needs more Eb/N0 than BPSK, and decodes cleanly once it has it."""
    alist = codes.alist_for("dvbs2:R1_2short")
    bpsk = BerEngine(alist, "Minstarapproxi8")
    psk8 = BerEngine(alist, "Minstarapproxi8", modulation="8PSK", interleaving=3)
    fer = lambda eng, e: (lambda c: c[2] / c[0])(eng.run(e, 30, 0, 512))
    assert fer(bpsk, 1.6) < 0.05
    assert fer(psk8, 1.6) > 0.5
    assert fer(psk8, 4.5) < 0.05


def test_psk8_rejects_bad_lengths():
    with pytest.raises(ValueError):
        BerEngine(codes.alist_for("ar4ja:1/2:1024"), "Phif64", modulation="8PSK")       # 2560 bits: not a multiple of 3
    with pytest.raises(ValueError):
        BerEngine(codes.alist_for("dvbs2:R1_2short"), "Phif64", modulation="QPSK")


def test_full_size_tile_shape_invariance(monkeypatch):
    """BASELINE.json configs[2] at a launch-sized batch (75 776 frames of DVB-S2 n=64800 r=1/2, Minstarapproxi8,
    25 iterations, waterfall point): the 512-frame-tile kernel and the 128-frame-tile kernel — the shape
    the oracle parity tests pin at small sizes — must return identical counters for the same global frames
    (bit errors, frame errors, iteration sums: a checksum over every decoded word and iteration count)."""
    import torch
    alist = codes.cached_alist_path("dvbs2:R1_2")
    frames = torch.cuda.get_device_properties(0).multi_processor_count * 512
    big = BerEngine(alist, "Minstarapproxi8")
    a = big.run(1.15, 25, 0, frames)
    big.close()
    monkeypatch.setenv("LDPC_B200_NW", "1")
    small = BerEngine(alist, "Minstarapproxi8")
    b = small.run(1.15, 25, 0, frames)
    small.close()
    assert a.tolist() == b.tolist()
    assert a[0] == frames and 0 < a[2] < frames                   # some frames fail, some converge
    assert a[4] < 25 * frames                                     # early termination happened
