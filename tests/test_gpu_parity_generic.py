"""GPU parity for K2 (float flooding) and K3 (horizontal layered) against the CPU checker, through
the C-ABI.  int8 layered decoders: bit-exact words and iteration counts.  Float decoders: the
transcendental functions come from libdevice instead of the host libm (except the f32 Phi rule, which runs
bit-exact ports of glibc's tanhf / logf), so parity is statistical — the decoded word and iteration count
must agree on (nearly) every frame (BASELINE.json: >= 99.99 % of frames; at scale in test_gpu_parity_scale.py,
here at most ONE frame of each 800-frame sample, none for f64 and Phi).  Needs a B200."""
import numpy as np
import pytest

import helpers
from ldpc_toolbox_b200 import Decoder, codes, implementation_names

pytestmark = pytest.mark.gpu

NAMES = implementation_names()
FLOAT_FLOOD = [n for n in NAMES if not n.startswith("HL") and "i8" not in n]
HL_I8 = [n for n in NAMES if n.startswith("HL") and "i8" in n]
HL_FLOAT = [n for n in NAMES if n.startswith("HL") and "i8" not in n]


def run_pair(oracle, alist, impl, llrs, max_iter, puncturing="", out_len=None):
    dec = Decoder(alist, impl, puncturing)
    ref = oracle.decoder(alist, impl, puncturing)
    out, its = dec.decode_batch(llrs, max_iter, output_len=out_len)
    rout, rits = ref.decode_batch(llrs, max_iter, out_len=out_len)
    bad = (its != rits) | (out != rout).any(axis=1)
    return int(bad.sum()), its, rits


def small_code_stimulus(seed, dtype=np.float32):
    rng = np.random.default_rng(seed)
    cases = []
    for n, m, heavy in ((96, 48, 0), (200, 80, 2)):
        alist = helpers.random_code_alist(rng, n, m, col_w=[1, 2, 3, 4, 9], extra_heavy_rows=heavy)
        llrs = np.concatenate([helpers.awgn_llrs(rng, np.zeros((100, n), dtype=np.uint8), s, dtype) for s in (0.3, 0.6, 0.9, 1.4)])
        llrs[3] = 0.0
        cases.append((alist, llrs))
    return cases


@pytest.mark.parametrize("path", ["tile", "smem"])
@pytest.mark.parametrize("impl", HL_I8)
def test_layered_i8_bit_exact(oracle, impl, path, monkeypatch):
    """Both layered kernels: K3 (frame-interleaved tiles) and K3q (frame per CTA, posteriors in shared memory)."""
    monkeypatch.setenv("LDPC_B200_LAYERED", path)
    for alist, llrs in small_code_stimulus(21):
        nbad, its, _ = run_pair(oracle, alist, impl, llrs, 12)
        assert nbad == 0, impl
        assert (its > 0).any()


@pytest.mark.parametrize("impl", HL_FLOAT)
def test_layered_float_smem_path_small_codes(oracle, impl, monkeypatch):
    monkeypatch.setenv("LDPC_B200_LAYERED", "smem")
    total = bad = 0
    for alist, llrs in small_code_stimulus(23, np.float64 if impl.endswith("f64") else np.float32):
        nbad, its, rits = run_pair(oracle, alist, impl, llrs, 12)
        bad += nbad
        total += len(its)
    # measured: 0 of 800 for every name (tools/float_small_code_counts.py); one last-ulp flip is tolerated for the f32
    # rules that still use libdevice transcendentals, none for f64 and for the bit-exact f32 Phi rule
    assert bad <= (0 if impl.endswith("f64") or "Phi" in impl else 1), f"{impl}: {bad} of {total} frames differ from the CPU checker"


def test_layered_smem_punctured_f64_input_and_zero_iterations(oracle, monkeypatch):
    """K3q reads the caller's LLRs itself: depuncturing, f64 input, max_iterations = 0 and output_len < n."""
    monkeypatch.setenv("LDPC_B200_LAYERED", "smem")
    alist = codes.alist_for("ar4ja:1/2:1024")
    rng = np.random.default_rng(35)
    enc = oracle.encoder(alist, "1,1,1,1,0")
    msgs = rng.integers(0, 2, size=(96, 1024), dtype=np.uint8)
    tx = np.stack([enc.encode(m, 2048) for m in msgs])
    llrs = helpers.awgn_llrs(rng, tx, helpers.sigma_for(1.8, 0.5), np.float64)
    for impl, iters in (("HLMinstarapproxi8", 30), ("HLAminstari8", 0), ("HLPhif64", 20)):
        nbad, its, rits = run_pair(oracle, alist, impl, llrs, iters, puncturing="1,1,1,1,0", out_len=1024)
        assert nbad <= (1 if impl.endswith("f64") else 0), (impl, nbad)


@pytest.mark.parametrize("impl", FLOAT_FLOOD + HL_FLOAT)
def test_float_rules_small_codes(oracle, impl):
    total = bad = 0
    for alist, llrs in small_code_stimulus(22, np.float64 if impl.endswith("f64") else np.float32):
        nbad, its, rits = run_pair(oracle, alist, impl, llrs, 12)
        bad += nbad
        total += len(its)
        assert (its > 0).any()
    # measured: 0 of 800 for every name (tools/float_small_code_counts.py).  f64: none tolerated; f32 rules on libdevice
    # transcendentals (Tanh, Min*-approx, A-Min*): one last-ulp flip; the f32 Phi rule runs bit-exact libm ports: none
    limit = 0 if impl.endswith("f64") or "Phi" in impl else 1
    assert bad <= limit, f"{impl}: {bad} of {total} frames differ from the CPU checker"


def test_johnson_phif64_reference_kat():
    """reference src/decoder/flooding.rs:161-189 through the GPU path."""
    johnson = "6 4\n2 3\n2 2 2 2 2 2\n3 3 3 3\n1 3\n1 2\n2 4\n1 4\n2 3\n3 4\n1 2 4\n2 3 5\n1 5 6\n3 4 6\n"
    dec = Decoder(johnson, "Phif64")
    good = [0, 0, 1, 0, 1, 1]
    to_llrs = lambda bits: np.array([1.3863 if b == 0 else -1.3863 for b in bits])
    out, it = dec.decode(to_llrs(good), 100)
    assert it == 0 and out.tolist() == good
    for j in range(6):
        bad = list(good)
        bad[j] ^= 1
        out, it = dec.decode(to_llrs(bad), 100)
        assert it == 1 and out.tolist() == good


def test_a10_layered_vectors():
    johnson = "6 4\n2 3\n2 2 2 2 2 2\n3 3 3 3\n1 3\n1 2\n2 4\n1 4\n2 3\n3 4\n1 2 4\n2 3 5\n1 5 6\n3 4 6\n"
    llr = np.array([-1.3863, 1.3863, -1.3863, 1.3863, -1.3863, -1.3863])
    for impl, iters in (("HLMinstarapproxi8", 1), ("HLAminstari8", 2)):
        out, it = Decoder(johnson, impl).decode(llr, 100)
        assert it == iters and out.tolist() == [0, 0, 1, 0, 1, 1]


def test_config1_ar4ja_phif64_punctured(oracle):
    """BASELINE.json configs[0]: CCSDS AR4JA r=1/2 k=1024, Phif64 flooding, puncturing 1,1,1,1,0."""
    alist = codes.alist_for("ar4ja:1/2:1024")
    rng = np.random.default_rng(31)
    enc = oracle.encoder(alist, "1,1,1,1,0")
    msgs = rng.integers(0, 2, size=(256, 1024), dtype=np.uint8)
    tx = np.stack([enc.encode(m, 2048) for m in msgs])
    llrs = np.concatenate([helpers.awgn_llrs(rng, tx[i * 64:(i + 1) * 64], helpers.sigma_for(e, 0.5), np.float64)
                           for i, e in enumerate((0.5, 1.5, 2.0, 3.0))])
    nbad, its, rits = run_pair(oracle, alist, "Phif64", llrs, 100, puncturing="1,1,1,1,0", out_len=1024)
    assert nbad <= 1
    assert (its == -1).any() and (its > 0).any()


def test_config2_nr_bg2_hl_minstarapprox_f32(oracle):
    """BASELINE.json configs[1] (reduced batch): 5G NR BG2 Z=384, HLMinstarapproxf32, 50 iterations."""
    alist = codes.alist_for("nr5g:2:384")
    n, k = 19968, 3840
    rng = np.random.default_rng(32)
    enc = oracle.encoder(alist)
    msgs, cws = helpers.encoded_frames(enc, rng, k, n, 192)
    llrs = np.concatenate([helpers.awgn_llrs(rng, cws[i * 64:(i + 1) * 64], helpers.sigma_for(e, k / n)) for i, e in enumerate((-1.0, 0.0, 1.0))])
    nbad, its, rits = run_pair(oracle, alist, "HLMinstarapproxf32", llrs, 50, out_len=k)
    assert nbad <= 2, nbad
    assert (its > 0).any()


def test_config4_nr_bg1_aminstar_f32_flooding_vs_layered(oracle):
    """BASELINE.json configs[3] (reduced): NR BG1 Z=384, Aminstarf32 flooding vs HLAminstarf32."""
    alist = codes.alist_for("nr5g:1:384")
    n, k = 26112, 8448
    rng = np.random.default_rng(33)
    enc = oracle.encoder(alist)
    msgs, cws = helpers.encoded_frames(enc, rng, k, n, 64)
    llrs = helpers.awgn_llrs(rng, cws, helpers.sigma_for(1.5, k / n))
    res = {}
    for impl in ("Aminstarf32", "HLAminstarf32"):
        nbad, its, rits = run_pair(oracle, alist, impl, llrs, 30, out_len=k)
        assert nbad <= 1, (impl, nbad)
        res[impl] = np.where(its < 0, 30, its).mean()
    assert res["HLAminstarf32"] < res["Aminstarf32"]      # layered converges in fewer iterations


def test_dvbs2_short_layered_i8(oracle):
    """Layered decoding of a staircase code: every row is chained to the next (one row per level)."""
    alist = codes.alist_for("dvbs2:R8_9short")
    n, k = 16200, 14400
    rng = np.random.default_rng(34)
    enc = oracle.encoder(alist)
    msgs, cws = helpers.encoded_frames(enc, rng, k, n, 40)
    llrs = helpers.awgn_llrs(rng, cws, helpers.sigma_for(4.2, k / n))
    nbad, its, _ = run_pair(oracle, alist, "HLMinstarapproxi8", llrs, 10, out_len=k)
    assert nbad == 0
