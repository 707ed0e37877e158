"""Posterior LLRs of the float decoders against the CPU checker (BASELINE.json north_star: "output LLRs within a
stated relative tolerance"), through the test hook ldpc_toolbox_decoder_decode_batch_posteriors_* — the flooding
decoder's output_llrs (reference src/decoder/flooding.rs:111-125) and the layered decoder's Qv
(src/decoder/horizontal_layered.rs:65-88) at the moment each frame stopped.

Stated tolerances, relative to max(|reference|, 1), on frames whose word and iteration count match:
  Min*-approx, A-Min*:  f64 every value within 1e-9; f32 99.9 % within 1e-4 and all within 1e-2 (a frame that needs
                        many iterations amplifies last-ulp differences: worst value seen 2.9e-4)   [SURVEY.md §8c proposal]
  Tanh:                 f64 every value within 1e-9; f32 99.9 % within 1e-4 and all within 1e-2 — 2 atanh(prod) is
                        ill-conditioned where the product of tanh values rounds towards 1 in f32
  Phi:                  99 % within 1e-4 (f32) / 1e-12 (f64) and no bound on the rest: phi(sum - phi_j) cancels
                        catastrophically when one input dominates the sum, so one ulp of difference in a libm result
                        moves a few isolated messages by O(1) in BOTH implementations' own arithmetic (the words still
                        agree).  (The f32 rule has since been given bit-exact ports of glibc's tanhf / logf, rules.cuh.)
Measured on a B200 (tests/posterior_probe.py): 256 of 256 frames match for all 16 names; maxima 1.4e-5 / 5e-14
(Min*-approx), 8.6e-6 / 1.3e-14 (A-Min*), 2.8e-3 / 1.7e-11 (Tanh)."""
import numpy as np
import pytest

import helpers
from ldpc_toolbox_b200 import Decoder, codes

pytestmark = pytest.mark.gpu

CASES = [("ar4ja:1/2:1024", "1,1,1,1,0", 1.8, 50, ""), ("nr5g:2:96", "", 1.2, 30, "HL")]
RULES = ["Phi", "Tanh", "Minstarapprox", "Aminstar"]


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("rule", RULES)
@pytest.mark.parametrize("code,punct,ebn0,max_iter,prefix", CASES)
def test_posterior_llrs_within_tolerance(oracle, code, punct, ebn0, max_iter, prefix, rule, dtype):
    impl = f"{prefix}{rule}{dtype}"
    alist = codes.alist_for(code)
    n, m = (int(x) for x in alist.split("\n")[0].split())
    k = n - m
    rng = np.random.default_rng(5)
    enc = oracle.encoder(alist, punct)
    n_tx = n * 4 // 5 if punct else n
    msgs = rng.integers(0, 2, size=(8, k), dtype=np.uint8)
    tx = np.stack([enc.encode(mm, n_tx) for mm in msgs])
    nframes = 96
    llrs = helpers.awgn_llrs(rng, tx[np.arange(nframes) % 8], helpers.sigma_for(ebn0, k / n_tx), np.float64 if dtype == "f64" else np.float32)
    out, its, post = Decoder(alist, impl, punct).decode_batch_posteriors(llrs, max_iter, output_len=k)
    ref = oracle.decoder(alist, impl, punct)
    rel = []
    for f in range(nframes):
        rout, rit = ref.decode(llrs[f], max_iter, out_len=k)
        if rit != its[f] or (rout != out[f]).any() or rit == 0:
            continue
        rp = ref.posteriors()
        assert ((post[f] <= 0) == (rp <= 0))[:k].all()            # same hard decisions as the returned word
        rel.append(np.abs(post[f] - rp) / np.maximum(np.abs(rp), 1.0))
    assert len(rel) >= nframes - 2, f"{impl}: only {len(rel)} of {nframes} frames agree in word and iteration count"
    rel = np.concatenate(rel)
    small = 1e-9 if dtype == "f64" else 1e-4
    if rule in ("Minstarapprox", "Aminstar", "Tanh"):
        if dtype == "f64":
            assert rel.max() <= small, f"{impl}: max relative error {rel.max():.3g}"
        else:
            assert np.quantile(rel, 0.999) <= 1e-4 and rel.max() <= 1e-2, f"{impl}: p99.9 {np.quantile(rel, 0.999):.3g} max {rel.max():.3g}"
    else:
        assert np.quantile(rel, 0.99) <= (1e-12 if dtype == "f64" else 1e-4), f"{impl}: p99 {np.quantile(rel, 0.99):.3g}"


def test_posteriors_refused_where_not_implemented():
    alist = codes.alist_for("ar4ja:1/2:1024")
    with pytest.raises(ValueError):
        Decoder(alist, "Minstarapproxi8", "1,1,1,1,0").decode_batch_posteriors(np.zeros((2, 2048), dtype=np.float32), 5)
