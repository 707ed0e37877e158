"""Committed golden vectors (tests/golden/golden_small.npz, made by tests/golden/make_golden.py from the
CPU checker): the checker must still reproduce them (CPU), and the GPU decoders must match them
(bit-exact for int8, within a stated handful of frames for the float rules)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_small.npz"))
MAX_ITER = int(G["max_iter"])


def cases():
    for tag in ("johnson", "random"):
        alist = str(G[f"alist_{tag}"])
        llrs = G[f"llrs_{tag}"]
        n = llrs.shape[1]
        for key in G.files:
            if key.startswith(f"bits_{tag}_"):
                impl = key[len(f"bits_{tag}_"):]
                bits = np.unpackbits(G[key], axis=1)[:, :n]
                yield tag, impl, alist, llrs, bits, G[f"its_{tag}_{impl}"].astype(np.int32)


CASES = list(cases())


def test_golden_file_shape():
    assert len(CASES) == 72
    assert G["llrs_random"].shape == (48, 96)


@pytest.mark.parametrize("tag,impl,alist,llrs,bits,its", CASES, ids=[f"{c[0]}-{c[1]}" for c in CASES])
def test_checker_reproduces_golden(oracle, tag, impl, alist, llrs, bits, its):
    out, got = oracle.decoder(alist, impl).decode_batch(llrs, MAX_ITER, nthreads=1)
    assert (got == its).all() and (out == bits).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,impl,alist,llrs,bits,its", CASES, ids=[f"{c[0]}-{c[1]}" for c in CASES])
def test_gpu_matches_golden(tag, impl, alist, llrs, bits, its):
    from ldpc_toolbox_b200 import Decoder
    out, got = Decoder(alist, impl).decode_batch(llrs, MAX_ITER)
    bad = int(((got != its) | (out != bits).any(axis=1)).sum())
    assert bad <= (0 if "i8" in impl else 1), f"{impl}: {bad} of {len(its)} frames differ from the golden vectors"


HL_CASES = [c for c in CASES if c[1].startswith("HL")]


@pytest.mark.gpu
@pytest.mark.parametrize("tag,impl,alist,llrs,bits,its", HL_CASES, ids=[f"{c[0]}-{c[1]}" for c in HL_CASES])
def test_gpu_frame_per_cta_layered_matches_golden(tag, impl, alist, llrs, bits, its, monkeypatch):
    """The same fixtures through K3q (frame per CTA, posteriors in shared memory), forced on these small codes."""
    from ldpc_toolbox_b200 import Decoder
    monkeypatch.setenv("LDPC_B200_LAYERED", "smem")
    out, got = Decoder(alist, impl).decode_batch(llrs, MAX_ITER)
    bad = int(((got != its) | (out != bits).any(axis=1)).sum())
    assert bad <= (0 if "i8" in impl else 1), f"{impl}: {bad} of {len(its)} frames differ from the golden vectors"
