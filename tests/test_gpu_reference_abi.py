"""The stock user's path: a C program linked against the reference's libldpc_toolbox.{so,a} sees exactly the nine
prototypes of the reference header (reference include/ldpc_toolbox.h:11-30) and nothing else.  This file binds ONLY
those nine symbols (its own ctypes declarations, not ldpc_toolbox_b200.capi) and walks the whole life cycle:
file-path constructor -> decode_f32 / decode_f64 -> destructor, encoder constructor -> encode -> destructor, and the
NULL returns of every constructor error (reference src/c_api/decoder.rs:76-137, src/c_api/encoder.rs:54-97).
Needs a B200 (constructors return NULL without a CUDA device: there is no CPU fallback)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from ldpc_toolbox_b200 import capi, codes

pytestmark = pytest.mark.gpu

JOHNSON = "6 4\n2 3\n2 2 2 2 2 2\n3 3 3 3\n1 3\n1 2\n2 4\n1 4\n2 3\n3 4\n1 2 4\n2 3 5\n1 5 6\n3 4 6\n"


@pytest.fixture(scope="module")
def ref_abi():
    """The nine reference prototypes, verbatim: names, argument order and C types."""
    lib = C.CDLL(capi.library_path())
    u8p, f64p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.POINTER(C.c_float)
    sigs = {
        "ldpc_toolbox_decoder_ctor": (C.c_void_p, [C.c_char_p, C.c_char_p, C.c_char_p]),
        "ldpc_toolbox_decoder_ctor_alist_string": (C.c_void_p, [C.c_char_p, C.c_char_p, C.c_char_p]),
        "ldpc_toolbox_decoder_dtor": (None, [C.c_void_p]),
        "ldpc_toolbox_decoder_decode_f64": (C.c_int32, [C.c_void_p, u8p, C.c_size_t, f64p, C.c_size_t, C.c_uint32]),
        "ldpc_toolbox_decoder_decode_f32": (C.c_int32, [C.c_void_p, u8p, C.c_size_t, f32p, C.c_size_t, C.c_uint32]),
        "ldpc_toolbox_encoder_ctor": (C.c_void_p, [C.c_char_p, C.c_char_p]),
        "ldpc_toolbox_encoder_ctor_alist_string": (C.c_void_p, [C.c_char_p, C.c_char_p]),
        "ldpc_toolbox_encoder_dtor": (None, [C.c_void_p]),
        "ldpc_toolbox_encoder_encode": (None, [C.c_void_p, u8p, C.c_size_t, u8p, C.c_size_t]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def _decode(lib, h, llrs, max_iter, out_len):
    out = np.full(out_len, 9, dtype=np.uint8)
    if llrs.dtype == np.float32:
        rc = lib.ldpc_toolbox_decoder_decode_f32(h, out.ctypes.data_as(C.POINTER(C.c_uint8)), out_len,
                                                 llrs.ctypes.data_as(C.POINTER(C.c_float)), llrs.size, max_iter)
    else:
        rc = lib.ldpc_toolbox_decoder_decode_f64(h, out.ctypes.data_as(C.POINTER(C.c_uint8)), out_len,
                                                 llrs.ctypes.data_as(C.POINTER(C.c_double)), llrs.size, max_iter)
    return out, rc


def test_reference_flooding_kat_through_file_path_ctor(ref_abi, tmp_path):
    """reference src/decoder/flooding.rs:161-189 (Phif64 on the 4x6 code), through the file-path constructor."""
    p = tmp_path / "johnson.alist"
    p.write_text(JOHNSON)
    h = ref_abi.ldpc_toolbox_decoder_ctor(str(p).encode(), b"Phif64", b"")
    assert h
    cw = np.array([0, 0, 1, 0, 1, 1], dtype=np.uint8)
    clean = np.where(cw == 1, -1.3863, 1.3863)
    out, it = _decode(ref_abi, h, clean, 100, 6)
    assert it == 0 and (out == cw).all()
    for j in range(6):
        llr = clean.copy()
        llr[j] *= -1
        out, it = _decode(ref_abi, h, llr, 100, 6)
        assert it == 1 and (out == cw).all()
        out, it = _decode(ref_abi, h, llr.astype(np.float32), 100, 4)      # output_len < n: the first bits (c_api/decoder.rs:61)
        assert it == 1 and (out == cw[:4]).all()
    ref_abi.ldpc_toolbox_decoder_dtor(h)


def test_stock_path_on_a_real_code(ref_abi, oracle, tmp_path):
    """File-path ctor + puncturing + decode_f32 / decode_f64 frame by frame, against the checker; then the encoder."""
    alist = codes.alist_for("ar4ja:1/2:1024")
    p = tmp_path / "ar4ja.alist"
    p.write_text(alist)
    rng = np.random.default_rng(21)
    eh = ref_abi.ldpc_toolbox_encoder_ctor(str(p).encode(), b"1,1,1,1,0")
    assert eh
    oenc = oracle.encoder(alist, "1,1,1,1,0")
    msgs = rng.integers(0, 2, size=(6, 1024), dtype=np.uint8)
    tx = np.zeros((6, 2048), dtype=np.uint8)
    for i, m in enumerate(msgs):
        ref_abi.ldpc_toolbox_encoder_encode(eh, tx[i].ctypes.data_as(C.POINTER(C.c_uint8)), 2048, m.ctypes.data_as(C.POINTER(C.c_uint8)), 1024)
        assert (tx[i] == oenc.encode(m, 2048)).all()
    ref_abi.ldpc_toolbox_encoder_dtor(eh)
    for impl in ("Minstarapproxi8", "HLAminstari8", "Phif64"):
        h = ref_abi.ldpc_toolbox_decoder_ctor(str(p).encode(), impl.encode(), b"1,1,1,1,0")
        assert h, impl
        odec = oracle.decoder(alist, impl, "1,1,1,1,0")
        for i in range(6):
            for dtype, ebn0 in ((np.float32, 2.2), (np.float64, 0.3)):
                llrs = helpers.awgn_llrs(rng, tx[i], helpers.sigma_for(ebn0, 0.5), dtype)
                out, it = _decode(ref_abi, h, llrs, 30, 1024)
                rout, rit = odec.decode(llrs, 30, out_len=1024)
                assert it == rit and (out == rout).all(), (impl, i, dtype)
        ref_abi.ldpc_toolbox_decoder_dtor(h)


def test_constructor_errors_return_null(ref_abi, tmp_path):
    p = tmp_path / "johnson.alist"
    p.write_text(JOHNSON)
    path = str(p).encode()
    assert not ref_abi.ldpc_toolbox_decoder_ctor(b"/nonexistent/file.alist", b"Phif64", b"")          # unreadable file
    assert not ref_abi.ldpc_toolbox_decoder_ctor(path, b"phif64", b"")                                # names are case-sensitive
    assert not ref_abi.ldpc_toolbox_decoder_ctor(path, b"Minstarapproxi9", b"")
    assert not ref_abi.ldpc_toolbox_decoder_ctor(path, b"Phif64", b"1,x,0")                           # bad puncturing pattern
    assert not ref_abi.ldpc_toolbox_decoder_ctor_alist_string(b"not an alist", b"Phif64", b"")
    assert not ref_abi.ldpc_toolbox_encoder_ctor(b"/nonexistent/file.alist", b"")
    assert not ref_abi.ldpc_toolbox_encoder_ctor_alist_string(b"3 2\n", b"")
    h = ref_abi.ldpc_toolbox_decoder_ctor_alist_string(JOHNSON.encode(), b"Tanhf32", b"")
    assert h
    ref_abi.ldpc_toolbox_decoder_dtor(h)
    # the 4x6 matrix has a redundant row: no systematic encoder exists (reference src/encoder.rs:59-97 returns Err)
    assert not ref_abi.ldpc_toolbox_encoder_ctor_alist_string(JOHNSON.encode(), b"")
    eh = ref_abi.ldpc_toolbox_encoder_ctor_alist_string(codes.alist_for("dvbs2:R1_2short").encode(), b"")
    assert eh
    ref_abi.ldpc_toolbox_encoder_dtor(eh)
