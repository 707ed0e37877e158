"""ctypes binding of include/ldpc_toolbox.h (the C-ABI of libldpc_toolbox.so).

This is the binding an out-of-tree user of the reference's libldpc_toolbox.so would write
(reference include/ldpc_toolbox.h:11-30); the extra batched / device-pointer entry points are
additive.  Loading fails loudly when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_LIB = None

_SIGS = {
    # name: (restype, argtypes)
    "ldpc_toolbox_decoder_ctor": (C.c_void_p, [C.c_char_p, C.c_char_p, C.c_char_p]),
    "ldpc_toolbox_decoder_ctor_alist_string": (C.c_void_p, [C.c_char_p, C.c_char_p, C.c_char_p]),
    "ldpc_toolbox_decoder_dtor": (None, [C.c_void_p]),
    "ldpc_toolbox_decoder_decode_f64": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32]),
    "ldpc_toolbox_decoder_decode_f32": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32]),
    "ldpc_toolbox_encoder_ctor": (C.c_void_p, [C.c_char_p, C.c_char_p]),
    "ldpc_toolbox_encoder_ctor_alist_string": (C.c_void_p, [C.c_char_p, C.c_char_p]),
    "ldpc_toolbox_encoder_dtor": (None, [C.c_void_p]),
    "ldpc_toolbox_encoder_encode": (None, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "ldpc_toolbox_last_error": (C.c_char_p, []),
    "ldpc_toolbox_decoder_ctor_ex": (C.c_void_p, [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int]),
    "ldpc_toolbox_decoder_decode_batch_f32": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p]),
    "ldpc_toolbox_decoder_decode_batch_f64": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p]),
    "ldpc_toolbox_decoder_submit_batch_f32": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p]),
    "ldpc_toolbox_decoder_submit_batch_f64": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p]),
    "ldpc_toolbox_decoder_decode_batch_posteriors_f32": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]),
    "ldpc_toolbox_decoder_decode_batch_posteriors_f64": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]),
    "ldpc_toolbox_decoder_wait": (C.c_int32, [C.c_void_p, C.c_int64]),
    "ldpc_toolbox_decoder_decode_batch_device_f32": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]),
    "ldpc_toolbox_decoder_decode_batch_device_f64": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]),
    "ldpc_toolbox_decoder_codeword_len": (C.c_size_t, [C.c_void_p]),
    "ldpc_toolbox_decoder_info_len": (C.c_size_t, [C.c_void_p]),
    "ldpc_toolbox_decoder_num_edges": (C.c_size_t, [C.c_void_p]),
    "ldpc_toolbox_decoder_llrs_len": (C.c_size_t, [C.c_void_p]),
    "ldpc_toolbox_decoder_last_timing": (C.c_int64, [C.c_void_p, C.c_void_p]),
    "ldpc_toolbox_decoder_average_decode_ms": (C.c_float, [C.c_void_p, C.c_void_p]),
    "ldpc_toolbox_ber_ctor": (C.c_void_p, [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int]),
    "ldpc_toolbox_ber_dtor": (None, [C.c_void_p]),
    "ldpc_toolbox_ber_set_modulation": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int32]),
    "ldpc_toolbox_ber_run": (C.c_int32, [C.c_void_p, C.c_float, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "ldpc_toolbox_ber_submit": (C.c_int64, [C.c_void_p, C.c_float, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "ldpc_toolbox_ber_wait": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p]),
    "ldpc_toolbox_ber_run_dump": (C.c_int32, [C.c_void_p, C.c_float, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldpc_toolbox_ber_dims": (None, [C.c_void_p, C.c_void_p]),
    "ldpc_toolbox_ber_rate": (C.c_double, [C.c_void_p]),
    "ldpc_toolbox_ber_noise_sigma": (C.c_double, [C.c_void_p, C.c_float]),
    "ldpc_toolbox_num_implementations": (C.c_int32, []),
    "ldpc_toolbox_implementation_name": (C.c_char_p, [C.c_int32]),
}

EXPORTED_SYMBOLS = tuple(_SIGS.keys())


def library_path() -> str:
    override = os.environ.get("LDPC_B200_LIB")     # experiment builds of the same CUDA library (tools/build_variant.py)
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libldpc_toolbox.so")


def load(build_if_missing: bool = True) -> C.CDLL:
    """Loads (building first if needed) libldpc_toolbox.so and declares every prototype."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise OSError(f"{path} is missing: run `python -m ldpc_toolbox_b200.build` (no CPU fallback exists)")
        from . import build as _build
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)      # AttributeError here = the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def last_error() -> str:
    return (load().ldpc_toolbox_last_error() or b"").decode()
