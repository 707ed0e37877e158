"""ctypes loader for the CPU checker (oracle/).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _cpu_tag():
    """The checker is built with -march=native; key the build directory by the host's ISA flags so a
    library built in the CPU container is never loaded on a GPU box with a different CPU."""
    import hashlib
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except Exception:
        flags = "unknown"
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


SO = os.path.join(ROOT, "oracle", "_build", _cpu_tag(), "libldpc_oracle.so")


def build(force=False):
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("ldpc_oracle.cpp", "ldpc_oracle_capi.cpp", "ldpc_oracle.hpp")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "OUT=" + SO], stdout=subprocess.DEVNULL)
    return SO


_u8p = C.POINTER(C.c_uint8)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.ldpc_oracle_decoder_ctor.restype = C.c_void_p
        L.ldpc_oracle_decoder_ctor.argtypes = [C.c_char_p] * 3
        L.ldpc_oracle_decoder_ctor_alist_string.restype = C.c_void_p
        L.ldpc_oracle_decoder_ctor_alist_string.argtypes = [C.c_char_p] * 3
        L.ldpc_oracle_decoder_dtor.argtypes = [C.c_void_p]
        L.ldpc_oracle_decoder_decode_f64.restype = C.c_int32
        L.ldpc_oracle_decoder_decode_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32]
        L.ldpc_oracle_decoder_decode_f32.restype = C.c_int32
        L.ldpc_oracle_decoder_decode_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32]
        for f in (L.ldpc_oracle_decoder_decode_batch_f32, L.ldpc_oracle_decoder_decode_batch_f64):
            f.restype = C.c_int32
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p, C.c_int]
        L.ldpc_oracle_decoder_set_linear_search.argtypes = [C.c_void_p, C.c_int]
        for f in (L.ldpc_oracle_decoder_n, L.ldpc_oracle_decoder_m, L.ldpc_oracle_decoder_edges):
            f.restype = C.c_size_t
            f.argtypes = [C.c_void_p]
        L.ldpc_oracle_decoder_posteriors.restype = C.c_size_t
        L.ldpc_oracle_decoder_posteriors.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ldpc_oracle_encoder_ctor.restype = C.c_void_p
        L.ldpc_oracle_encoder_ctor.argtypes = [C.c_char_p] * 2
        L.ldpc_oracle_encoder_ctor_alist_string.restype = C.c_void_p
        L.ldpc_oracle_encoder_ctor_alist_string.argtypes = [C.c_char_p] * 2
        L.ldpc_oracle_encoder_dtor.argtypes = [C.c_void_p]
        L.ldpc_oracle_encoder_is_staircase.argtypes = [C.c_void_p]
        L.ldpc_oracle_encoder_encode.restype = C.c_int32
        L.ldpc_oracle_encoder_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ldpc_oracle_implementation_name.restype = C.c_char_p
        L.ldpc_oracle_implementation_name.argtypes = [C.c_int]
        L.ldpc_oracle_alist_roundtrip.restype = C.c_size_t
        L.ldpc_oracle_alist_roundtrip.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.ldpc_oracle_ber_run.restype = C.c_int32
        L.ldpc_oracle_ber_run.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_float, C.c_uint32, C.c_uint64, C.c_uint64,
                                          C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.ldpc_oracle_noise_sigma.restype = C.c_double
        L.ldpc_oracle_noise_sigma.argtypes = [C.c_double, C.c_double, C.c_float]

    def implementations(self):
        return [self.lib.ldpc_oracle_implementation_name(i).decode() for i in range(self.lib.ldpc_oracle_num_implementations())]

    def decoder(self, alist_text, impl, puncturing=""):
        return OracleDecoder(self, alist_text, impl, puncturing)

    def encoder(self, alist_text, puncturing=""):
        return OracleEncoder(self, alist_text, puncturing)

    def alist_roundtrip(self, text, padding=True):
        need = self.lib.ldpc_oracle_alist_roundtrip(text.encode(), int(padding), None, 0)
        if need == 0:
            return None
        buf = C.create_string_buffer(need)
        self.lib.ldpc_oracle_alist_roundtrip(text.encode(), int(padding), buf, need)
        return buf.value.decode()

    def ber_run(self, alist_text, impl, puncturing, ebn0_db, max_iter, frames=0, max_frame_errors=100, nthreads=1, seed=1,
                linear_search=False):
        counters = (C.c_uint64 * 6)()
        el = C.c_double()
        rc = self.lib.ldpc_oracle_ber_run(alist_text.encode(), impl.encode(), puncturing.encode(), ebn0_db, max_iter, frames,
                                          max_frame_errors, nthreads, seed, int(linear_search), counters, C.byref(el))
        if rc != 0:
            raise RuntimeError("oracle ber_run failed")
        keys = ["frames", "bit_errors", "frame_errors", "false_decodes", "total_iterations", "correct_iterations"]
        d = dict(zip(keys, [int(x) for x in counters]))
        d["elapsed_s"] = el.value
        return d


class OracleDecoder:
    def __init__(self, o, alist_text, impl, puncturing=""):
        self.o = o
        self.h = o.lib.ldpc_oracle_decoder_ctor_alist_string(alist_text.encode(), impl.encode(), puncturing.encode())
        if not self.h:
            raise ValueError("oracle decoder ctor returned NULL")
        self.n = o.lib.ldpc_oracle_decoder_n(self.h)
        self.m = o.lib.ldpc_oracle_decoder_m(self.h)
        self.edges = o.lib.ldpc_oracle_decoder_edges(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.o.lib.ldpc_oracle_decoder_dtor(self.h)
            self.h = None

    def set_linear_search(self, on):
        self.o.lib.ldpc_oracle_decoder_set_linear_search(self.h, int(on))

    def decode(self, llrs, max_iter, out_len=None):
        llrs = np.ascontiguousarray(llrs)
        out_len = self.n if out_len is None else out_len
        out = np.zeros(out_len, dtype=np.uint8)
        if llrs.dtype == np.float32:
            rc = self.o.lib.ldpc_oracle_decoder_decode_f32(self.h, out.ctypes.data, out_len, llrs.ctypes.data, llrs.size, max_iter)
        else:
            llrs = llrs.astype(np.float64)
            rc = self.o.lib.ldpc_oracle_decoder_decode_f64(self.h, out.ctypes.data, out_len, llrs.ctypes.data, llrs.size, max_iter)
        return out, rc

    def posteriors(self):
        p = np.zeros(self.n, dtype=np.float64)
        self.o.lib.ldpc_oracle_decoder_posteriors(self.h, p.ctypes.data, p.size)
        return p

    def decode_batch(self, llrs, max_iter, out_len=None, nthreads=0):
        llrs = np.ascontiguousarray(llrs)
        nframes, per = llrs.shape
        out_len = self.n if out_len is None else out_len
        out = np.zeros((nframes, out_len), dtype=np.uint8)
        its = np.zeros(nframes, dtype=np.int32)
        fn = self.o.lib.ldpc_oracle_decoder_decode_batch_f32 if llrs.dtype == np.float32 else self.o.lib.ldpc_oracle_decoder_decode_batch_f64
        if llrs.dtype not in (np.float32, np.float64):
            llrs = llrs.astype(np.float64)
        fn(self.h, out.ctypes.data, out_len, llrs.ctypes.data, per, nframes, max_iter, its.ctypes.data, nthreads)
        return out, its


class OracleEncoder:
    def __init__(self, o, alist_text, puncturing=""):
        self.o = o
        self.h = o.lib.ldpc_oracle_encoder_ctor_alist_string(alist_text.encode(), puncturing.encode())
        if not self.h:
            raise ValueError("oracle encoder ctor returned NULL")

    def __del__(self):
        if getattr(self, "h", None):
            self.o.lib.ldpc_oracle_encoder_dtor(self.h)
            self.h = None

    @property
    def is_staircase(self):
        return bool(self.o.lib.ldpc_oracle_encoder_is_staircase(self.h))

    def encode(self, msg, out_len):
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        out = np.zeros(out_len, dtype=np.uint8)
        rc = self.o.lib.ldpc_oracle_encoder_encode(self.h, out.ctypes.data, out_len, msg.ctypes.data, msg.size)
        if rc != 0:
            raise ValueError("oracle encode: length mismatch")
        return out


_cached = None


def load():
    global _cached
    if _cached is None:
        _cached = Oracle(C.CDLL(build()))
    return _cached
