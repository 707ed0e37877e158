"""Host-side mirror of the reference's decoder interface on top of the C-ABI.

  DecoderImplementation   reference src/decoder/factory.rs:31-188 (the 36 names, verbatim)
  Decoder.decode          reference src/decoder.rs:19-35 via src/c_api/decoder.rs:50-72:
                          returns (hard bits, iterations) with iterations == -1 on failure
  Decoder.decode_batch    the same call for many frames at once (native shape of the GPU path)
  Encoder.encode          reference src/c_api/encoder.rs:37-52
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi


def implementation_names() -> list[str]:
    lib = capi.load()
    return [lib.ldpc_toolbox_implementation_name(i).decode() for i in range(lib.ldpc_toolbox_num_implementations())]


class DecoderImplementation(str):
    """One of the reference's implementation names, e.g. ``DecoderImplementation("Minstarapproxi8")``."""

    def __new__(cls, name: str):
        if name not in implementation_names():
            raise ValueError("invalid decoder implementation")     # factory.rs:219
        return super().__new__(cls, name)

    def build_decoder(self, alist: str, puncturing: str = "", **kw) -> "Decoder":   # DecoderFactory::build_decoder
        return Decoder(alist, str(self), puncturing, **kw)


def _alist_arg(alist: str):
    is_path = "\n" not in alist and os.path.exists(alist)
    return alist.encode(), int(is_path)


class Decoder:
    def __init__(self, alist: str, implementation: str = "Phif64", puncturing: str = "", device: int = -1, max_tiles: int = 0):
        self._lib = capi.load()
        text, is_path = _alist_arg(alist)
        self._h = self._lib.ldpc_toolbox_decoder_ctor_ex(text, is_path, implementation.encode(), puncturing.encode(), device, max_tiles)
        if not self._h:
            raise ValueError(f"ldpc_toolbox_decoder_ctor returned NULL: {capi.last_error()}")
        self.implementation = implementation
        self.n = self._lib.ldpc_toolbox_decoder_codeword_len(self._h)
        self.k = self._lib.ldpc_toolbox_decoder_info_len(self._h)
        self.num_edges = self._lib.ldpc_toolbox_decoder_num_edges(self._h)
        self.llrs_len = self._lib.ldpc_toolbox_decoder_llrs_len(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ldpc_toolbox_decoder_dtor(self._h)
            self._h = None

    __del__ = close

    def decode(self, llrs, max_iterations: int, output_len: int | None = None):
        llrs = np.ascontiguousarray(llrs)
        if llrs.dtype != np.float32:
            llrs = llrs.astype(np.float64, copy=False)
        output_len = self.n if output_len is None else output_len
        out = np.zeros(output_len, dtype=np.uint8)
        fn = self._lib.ldpc_toolbox_decoder_decode_f32 if llrs.dtype == np.float32 else self._lib.ldpc_toolbox_decoder_decode_f64
        it = fn(self._h, out.ctypes.data, output_len, llrs.ctypes.data, llrs.size, max_iterations)
        if it == -2:
            raise ValueError(f"decode: {capi.last_error()}")
        return out, it

    def decode_batch(self, llrs, max_iterations: int, output_len: int | None = None, out=None, iterations=None):
        llrs = np.ascontiguousarray(llrs)
        if llrs.dtype != np.float32:
            llrs = llrs.astype(np.float64, copy=False)
        nframes, per = llrs.shape
        output_len = self.n if output_len is None else output_len
        if out is None:
            out = np.zeros((nframes, output_len), dtype=np.uint8)
        if iterations is None:
            iterations = np.zeros(nframes, dtype=np.int32)
        fn = self._lib.ldpc_toolbox_decoder_decode_batch_f32 if llrs.dtype == np.float32 else self._lib.ldpc_toolbox_decoder_decode_batch_f64
        rc = fn(self._h, out.ctypes.data, output_len, out.strides[0], llrs.ctypes.data, per, nframes, max_iterations, iterations.ctypes.data)
        if rc != 0:
            raise ValueError(f"decode_batch: {capi.last_error()}")
        return out, iterations

    def decode_batch_posteriors(self, llrs, max_iterations: int, output_len: int | None = None):
        """Test hook (float decoders): (hard words, iterations, posterior LLRs [nframes][n] as f64)."""
        llrs = np.ascontiguousarray(llrs)
        if llrs.dtype != np.float32:
            llrs = llrs.astype(np.float64, copy=False)
        nframes, per = llrs.shape
        output_len = self.n if output_len is None else output_len
        out = np.zeros((nframes, output_len), dtype=np.uint8)
        iterations = np.zeros(nframes, dtype=np.int32)
        post = np.zeros((nframes, self.n), dtype=np.float64)
        fn = (self._lib.ldpc_toolbox_decoder_decode_batch_posteriors_f32 if llrs.dtype == np.float32
              else self._lib.ldpc_toolbox_decoder_decode_batch_posteriors_f64)
        rc = fn(self._h, out.ctypes.data, output_len, out.strides[0], llrs.ctypes.data, per, nframes, max_iterations,
                iterations.ctypes.data, post.ctypes.data)
        if rc != 0:
            raise ValueError(f"decode_batch_posteriors: {capi.last_error()}")
        return out, iterations, post

    def decode_batch_ptr(self, llrs_ptr: int, is_f64: bool, llrs_len: int, nframes: int, max_iterations: int,
                         out_ptr: int, output_len: int, output_stride: int, iters_ptr: int, device: bool, stream: int = 0):
        """Raw-pointer call (host or device buffers), used with torch pinned / CUDA tensors."""
        if device:
            fn = self._lib.ldpc_toolbox_decoder_decode_batch_device_f64 if is_f64 else self._lib.ldpc_toolbox_decoder_decode_batch_device_f32
            rc = fn(self._h, out_ptr, output_len, output_stride, llrs_ptr, llrs_len, nframes, max_iterations, iters_ptr, stream)
        else:
            fn = self._lib.ldpc_toolbox_decoder_decode_batch_f64 if is_f64 else self._lib.ldpc_toolbox_decoder_decode_batch_f32
            rc = fn(self._h, out_ptr, output_len, output_stride, llrs_ptr, llrs_len, nframes, max_iterations, iters_ptr)
        if rc != 0:
            raise ValueError(f"decode_batch: {capi.last_error()}")

    def submit_batch_ptr(self, llrs_ptr: int, is_f64: bool, llrs_len: int, nframes: int, max_iterations: int,
                         out_ptr: int, output_len: int, output_stride: int, iters_ptr: int) -> int:
        """Asynchronous host-buffer decode (pinned buffers): returns a ticket for wait()."""
        fn = self._lib.ldpc_toolbox_decoder_submit_batch_f64 if is_f64 else self._lib.ldpc_toolbox_decoder_submit_batch_f32
        t = fn(self._h, out_ptr, output_len, output_stride, llrs_ptr, llrs_len, nframes, max_iterations, iters_ptr)
        if t < 0:
            raise ValueError(f"submit_batch: {capi.last_error()}")
        return int(t)

    def wait(self, ticket: int) -> None:
        if self._lib.ldpc_toolbox_decoder_wait(self._h, ticket) != 0:
            raise RuntimeError(f"wait: {capi.last_error()}")

    def average_decode_ms(self):
        """(average BP-kernel ms, launches) over the launches since the previous call (at most the last 32)."""
        n = C.c_int64(0)
        ms = self._lib.ldpc_toolbox_decoder_average_decode_ms(self._h, C.byref(n))
        return float(ms), int(n.value)

    def last_timing(self):
        ms = (C.c_float * 3)()
        launches = self._lib.ldpc_toolbox_decoder_last_timing(self._h, ms)
        return {"ingest_ms": ms[0], "decode_ms": ms[1], "emit_ms": ms[2], "kernel_launches": int(launches)}


class Encoder:
    def __init__(self, alist: str, puncturing: str = ""):
        self._lib = capi.load()
        text, is_path = _alist_arg(alist)
        ctor = self._lib.ldpc_toolbox_encoder_ctor if is_path else self._lib.ldpc_toolbox_encoder_ctor_alist_string
        self._h = ctor(text, puncturing.encode())
        if not self._h:
            raise ValueError(f"ldpc_toolbox_encoder_ctor returned NULL: {capi.last_error()}")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ldpc_toolbox_encoder_dtor(self._h)
            self._h = None

    __del__ = close

    def encode(self, message, output_len: int):
        message = np.ascontiguousarray(message, dtype=np.uint8)
        out = np.full(output_len, 255, dtype=np.uint8)
        self._lib.ldpc_toolbox_encoder_encode(self._h, out.ctypes.data, output_len, message.ctypes.data, message.size)
        if output_len and out[0] == 255:
            raise ValueError(f"encode: {capi.last_error()}")
        return out
