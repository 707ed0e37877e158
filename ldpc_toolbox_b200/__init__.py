"""ldpc_toolbox_b200 — B200-native batched LDPC decoding behind the ldpc-toolbox API surface.

Python here is plumbing only (ctypes over the C-ABI in include/ldpc_toolbox.h); the product is
the CUDA library built from csrc/.  See DESIGN.md and INTEGRATION.md.
"""
from .decoder import Decoder, DecoderImplementation, Encoder, implementation_names  # noqa: F401
