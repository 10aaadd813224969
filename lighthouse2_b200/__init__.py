"""lighthouse2_b200 - B200-native render core for Lighthouse 2 (hot path only).

The product is the shared library csrc/libRenderCore_B200.so (C ABI in include/lh2b.h plus the
`CreateCore` symbol the reference RenderSystem loads). This package is the thin Python host mirror
of the reference's CoreAPI_Base interface used by tests and bench.py; it fails loudly when the
CUDA library is missing - there is no CPU fallback.
"""
from .capi import load_library, LibraryMissing  # noqa: F401
from .core import RenderCore, CoreError  # noqa: F401
