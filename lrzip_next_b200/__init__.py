"""lrzip-next_b200: B200-native rzip + block-backend compressor writing lrzip-next archives.

The product is the C-ABI shared library ``liblrzgpu.so`` (include/lrzgpu.h, sources under
``csrc/``); this package is the thin ctypes host mirror used by the tests, bench.py and the
multi-GPU launcher.  There is no CPU fallback: importing works anywhere, calling needs a B200.
"""
from .api import (  # noqa: F401
    BACKEND_LZMA, BACKEND_NONE, BACKEND_ZSTD, Context, LrzGpuError, Params, Sizing, Stats,
    lib_path, load_library, make_params, sizing,
)
