"""imageprocessingtools_b200 -- B200-native per-pixel path of ppmx-edward.c.

The product is native: CUDA kernels behind the C ABI of include/ppmx_gpu.h
(libppmx_gpu.so) and the C host layer of include/ppmx_host.h (libppmx_host.so, ppmx-b200).
This Python package is a thin ctypes front end used by the tests and bench.py; it carries no
arithmetic and has no CPU fallback -- without the built libraries or without a B200 every
call raises.
"""
from .ppmx import (FT_PBM, FT_PGM, FT_PPM, LAYOUT_BITS, LAYOUT_R8, LAYOUT_RGB8, Ppmx, PpmxError, gpu_lib, host_lib,
                   PpmxOp, PpmxBand)

__all__ = ["Ppmx", "PpmxError", "PpmxOp", "PpmxBand", "gpu_lib", "host_lib", "FT_PPM", "FT_PGM", "FT_PBM",
           "LAYOUT_RGB8", "LAYOUT_R8", "LAYOUT_BITS"]
