"""kvazzup_b200 -- B200-native HEVC media path for the uvgComm Filter chain.

The product is ``libb200media.so`` (hand-written sm_100a CUDA kernels behind
the C ABI in ``include/*.h``).  This package is only the Python binding used by
the tests and the benchmark: it loads the shared library with ctypes and
mirrors the reference's filter interfaces.  There is no CPU fallback anywhere
in this package; compute calls raise if the library or a GPU is missing.
"""
from .capi import B200Error, lib, lib_path, load  # noqa: F401

__all__ = ["B200Error", "lib", "lib_path", "load"]
