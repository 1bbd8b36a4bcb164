"""ctypes signatures of the encoder/decoder C ABI (include/b200_kvazaar.h, b200_openhevc.h)."""
from __future__ import annotations

import ctypes as C

SIGNATURES: dict = {}


def bind(l) -> None:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
