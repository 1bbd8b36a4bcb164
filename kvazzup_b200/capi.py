"""ctypes binding of the C ABI declared in include/b200media.h (+ codec headers)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
_LIB = None


class B200Error(RuntimeError):
    pass


def lib_path() -> Path:
    return PKG_DIR / "libb200media.so"


def fourcc(s: str) -> int:
    a, b, c, d = (ord(ch) for ch in s)
    return a | (b << 8) | (c << 16) | (d << 24)


FOURCC = {name: fourcc(code) for name, code in {
    "I420": "I420", "I422": "I422", "NV12": "NV12", "NV21": "NV21", "YUY2": "YUY2",
    "YUYV": "YUYV", "UYVY": "UYVY", "ARGB": "ARGB", "BGRA": "BGRA", "ABGR": "ABGR",
    "RGBA": "RGBA", "24BG": "24BG", "RAW": "raw ", "MJPG": "MJPG"}.items()}

u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); every symbol include/b200media.h declares.
SIGNATURES = {
    "b200_device_count": (C.c_int, []),
    "b200_set_device": (C.c_int, [C.c_int]),
    "b200_last_error": (C.c_char_p, []),
    "b200_version": (C.c_char_p, []),
    "b200_launch_count": (C.c_ulonglong, []),
    "b200_int_peak": (C.c_double, [C.c_int]),
    "b200_yuv420_to_rgb32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint16, C.c_uint16]),
    "b200_half_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint16, C.c_uint16]),
    "b200_flip_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint16, C.c_uint16, C.c_int, C.c_int]),
    "b200_selfview": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint16, C.c_uint16, C.c_int, C.c_int, C.c_int]),
    "b200_selfview_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b200_ConvertToI420": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_uint32]),
    "b200_frame_bytes": (C.c_size_t, [C.c_uint32, C.c_int, C.c_int]),
    "b200_i420_to_rgb32_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b200_half_rgb_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b200_flip_rgb_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b200_convert_to_i420_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p]),
}


def load(rebuild: bool = False):
    """Load libb200media.so (building it in-tree first if it is missing)."""
    global _LIB
    if _LIB is not None and not rebuild:
        return _LIB
    p = lib_path()
    if rebuild or not p.exists():
        from .build import build
        build()
    if not p.exists():
        raise B200Error(f"{p} is missing: the CUDA library is the product, there is no fallback")
    l = C.CDLL(str(p), mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
    from . import codec_abi
    codec_abi.bind(l)
    _LIB = l
    return l


def lib():
    return load()


def check(rc: int, what: str, ok=(0,)):
    if rc not in ok:
        msg = lib().b200_last_error().decode(errors="replace")
        raise B200Error(f"{what} failed (rc={rc}): {msg}")
    return rc
