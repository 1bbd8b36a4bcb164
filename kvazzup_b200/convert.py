"""Python mirror of the conversion filters (LibYUVConverter, YUVtoRGB32, HalfRGBFilter, flip).

Each function calls straight into libb200media.so; names follow the reference filters:
  src/media/processing/libyuvconverter.cpp:20-136, yuvtorgb32.cpp:29-64, halfrgbfilter.cpp:21-43,
  filter.cpp:263-294.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import B200Error, FOURCC, check, lib


def _p(a: np.ndarray) -> C.c_void_p:
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


# ---- host-buffer path (what the reference filters call) ---------------------------------------

def yuv420_to_rgb32(i420: np.ndarray, w: int, h: int) -> np.ndarray:
    out = np.empty(w * h * 4, np.uint8)
    check(lib().b200_yuv420_to_rgb32(_p(i420), _p(out), w, h), "b200_yuv420_to_rgb32", ok=(1,))
    return out


def half_rgb(rgb: np.ndarray, w: int, h: int) -> np.ndarray:
    out = np.empty((w // 2) * ((h + 1) // 2) * 4, np.uint8)
    check(lib().b200_half_rgb(_p(rgb), _p(out), w, h), "b200_half_rgb")
    return out


def flip_rgb(rgb: np.ndarray, w: int, h: int, horizontally: bool, vertically: bool, out: np.ndarray | None = None):
    if out is None:
        out = np.zeros(w * h * 4, np.uint8)
    check(lib().b200_flip_rgb(_p(rgb), _p(out), w, h, int(horizontally), int(vertically)), "b200_flip_rgb")
    return out


def selfview(i420: np.ndarray, w: int, h: int, half: bool, horizontally: bool, vertically: bool) -> np.ndarray:
    """Fused self-view chain: YUVtoRGB32 -> HalfRGBFilter -> mirror (filtergraph.cpp:247-325)."""
    ow, oh = (w // 2, (h + 1) // 2) if half else (w, h)
    out = np.empty(ow * oh * 4, np.uint8)
    check(lib().b200_selfview(_p(i420), _p(out), w, h, int(half), int(horizontally), int(vertically)), "b200_selfview")
    return out


def convert_to_i420(sample: np.ndarray, w: int, h: int, fourcc: int, fill: int | None = None):
    """Returns (rc, packed I420).  rc follows libyuv: 0 ok, -1 unsupported (output untouched)."""
    out = np.empty(w * h * 3 // 2, np.uint8) if fill is None else np.full(w * h * 3 // 2, fill, np.uint8)
    ysz = w * h
    base = out.ctypes.data
    rc = lib().b200_ConvertToI420(_p(sample), sample.size, C.c_void_p(base), w,
                                  C.c_void_p(base + ysz), (w + 1) // 2,
                                  C.c_void_p(base + ysz + ysz // 4), (w + 1) // 2,
                                  0, 0, w, h, w, h, 0, fourcc)
    if rc == -2:
        raise B200Error(lib().b200_last_error().decode())
    return rc, out


# ---- device-resident batched path ---------------------------------------------------------------

class _Ptr:
    """Lets the *_dev wrappers take either a torch tensor or a raw device address."""
    def __init__(self, a):
        self.a = a

    def data_ptr(self):
        return self.a.data_ptr() if hasattr(self.a, "data_ptr") else int(self.a)


def i420_to_rgb32_dev(d_in, d_out, w: int, h: int, n: int, stream: int = 0):
    check(lib().b200_i420_to_rgb32_dev(_Ptr(d_in).data_ptr(), _Ptr(d_out).data_ptr(), w, h, n, stream), "b200_i420_to_rgb32_dev")


def half_rgb_dev(d_in, d_out, w: int, h: int, n: int, stream: int = 0):
    check(lib().b200_half_rgb_dev(_Ptr(d_in).data_ptr(), _Ptr(d_out).data_ptr(), w, h, n, stream), "b200_half_rgb_dev")


def flip_rgb_dev(d_in, d_out, w: int, h: int, hor: bool, ver: bool, n: int, stream: int = 0):
    check(lib().b200_flip_rgb_dev(_Ptr(d_in).data_ptr(), _Ptr(d_out).data_ptr(), w, h, int(hor), int(ver), n, stream),
          "b200_flip_rgb_dev")


def convert_to_i420_dev(d_src, d_dst, w: int, h: int, fourcc: int, n: int, stream: int = 0):
    check(lib().b200_convert_to_i420_dev(_Ptr(d_src).data_ptr(), _Ptr(d_dst).data_ptr(), w, h, fourcc, n, stream),
          "b200_convert_to_i420_dev")


__all__ = ["FOURCC", "yuv420_to_rgb32", "half_rgb", "flip_rgb", "convert_to_i420", "i420_to_rgb32_dev",
           "half_rgb_dev", "flip_rgb_dev", "convert_to_i420_dev"]
