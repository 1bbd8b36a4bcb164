"""Helpers of the MJPG tests (test infrastructure): JPEG frames from cv2's encoder (libjpeg-turbo) and the
oracle's decoder through ctypes."""
from __future__ import annotations

import ctypes as C

import numpy as np

from kvazzup_b200 import synth

SAMPLING = {"444": 0x111111, "422": 0x211111, "420": 0x221111}      # cv2.IMWRITE_JPEG_SAMPLING_FACTOR_*


def have_cv2() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def make_jpeg(w, h, quality=80, sampling="422", restart=0, kind="camera", t=0, grey=False) -> bytes:
    import cv2
    i420 = synth.noise(7 + t, w * h * 3 // 2) if kind == "noise" else synth.camera_i420(w, h, t)
    img = cv2.cvtColor(i420.reshape(h * 3 // 2, w), cv2.COLOR_YUV2BGR_I420)
    if grey:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    params = [cv2.IMWRITE_JPEG_QUALITY, quality]
    if not grey:
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sampling]]
    if restart:
        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, restart]
    ok, enc = cv2.imencode(".jpg", img, params)
    assert ok
    return enc.tobytes()


def strip_dht(jpeg: bytes) -> bytes:
    """Removes the DHT segments (UVC cameras and AVI MJPG leave them out: the decoder then uses the tables
    of T.81 Annex K, which are the ones cv2 / libjpeg write when optimisation is off)."""
    out, p = bytearray(jpeg[:2]), 2
    while p < len(jpeg):
        assert jpeg[p] == 0xFF
        m = jpeg[p + 1]
        if m == 0xDA:
            out += jpeg[p:]
            break
        n = (jpeg[p + 2] << 8) | jpeg[p + 3]
        if m != 0xC4:
            out += jpeg[p:p + 2 + n]
        p += 2 + n
    return bytes(out)


def oracle_planes(lib, jpeg: bytes):
    """-> (component planes padded to whole MCUs, width, height)"""
    a = np.frombuffer(jpeg, np.uint8)
    pw, ph, w, h = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int(), C.c_int()
    nc = lib.oracle_mjpg_planes(a.ctypes.data, a.size, None, pw, ph, C.byref(w), C.byref(h))
    if nc <= 0:
        raise ValueError("oracle cannot decode this JPEG")
    bufs = [np.empty(pw[i] * ph[i], np.uint8) for i in range(nc)]
    ptrs = (C.c_void_p * 3)(*[b.ctypes.data for b in bufs] + [None] * (3 - nc))
    lib.oracle_mjpg_planes(a.ctypes.data, a.size, ptrs, pw, ph, C.byref(w), C.byref(h))
    return [bufs[i].reshape(ph[i], pw[i]) for i in range(nc)], w.value, h.value


def oracle_mjpg_to_i420(lib, jpeg: bytes, w, h):
    a = np.frombuffer(jpeg, np.uint8)
    out = np.zeros(w * h * 3 // 2, np.uint8)
    ysz = w * h
    base = out.ctypes.data
    rc = lib.oracle_mjpg_to_i420(a.ctypes.data, a.size, base, w, base + ysz, w // 2, base + ysz + ysz // 4, w // 2, w, h)
    return rc, out
