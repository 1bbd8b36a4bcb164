"""Shared helpers for the parity tests."""
from __future__ import annotations

import ctypes as C

import numpy as np

RESOLUTIONS = [(640, 480), (1280, 720), (1920, 1080), (2560, 1440), (3840, 2160)]


def ptr(a: np.ndarray) -> C.c_void_p:
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def oracle_i420_to_rgb32(olib, i420: np.ndarray, w: int, h: int) -> np.ndarray:
    out = np.empty(w * h * 4, np.uint8)
    olib.oracle_i420_to_rgb32(ptr(i420), ptr(out), w, h)
    return out


def oracle_convert_to_i420(olib, src: np.ndarray, w: int, h: int, fourcc: int, fill: int = 0xAA):
    out = np.full(w * h * 3 // 2, fill, np.uint8)
    ysz = w * h
    y = C.c_void_p(out.ctypes.data)
    u = C.c_void_p(out.ctypes.data + ysz)
    v = C.c_void_p(out.ctypes.data + ysz + ysz // 4)
    rc = olib.oracle_convert_to_i420(ptr(src), src.size, y, w, u, (w + 1) // 2, v, (w + 1) // 2, w, h, fourcc)
    return rc, out


def edge_i420_frames(w: int, h: int):
    """Known-answer style inputs (SURVEY.md 8d `edges`)."""
    n = w * h * 3 // 2
    ysz = w * h
    frames = {}
    for name, val in (("zeros", 0), ("ones", 255), ("mid", 128)):
        frames[name] = np.full(n, val, np.uint8)
    f = np.full(n, 240, np.uint8)
    f[:ysz] = 16
    frames["y16_uv240"] = f
    f = np.zeros(n, np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    f[:ysz] = (((yy + xx) & 1) * 255).astype(np.uint8).ravel()
    cy, cx = np.mgrid[0:h // 2, 0:w // 2]
    f[ysz:ysz + ysz // 4] = (((cy + cx) & 1) * 255).astype(np.uint8).ravel()
    f[ysz + ysz // 4:] = (((cy + cx + 1) & 1) * 255).astype(np.uint8).ravel()
    frames["checker"] = f
    f = np.full(n, 128, np.uint8)
    f[(h // 2) * w + w // 2] = 255
    frames["impulse"] = f
    return frames


def all_uv_frame():
    """512x512 frame in which every (U,V) pair occurs, with luma sweeping 0..255."""
    w = h = 512
    ysz = w * h
    f = np.empty(ysz * 3 // 2, np.uint8)
    cy, cx = np.mgrid[0:256, 0:256]
    f[ysz:ysz + ysz // 4] = cx.astype(np.uint8).ravel()
    f[ysz + ysz // 4:] = cy.astype(np.uint8).ravel()
    yy, xx = np.mgrid[0:h, 0:w]
    f[:ysz] = ((xx * 7 + yy * 13) & 255).astype(np.uint8).ravel()
    return f, w, h
