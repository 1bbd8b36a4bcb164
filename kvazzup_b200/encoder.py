"""Python handle on the B200 HEVC encoder engine (include/b200_hevc.h).

The reference-shaped boundary is kvz_api (kvazzup_b200/kvazaar.py); this thinner wrapper is what the
parity tests and the benchmark use to keep frames in HBM and to read intermediate state.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import B200Error, lib

CU_DTYPE = np.dtype([("mvx", "<i2"), ("mvy", "<i2"), ("log2_size", "u1"), ("pred_mode", "u1"),
                     ("intra_mode", "u1"), ("cbf", "u1"), ("skip", "u1"), ("merge_idx", "u1"),
                     ("mvp_idx", "u1"), ("qp", "u1"), ("ref_idx", "u1"), ("chroma_mode", "u1"), ("tu_log2", "u1"),
                     ("flags", "u1")])
assert CU_DTYPE.itemsize == 16


class EncParams(C.Structure):
    """b200_enc_params (include/b200_hevc.h)."""
    _fields_ = [(n, C.c_int) for n in ("struct_size", "width", "height", "qp", "intra_period", "search_range", "deblock",
                                       "debug", "depth", "qp_delta", "fps_num", "fps_den", "sao", "intra_in_p", "me_coarse",
                                       "intra_satd", "subme_satd", "vaq", "scaling_list", "src_width", "src_height", "mv_edges", "vps_period")]


def preset_options(preset: str) -> dict:
    """Engine options (search_range, me_coarse, sao, intra_in_p, intra_satd, subme_satd) the kvz_api preset of that name selects."""
    p = EncParams()
    lib().b200_enc_params_default(C.byref(p))
    if lib().b200_enc_params_from_preset(preset.encode(), C.byref(p)) != 0:
        raise B200Error("unknown preset " + preset)
    return {"search_range": p.search_range, "me_coarse": p.me_coarse, "sao": p.sao, "intra_in_p": p.intra_in_p,
            "intra_satd": p.intra_satd, "subme_satd": p.subme_satd}


class GpuEncoder:
    def __init__(self, w, h, qp=32, intra_period=64, search_range=8, deblock=1, debug=0, depth=1, **options):
        """`options`: further fields of b200_enc_params by name (qp_delta, fps_num, fps_den, sao, ...)."""
        self.l = lib()
        self.w, self.h = w, h
        p = EncParams()
        self.l.b200_enc_params_default(C.byref(p))
        assert p.struct_size == C.sizeof(EncParams), "EncParams out of step with include/b200_hevc.h"
        p.width, p.height, p.qp, p.intra_period, p.search_range = w, h, qp, intra_period, search_range
        p.deblock, p.debug, p.depth = deblock, debug, depth
        known = {f[0] for f in EncParams._fields_}
        for k, val in options.items():
            if k not in known:
                raise TypeError(f"unknown encoder option {k!r}")
            setattr(p, k, int(val))
        self.h_enc = self.l.b200_enc_open_params(C.byref(p))
        if not self.h_enc:
            raise B200Error("b200_enc_open failed: " + self.l.b200_last_error().decode())
        # pictures passed in are src_w x src_h (= w x h unless a conformance window is in use)
        self.src_w, self.src_h = p.src_width or w, p.src_height or h
        self.out = np.empty(w * h * 3 + 65536, np.uint8)

    def _ret(self, n):
        if n < 0:
            raise B200Error(f"encode failed ({n}): " + self.l.b200_last_error().decode())
        return self.out[:n].tobytes()

    def set_ctu_dqp(self, dqp):
        """Per-CTU QP offsets (int8, raster) for the following pictures; None clears.  Needs qp_delta=1."""
        if dqp is None:
            rc = self.l.b200_enc_set_ctu_dqp(self.h_enc, None, 0)
        else:
            a = np.ascontiguousarray(dqp, dtype=np.int8)
            rc = self.l.b200_enc_set_ctu_dqp(self.h_enc, C.c_void_p(a.ctypes.data), a.size)
        if rc != 0:
            raise B200Error("b200_enc_set_ctu_dqp failed: " + self.l.b200_last_error().decode())

    def encode(self, i420: np.ndarray) -> bytes:
        assert i420.dtype == np.uint8 and i420.size == self.src_w * self.src_h * 3 // 2
        f = np.ascontiguousarray(i420)
        return self._ret(self.l.b200_enc_encode(self.h_enc, C.c_void_p(f.ctypes.data), C.c_void_p(self.out.ctypes.data), self.out.size))

    def encode_dev(self, d_i420) -> bytes:
        return self._ret(self.l.b200_enc_encode_dev(self.h_enc, C.c_void_p(d_i420.data_ptr()), C.c_void_p(self.out.ctypes.data), self.out.size))

    def flush(self) -> bytes:
        """Next pending access unit (b'' when the pipeline is drained)."""
        return self._ret(self.l.b200_enc_flush(self.h_enc, C.c_void_p(self.out.ctypes.data), self.out.size))

    def pending(self) -> int:
        return int(self.l.b200_enc_pending(self.h_enc))

    KERNELS = ("intra", "me", "recon", "modes", "deblock", "binarise", "arith", "pack", "sao")

    def set_profile(self, on: bool):
        self.l.b200_enc_set_profile(self.h_enc, int(on))

    def profile(self) -> dict:
        """{kernel: (total_ms, launches)} measured with CUDA events on the launching streams."""
        ms = (C.c_double * 9)()
        cnt = (C.c_ulonglong * 9)()
        self.l.b200_enc_get_profile(self.h_enc, ms, cnt, 9)
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.KERNELS)}

    def me_stats(self) -> dict:
        """Work counters of the motion search since set_profile(True)."""
        a = (C.c_ulonglong * 4)()
        self.l.b200_enc_get_me_stats(self.h_enc, a, 4)
        return {"ctus": int(a[0]), "second_set_quadrants": int(a[1]), "intra_searches": int(a[2]), "intra_cus": int(a[3])}

    def timeline(self) -> dict:
        """{kernel: (begin_ms, end_ms)} of the last returned picture, since the encoder was opened."""
        t = (C.c_float * 18)()
        self.l.b200_enc_get_timeline(self.h_enc, t, 18)
        return {k: (t[2 * i], t[2 * i + 1]) for i, k in enumerate(self.KERNELS)}

    def _read(self, what, dtype, count):
        a = np.empty(count, dtype)
        rc = self.l.b200_enc_debug_read(self.h_enc, what, C.c_void_p(a.ctypes.data), a.nbytes)
        if rc != 0:
            raise B200Error(f"debug_read({what}) failed: " + self.l.b200_last_error().decode())
        return a

    def recon(self):
        return self._read(0, np.uint8, self.w * self.h * 3 // 2)

    def recon_predeblock(self):
        return self._read(1, np.uint8, self.w * self.h * 3 // 2)

    def cu_map(self):
        return self._read(2, CU_DTYPE, (self.w // 8) * (self.h // 8))

    def levels(self):
        return self._read(3, np.int16, self.w * self.h * 3 // 2)

    def set_reference(self, i420: np.ndarray):
        f = np.ascontiguousarray(i420)
        if self.l.b200_enc_debug_set_reference(self.h_enc, C.c_void_p(f.ctypes.data)) != 0:
            raise B200Error("set_reference failed")

    def last_was_idr(self):
        return bool(self.l.b200_enc_last_was_idr(self.h_enc))

    def bins(self):
        return int(self.l.b200_enc_last_bins(self.h_enc))

    def close(self):
        if self.h_enc:
            self.l.b200_enc_close(self.h_enc)
            self.h_enc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TiledParams(C.Structure):
    """b200_tiled_params (include/b200_hevc.h)."""
    _fields_ = [(n, C.c_int) for n in ("struct_size", "width", "height", "qp", "intra_period", "search_range", "deblock",
                                       "depth", "tile_cols", "wpp", "fps_num", "fps_den", "sao", "intra_in_p", "me_coarse",
                                       "intra_satd", "subme_satd", "tile_rows", "scaling_list", "mv_edges", "vps_period")]


class GpuTiledEncoder:
    """Tile columns as independent strip encoders, optionally one GPU per strip (include/b200_hevc.h)."""

    def __init__(self, w, h, tiles, qp=32, intra_period=64, search_range=8, deblock=1, depth=1, wpp=0, devices=(), **options):
        self.l = lib()
        self.w, self.h = w, h
        devs = (C.c_int * max(len(devices), 1))(*devices)
        p = TiledParams()
        self.l.b200_tiled_params_default(C.byref(p))
        assert p.struct_size == C.sizeof(TiledParams), "TiledParams out of step with include/b200_hevc.h"
        p.width, p.height, p.qp, p.intra_period, p.search_range = w, h, qp, intra_period, search_range
        p.deblock, p.depth, p.tile_cols, p.wpp = deblock, depth, tiles, wpp
        known = {f[0] for f in TiledParams._fields_}
        for k, val in options.items():
            if k not in known:
                raise TypeError(f"unknown tiled encoder option {k!r}")
            setattr(p, k, int(val))
        self.h_enc = self.l.b200_tiled_open_params(C.byref(p), devs, len(devices))
        if not self.h_enc:
            raise B200Error("b200_tiled_open failed: " + self.l.b200_last_error().decode())
        self.out = np.empty(w * h * 3 + 65536, np.uint8)

    def _ret(self, n):
        if n < 0:
            raise B200Error(f"tiled encode failed ({n}): " + self.l.b200_last_error().decode())
        return self.out[:n].tobytes()

    def encode(self, i420: np.ndarray) -> bytes:
        f = np.ascontiguousarray(i420)
        assert f.size == self.w * self.h * 3 // 2
        return self._ret(self.l.b200_tiled_encode(self.h_enc, C.c_void_p(f.ctypes.data), C.c_void_p(self.out.ctypes.data), self.out.size))

    def set_fps(self, num: int, den: int):
        self.l.b200_tiled_set_fps(self.h_enc, num, den)

    def flush(self) -> bytes:
        return self._ret(self.l.b200_tiled_flush(self.h_enc, C.c_void_p(self.out.ctypes.data), self.out.size))

    def pending(self) -> int:
        return self.l.b200_tiled_pending(self.h_enc)

    def recon(self) -> np.ndarray:
        out = np.empty(self.w * self.h * 3 // 2, np.uint8)
        if self.l.b200_tiled_recon(self.h_enc, C.c_void_p(out.ctypes.data), out.size) != 0:
            raise B200Error("b200_tiled_recon failed: " + self.l.b200_last_error().decode())
        return out

    def close(self):
        if self.h_enc:
            self.l.b200_tiled_close(self.h_enc)
            self.h_enc = None


def satd8x8(a: np.ndarray, b: np.ndarray, w: int, h: int) -> np.ndarray:
    """SATD (8x8 Hadamard, HM convention) of every 8x8 block between two w x h planes -> (h/8, w/8) uint32."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    assert a.size == b.size == w * h
    out = np.empty((h // 8, w // 8), np.uint32)
    rc = lib().b200_satd8x8(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), w, h, C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise B200Error("b200_satd8x8 failed: " + lib().b200_last_error().decode())
    return out
