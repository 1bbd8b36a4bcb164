"""TEST INFRASTRUCTURE ONLY: Python wrapper of the oracle HEVC encoder (oracle/hevc_enc.c)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .binding import load
from .sigs import OrcCu, OrcEncCfg

CU_DTYPE = np.dtype([("mvx", "<i2"), ("mvy", "<i2"), ("log2_size", "u1"), ("pred_mode", "u1"),
                     ("intra_mode", "u1"), ("cbf", "u1"), ("skip", "u1"), ("merge_idx", "u1"),
                     ("mvp_idx", "u1"), ("qp", "u1"), ("ref_idx", "u1"), ("chroma_mode", "u1"), ("tu_log2", "u1"),
                     ("flags", "u1")])
assert CU_DTYPE.itemsize == C.sizeof(OrcCu) == 16


class OracleEncoder:
    def __init__(self, w, h, qp=32, intra_period=64, search_range=8, deblock=1, **options):
        """`options`: any further field of orc_enc_cfg_t (oracle/hevc_enc.h) by name, e.g. sao=1, fps_num=30."""
        self.lib = load()
        self.w, self.h = w, h
        cfg = OrcEncCfg(width=w, height=h, qp=qp, intra_period=intra_period, search_range=search_range, deblock=deblock)
        known = {f[0] for f in OrcEncCfg._fields_}
        for k, val in options.items():
            if k not in known:
                raise TypeError(f"unknown oracle encoder option {k!r}")
            setattr(cfg, k, int(val))
        self.h_enc = self.lib.orc_enc_open(C.byref(cfg))
        if not self.h_enc:
            raise ValueError("orc_enc_open rejected the configuration")
        self.out = np.empty(w * h * 3 + 65536, np.uint8)

    def encode(self, i420: np.ndarray) -> bytes:
        assert i420.dtype == np.uint8 and i420.size == self.w * self.h * 3 // 2
        frame = np.ascontiguousarray(i420)
        n = self.lib.orc_enc_encode(self.h_enc, C.c_void_p(frame.ctypes.data), C.c_void_p(self.out.ctypes.data),
                                    self.out.size)
        if n < 0:
            raise RuntimeError(f"orc_enc_encode failed ({n})")
        return self.out[:n].tobytes()

    def set_qp(self, qp):
        """Slice QP of the following pictures."""
        if self.lib.orc_enc_set_qp(self.h_enc, int(qp)) != 0:
            raise ValueError("qp out of range")

    def set_ctu_dqp(self, dqp):
        """Per-CTU QP offsets (int8, ctb_cols*ctb_rows, raster) for the following pictures; None clears."""
        if dqp is None:
            rc = self.lib.orc_enc_set_ctu_dqp(self.h_enc, None)
        else:
            a = np.ascontiguousarray(dqp, dtype=np.int8)
            assert a.size == ((self.w + 63) // 64) * ((self.h + 63) // 64)
            rc = self.lib.orc_enc_set_ctu_dqp(self.h_enc, C.c_void_p(a.ctypes.data))
        if rc != 0:
            raise ValueError("set_ctu_dqp needs qp_delta=1")

    def _view(self, ptr, dtype, count):
        t = (C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(t, dtype=dtype, count=count).copy()

    def recon(self):
        return self._view(self.lib.orc_enc_recon(self.h_enc), np.uint8, self.w * self.h * 3 // 2)

    def recon_predeblock(self):
        return self._view(self.lib.orc_enc_recon_predeblock(self.h_enc), np.uint8, self.w * self.h * 3 // 2)

    def levels(self):
        return self._view(self.lib.orc_enc_levels(self.h_enc), np.int16, self.w * self.h * 3 // 2)

    def cu_map(self):
        return self._view(self.lib.orc_enc_cu_map(self.h_enc), CU_DTYPE, (self.w // 8) * (self.h // 8))

    def last_was_idr(self):
        return bool(self.lib.orc_enc_last_was_idr(self.h_enc))

    def bins(self):
        return int(self.lib.orc_enc_bins(self.h_enc))

    def close(self):
        if self.h_enc:
            self.lib.orc_enc_close(self.h_enc)
            self.h_enc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleTiledEncoder:
    """Tile columns coded as independent strips (oracle/hevc_enc.c: orc_tiled_*)."""

    def __init__(self, w, h, tiles, qp=32, intra_period=64, search_range=8, deblock=1, hash_sei=0, wpp=0, tile_rows=1, **options):
        self.lib = load()
        self.w, self.h = w, h
        cfg = OrcEncCfg(width=w, height=h, qp=qp, intra_period=intra_period, search_range=search_range, deblock=deblock,
                        hash_sei=hash_sei, no_wpp=0 if wpp else 1)
        known = {f[0] for f in OrcEncCfg._fields_}
        for k, val in options.items():
            if k not in known:
                raise TypeError(f"unknown oracle encoder option {k!r}")
            setattr(cfg, k, int(val))
        self.h_enc = self.lib.orc_tiled_open2(C.byref(cfg), tiles, tile_rows)
        if not self.h_enc:
            raise ValueError("orc_tiled_open rejected the configuration")
        self.out = np.empty(w * h * 3 + 65536, np.uint8)

    def encode(self, i420: np.ndarray) -> bytes:
        frame = np.ascontiguousarray(i420)
        n = self.lib.orc_tiled_encode(self.h_enc, C.c_void_p(frame.ctypes.data), C.c_void_p(self.out.ctypes.data), self.out.size)
        if n < 0:
            raise RuntimeError(f"orc_tiled_encode failed ({n})")
        return self.out[:n].tobytes()

    def recon(self):
        n = self.w * self.h * 3 // 2
        t = (C.c_uint8 * n).from_address(self.lib.orc_tiled_recon(self.h_enc))
        return np.frombuffer(t, dtype=np.uint8, count=n).copy()

    def close(self):
        if self.h_enc:
            self.lib.orc_tiled_close(self.h_enc)
            self.h_enc = None
