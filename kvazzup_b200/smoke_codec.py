"""Codec part of __graft_entry__.smoke(): a tiny encode through kvz_api, checked by the caller's oracle."""
from __future__ import annotations


def run(olib) -> None:
    # `olib` (the loaded oracle library) is handed in by smoke(); this module never imports oracle/ itself.
    import ctypes as C

    import numpy as np

    from . import synth
    from .kvazaar import KvazaarFilter

    w, h, n = 192, 136, 3
    frames = [synth.camera_i420(w, h, t) for t in range(n)]
    f = KvazaarFilter({"video/ResolutionWidth": w, "video/ResolutionHeight": h, "video/QP": 30, "video/Preset": "ultrafast"})
    if not f.init():
        raise RuntimeError("KvazaarFilter.init failed")
    aus = [f.feed_input(fr)[0] for fr in frames]
    f.close()

    class Cfg(C.Structure):
        _fields_ = [(k, C.c_int) for k in ("width", "height", "qp", "intra_period", "search_range", "deblock", "hash_sei")]

    olib.orc_enc_open.restype = C.c_void_p
    enc = C.c_void_p(olib.orc_enc_open(C.byref(Cfg(w, h, 30, 64, 8, 1, 0))))
    out = np.empty(w * h * 3 + 65536, np.uint8)
    for fr, au in zip(frames, aus):
        k = olib.orc_enc_encode(enc, C.c_void_p(fr.ctypes.data), C.c_void_p(out.ctypes.data), out.size)
        assert k == len(au) and out[:k].tobytes() == au, "GPU access unit differs from the oracle's"
    olib.orc_enc_close(enc)
