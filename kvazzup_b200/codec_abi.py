"""ctypes signatures of the encoder/decoder C ABI (include/b200_kvazaar.h, b200_openhevc.h)."""
from __future__ import annotations

import ctypes as C

v = C.c_void_p
i = C.c_int

SIGNATURES: dict = {
    "b200_enc_open": (v, [i, i, i, i, i, i, i, i]),
    "b200_enc_open_roi": (v, [i, i, i, i, i, i, i, i]),
    "b200_enc_params_default": (None, [v]),
    "b200_enc_open_params": (v, [v]),
    "b200_enc_parameter_sets": (i, [v, i, i, i, v, i]),
    "b200_kvz_set_bitrate": (i, [v, i]),
    "b200_rtcp_bitrate_update": (i, [i, i, i]),
    "b200_enc_params_from_preset": (i, [C.c_char_p, v]),
    "b200_enc_set_ctu_dqp": (i, [v, v, i]),
    "b200_enc_flush": (i, [v, v, i]),
    "b200_enc_pending": (i, [v]),
    "b200_enc_set_profile": (i, [v, i]),
    "b200_enc_get_profile": (i, [v, v, v, i]),
    "b200_enc_get_me_stats": (i, [v, v, i]),
    "b200_enc_get_timeline": (i, [v, v, i]),
    "b200_enc_close": (None, [v]),
    "b200_enc_encode": (i, [v, v, v, i]),
    "b200_enc_encode_dev": (i, [v, v, v, i]),
    "b200_enc_last_was_idr": (i, [v]),
    "b200_enc_last_bins": (C.c_ulonglong, [v]),
    "b200_enc_debug_read": (i, [v, i, v, C.c_size_t]),
    "b200_enc_debug_set_reference": (i, [v, v]),
    "b200_tiled_open": (v, [i, i, i, i, i, i, i, i, i, v, i]),
    "b200_tiled_params_default": (None, [v]),
    "b200_tiled_open_params": (v, [v, v, i]),
    "b200_tiled_close": (None, [v]),
    "b200_tiled_set_fps": (None, [v, i, i]),
    "b200_tiled_encode": (i, [v, v, v, i]),
    "b200_tiled_flush": (i, [v, v, i]),
    "b200_tiled_pending": (i, [v]),
    "b200_tiled_recon": (i, [v, v, C.c_size_t]),
    "b200_satd8x8": (i, [v, v, i, i, v]),
    "b200_satd8x8_dev": (i, [v, v, i, i, v, v]),
    "libOpenHevcInit": (v, [i, i]),
    "libOpenHevcStartDecoder": (i, [v]),
    "libOpenHevcDecode": (i, [v, v, i, C.c_int64]),
    "libOpenHevcGetOutput": (i, [v, i, v]),
    "libOpenHevcGetPictureInfo": (None, [v, v]),
    "libOpenHevcSetTemporalLayer_id": (None, [v, i]),
    "libOpenHevcSetActiveDecoders": (None, [v, i]),
    "libOpenHevcSetViewLayers": (None, [v, i]),
    "libOpenHevcSetDebugMode": (None, [v, i]),
    "libOpenHevcSetCheckMD5": (None, [v, i]),
    "libOpenHevcVersion": (C.c_char_p, [v]),
    "libOpenHevcFlush": (None, [v]),
    "libOpenHevcClose": (None, [v]),
    "b200_dec_last_picture": (i, [v, v, i]),
    "b200_dec_output_dev": (v, [v]),
    "b200_dec_missing_refs": (i, [v]),
    "b200_dec_set_host_output": (None, [v, i]),
    "b200_dec_probe": (i, [v, C.c_size_t, v, v]),
}


def bind(l) -> None:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
