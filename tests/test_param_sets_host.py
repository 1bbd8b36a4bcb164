"""Host logic of the encoder's bitstream headers, without a GPU: the VPS / SPS / PPS the product writes
(b200_enc_parameter_sets: hevc_encoder.cu write_parameter_sets, what every IDR access unit and kvz_api's
encoder_headers carry) equal the oracle encoder's byte for byte -- FFmpeg decodes the oracle's streams -- and the
product's own parser (b200_dec_probe) reads them back."""
import ctypes as C

import numpy as np
import pytest

from kvazzup_b200.capi import lib
from kvazzup_b200.encoder import EncParams
from kvazzup_b200.openhevc import probe
from tests.test_dec_probe import headers


def product_sets(w, h, tiles=(1, 1), wpp=1, **opts):
    p = EncParams()
    lib().b200_enc_params_default(C.byref(p))
    p.width, p.height = w, h
    for k, v in opts.items():
        setattr(p, k, v)
    out = np.zeros(4096, np.uint8)
    n = lib().b200_enc_parameter_sets(C.byref(p), tiles[0], tiles[1], wpp, C.c_void_p(out.ctypes.data), out.size)
    assert n > 0, lib().b200_last_error().decode()
    return out[:n].tobytes()


@pytest.mark.parametrize("w,h,tiles,gpu,orc", [
    (416, 240, None, {}, {}),
    (1920, 1080, None, {"sao": 2, "fps_num": 30, "fps_den": 1}, {"sao": 2, "fps_num": 30, "fps_den": 1}),
    (640, 480, None, {"qp_delta": 1, "deblock": 0}, {"qp_delta": 1, "deblock": 0}),
    (416, 240, None, {"scaling_list": 1, "sao": 1}, {"scaling_list": 1, "sao": 1}),
    (1368, 768, None, {"src_width": 1366, "src_height": 768}, {"conf_right": 2}),
    (416, 240, None, {"src_width": 410, "src_height": 234, "fps_num": 30000, "fps_den": 1001}, {"conf_right": 6, "conf_bottom": 6, "fps_num": 30000, "fps_den": 1001}),
    (640, 256, (3, 2), {"sao": 2}, {"sao": 2}),
    (3840, 2160, (8, 1), {}, {}),
])
def test_parameter_sets_equal_the_oracles(w, h, tiles, gpu, orc):
    if tiles:
        for wpp in (0, 1):
            want = headers(w, h, tiles=tiles, wpp=wpp, **orc)
            assert product_sets(w, h, tiles, wpp, **gpu) == want, wpp
    else:
        want = headers(w, h, **orc)
        assert product_sets(w, h, **gpu) == want
    info = probe(product_sets(w, h, tiles or (1, 1), 0 if tiles else 1, **gpu))
    assert info["decodable"] == 1
    assert (info["width"], info["height"]) == (gpu.get("src_width", w), gpu.get("src_height", h))
    assert (info["tile_cols"], info["tile_rows"]) == (tiles or (1, 1))
    assert info["scaling_list"] == gpu.get("scaling_list", 0) and info["sao"] == (1 if gpu.get("sao") else 0)


def test_bad_arguments_are_refused():
    p = EncParams()
    lib().b200_enc_params_default(C.byref(p))
    p.width, p.height = 100, 64                       # coded sizes are multiples of 8
    out = np.zeros(64, np.uint8)
    assert lib().b200_enc_parameter_sets(C.byref(p), 1, 1, 1, C.c_void_p(out.ctypes.data), out.size) < 0
    p.width = 128
    n = lib().b200_enc_parameter_sets(C.byref(p), 1, 1, 1, C.c_void_p(out.ctypes.data), 8)
    assert n < -8                                     # too small: minus the size needed
    big = np.zeros(-n, np.uint8)
    assert lib().b200_enc_parameter_sets(C.byref(p), 1, 1, 1, C.c_void_p(big.ctypes.data), big.size) == -n
