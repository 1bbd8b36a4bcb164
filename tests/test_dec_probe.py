"""Host logic of the decoder's parameter-set handling, without a GPU: b200_dec_probe parses the VPS / SPS /
PPS the oracle encoder writes (FFmpeg-verified streams, tests/test_oracle_hevc.py) and reports geometry,
tools, the scaling factors in force and whether the GPU decoder takes such a stream."""
import ctypes as C

import numpy as np
import pytest

from kvazzup_b200.capi import B200Error
from kvazzup_b200.openhevc import probe
from oracle.binding import load
from oracle.encoder import OracleEncoder, OracleTiledEncoder
from tests.test_oracle_hevc import frames_of


def headers(w, h, tiles=None, **kw):
    enc = OracleTiledEncoder(w, h, tiles[0], tile_rows=tiles[1], qp=30, **kw) if tiles else OracleEncoder(w, h, qp=30, **kw)
    au = enc.encode(frames_of("camera", w, h, 1)[0])
    enc.close()
    return au[:au.index(b"\x00\x00\x00\x01\x26")]          # everything before the IDR slice NAL (type 19)


def test_geometry_tools_and_timing_are_reported():
    info = probe(headers(416, 240, sao=2, sign_hiding=1, qp_delta=1, tmvp=1, refs=3, tr_depth=2, strong_intra=1, cabac_init=1,
                         fps_num=30000, fps_den=1001))
    assert (info["width"], info["height"], info["coded_width"], info["coded_height"]) == (416, 240, 416, 240)
    assert (info["fps_num"], info["fps_den"]) == (30000, 1001)
    assert info["sao"] == 1 and info["sign_hiding"] == 1 and info["qp_delta"] == 1 and info["tmvp"] == 1
    assert info["strong_intra"] == 1 and info["cabac_init_present"] == 1
    assert info["max_tr_depth_inter"] == 2 and info["max_tr_depth_intra"] == 2
    assert info["wpp"] == 1 and info["tile_cols"] == 1 and info["scaling_list"] == 0
    assert info["decodable"] == 1 and info["reason"] == ""
    plain = probe(headers(64, 64))
    assert plain["sao"] == 0 and plain["fps_num"] == 0 and plain["decodable"] == 1


def test_conformance_window_gives_the_output_size():
    info = probe(headers(1368, 768, conf_right=2))
    assert (info["width"], info["height"], info["coded_width"], info["coded_height"]) == (1366, 768, 1368, 768)
    assert (info["crop_left"], info["crop_top"]) == (0, 0) and info["decodable"] == 1
    info = probe(headers(416, 240, conf_right=6, conf_bottom=4))
    assert (info["width"], info["height"]) == (410, 236)


def test_tile_grids_are_reported():
    info = probe(headers(640, 256, tiles=(3, 2)))
    assert (info["tile_cols"], info["tile_rows"], info["wpp"]) == (3, 2, 0) and info["decodable"] == 1


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_scaling_factors_in_force_equal_the_oracle_expansion(mode):
    """Default lists, lists coded in the SPS (coded / copied / inferred-default entries, DC factors) and in the PPS:
    the table the kernels will read equals the oracle's independent expansion of the same lists."""
    info = probe(headers(128, 72, scaling_list=mode), with_scaling_table=True)
    assert info["scaling_list"] == mode and info["decodable"] == 1
    want = np.zeros(1548, np.uint8)
    load().orc_scaling_table(mode, C.c_void_p(want.ctypes.data))
    got = info["scaling_table"][:1548].copy()
    # 32x32 chroma lists do not exist in 4:2:0: matrices 1, 2, 4, 5 of sizeId 3 are never read
    for mid in (1, 2, 4, 5):
        got[(3 * 6 + mid) * 64:(3 * 6 + mid + 1) * 64] = want[(3 * 6 + mid) * 64:(3 * 6 + mid + 1) * 64]
        got[1536 + 6 + mid] = want[1536 + 6 + mid]
    assert np.array_equal(got, want)
    if mode == 1:
        m = got[:1536].reshape(4, 6, 64)
        assert (m[0] == 16).all() and m[1, 0, 63] == 115 and m[1, 3, 63] == 91 and m[2, 1, 0] == 16


def test_streams_outside_the_scope_are_named():
    hdr = bytearray(headers(64, 64))
    sps_at = hdr.index(b"\x00\x00\x00\x01\x42")
    pps_at = hdr.index(b"\x00\x00\x00\x01\x44")
    hdr[pps_at - 3] ^= 0x10                        # amp_enabled_flag (tests/test_dec_gpu.py flips the same bit)
    info = probe(bytes(hdr))
    assert info["decodable"] == 0 and "AMP" in info["reason"]
    assert sps_at < pps_at
    with pytest.raises(B200Error):
        probe(bytes(hdr[:pps_at]))                 # no PPS
    with pytest.raises(B200Error):
        probe(b"\x00\x00\x01\x42\x01" + bytes(3))  # truncated SPS


@pytest.mark.parametrize("w,h,n,tiles,kw,expect", [
    (416, 240, 3, None, {}, {"entry_points": 3, "num_ref_idx_l0": 1, "rps_pictures": 1}),
    (416, 240, 4, None, {"refs": 3, "tmvp": 1}, {"entry_points": 3, "num_ref_idx_l0": 3, "rps_pictures": 3}),
    (416, 240, 2, None, {"no_wpp": 1, "sao": 2, "tr_depth": 2}, {"entry_points": 0}),
    (640, 256, 2, (3, 2), {}, {"entry_points": 5}),                                   # one substream per tile
    (640, 256, 2, (2, 2), {"wpp": 1}, {"entry_points": 7}),                           # WPP rows inside the tiles: 2 x (2 + 2) substreams
    (416, 240, 3, None, {"tr_depth": 2, "tu4": 1, "intra_sizes": 7, "chroma_modes": 1, "sign_hiding": 1, "strong_intra": 1, "cb_qp_offset": 2,
                         "cr_qp_offset": -2, "beta_offset_div2": 1, "tc_offset_div2": 1, "sao": 2, "intra_in_p": 1, "refs": 2, "tmvp": 1,
                         "qp_delta": 1, "cabac_init": 1, "scaling_list": 2}, {"entry_points": 3, "num_ref_idx_l0": 2}),
])
def test_slice_headers_of_the_decoder_test_matrix_are_within_scope(w, h, n, tiles, kw, expect):
    """The decoder's host side (parameter-set activation, slice header parsing, the scope check) on whole oracle
    streams: every slice header is parsed and found decodable; entry points, reference indices and the reference
    picture set are what the stream was coded with."""
    qp = 29
    enc = OracleTiledEncoder(w, h, tiles[0], tile_rows=tiles[1], qp=qp, intra_period=0, **kw) if tiles else OracleEncoder(w, h, qp=qp, intra_period=0, **kw)
    stream = b"".join(enc.encode(f) for f in frames_of("sports", w, h, n))
    enc.close()
    info = probe(stream)
    assert info["decodable"] == 1, info["reason"]
    assert info["slices"] == n and info["slice_type"] == 1 and info["slice_qp"] == qp
    for k, v in expect.items():
        # the last picture of a stream with `refs` pictures in flight has min(refs, pictures before it) references
        assert info[k] == (min(v, n - 1) if k in ("num_ref_idx_l0", "rps_pictures") else v), (k, info[k])


def test_an_intra_picture_alone_reports_an_i_slice():
    enc = OracleEncoder(192, 136, qp=35)
    info = probe(enc.encode(frames_of("camera", 192, 136, 1)[0]))
    enc.close()
    assert info["slices"] == 1 and info["slice_type"] == 2 and info["slice_qp"] == 35 and info["num_ref_idx_l0"] == 0
    assert info["entry_points"] == 2 and info["decodable"] == 1
