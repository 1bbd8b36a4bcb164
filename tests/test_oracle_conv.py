"""CPU-only: pin the conversion oracle against the reference's own object code and golden fixtures."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from kvazzup_b200 import synth
from kvazzup_b200.capi import FOURCC
from tests.helpers import (all_uv_frame, edge_i420_frames, oracle_convert_to_i420,
                           oracle_i420_to_rgb32, ptr)

GOLDEN = Path(__file__).parent / "golden" / "conv_golden.json"


def ref_rgb(ref, variant, i420, w, h):
    out = np.full(w * h * 4, 0x5A, np.uint8)
    src = i420.copy()
    if variant == "avx2_mt":
        ref.ref_yuv420_to_rgb_i_avx2_mt(ptr(src), ptr(out), w, h, 4)
    elif variant == "avx2":
        ref.ref_yuv420_to_rgb_i_avx2(ptr(src), ptr(out), w, h)
    else:
        ref.ref_yuv420_to_rgb_i_sse41(ptr(src), ptr(out), w, h)
    return out


needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("variant", ["avx2_mt", "avx2", "sse41"])
@pytest.mark.parametrize("wh", [(640, 480), (1280, 720), (64, 16), (1920, 1080)])
def test_oracle_matches_reference_simd(oracle_lib, variant, wh):
    w, h = wh
    ref = oracle.load_ref()
    if not ref.ref_has_avx2():
        pytest.skip("host lacks AVX2")
    i420 = synth.noise(1234, w * h * 3 // 2)
    assert np.array_equal(oracle_i420_to_rgb32(oracle_lib, i420, w, h), ref_rgb(ref, variant, i420, w, h))


@needs_ref
def test_oracle_matches_reference_all_uv_pairs(oracle_lib):
    ref = oracle.load_ref()
    f, w, h = all_uv_frame()
    assert np.array_equal(oracle_i420_to_rgb32(oracle_lib, f, w, h), ref_rgb(ref, "avx2", f, w, h))


@needs_ref
def test_oracle_matches_reference_edges(oracle_lib):
    ref = oracle.load_ref()
    for name, f in edge_i420_frames(64, 32).items():
        assert np.array_equal(oracle_i420_to_rgb32(oracle_lib, f, 64, 32), ref_rgb(ref, "avx2", f, 64, 32)), name


@needs_ref
def test_reference_scalar_fallback_is_the_documented_swap(oracle_lib):
    """`_c` == SIMD with U/V swapped on input and bytes 0/2 swapped on output."""
    ref = oracle.load_ref()
    w, h = 64, 32
    ysz = w * h
    f = synth.noise(7, ysz * 3 // 2)
    swapped = f.copy()
    swapped[ysz:ysz + ysz // 4] = f[ysz + ysz // 4:]
    swapped[ysz + ysz // 4:] = f[ysz:ysz + ysz // 4]
    out_c = np.zeros(ysz * 4, np.uint8)
    ref.ref_yuv420_to_rgb_i_c(ptr(swapped), ptr(out_c), w, h)
    simd = oracle_i420_to_rgb32(oracle_lib, f, w, h).reshape(-1, 4)
    c = out_c.reshape(-1, 4)
    assert np.array_equal(c[:, 0], simd[:, 2]) and np.array_equal(c[:, 1], simd[:, 1]) and np.array_equal(c[:, 2], simd[:, 0])


@needs_ref
@pytest.mark.parametrize("wh", [(64, 32), (1280, 720)])
def test_half_and_flip_match_reference(oracle_lib, wh):
    w, h = wh
    ref = oracle.load_ref()
    rgb = synth.noise(3, w * h * 4)
    a = np.zeros(w * h, np.uint8)
    b = np.zeros(w * h, np.uint8)
    oracle_lib.oracle_half_rgb(ptr(rgb), ptr(a), w, h)
    ref.ref_half_rgb(ptr(rgb), ptr(b), w, h)
    assert np.array_equal(a, b)
    for hor, ver in ((1, 0), (0, 1), (1, 1), (0, 0)):
        a = np.full(w * h * 4, 9, np.uint8)
        b = np.full(w * h * 4, 9, np.uint8)
        oracle_lib.oracle_flip_rgb(ptr(rgb), ptr(a), w, h, hor, ver)
        ref.ref_flip_rgb(ptr(rgb), ptr(b), w, h, hor, ver)
        assert np.array_equal(a, b), (hor, ver)


def test_oracle_matches_golden_fixtures(oracle_lib):
    """Fixtures were produced by the reference object code (tests/golden/make_conv_golden.py)."""
    g = json.loads(GOLDEN.read_text())
    for case in g["i420_to_rgb32"]:
        w, h = case["w"], case["h"]
        i420 = synth.noise(case["seed"], w * h * 3 // 2)
        out = oracle_i420_to_rgb32(oracle_lib, i420, w, h)
        assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"], case
    f, w, h = all_uv_frame()
    out = oracle_i420_to_rgb32(oracle_lib, f, w, h)
    assert hashlib.sha256(out.tobytes()).hexdigest() == g["all_uv_sha256"]
    px = g["known_pixels"]
    for (Y, U, V), bgr0 in zip(px["yuv"], px["bgr0"]):
        fr = np.array([Y] * 4 + [U] + [V], np.uint8)
        assert oracle_i420_to_rgb32(oracle_lib, fr, 2, 2)[:4].tolist() == bgr0


# ---- camera formats -> I420 (restated libyuv contract; unpinned) -------------------------------

def test_yuyv_known_answer(oracle_lib):
    w, h = 4, 2
    #        Y0  U   Y1  V    Y2  U   Y3  V
    row0 = [10, 100, 20, 200, 30, 101, 40, 201]
    row1 = [50, 103, 60, 203, 70, 104, 80, 206]
    src = np.array(row0 + row1, np.uint8)
    rc, out = oracle_convert_to_i420(oracle_lib, src, w, h, FOURCC["YUYV"])
    assert rc == 0
    assert out[:8].tolist() == [10, 20, 30, 40, 50, 60, 70, 80]
    assert out[8:10].tolist() == [(100 + 103 + 1) >> 1, (101 + 104 + 1) >> 1]
    assert out[10:12].tolist() == [(200 + 203 + 1) >> 1, (201 + 206 + 1) >> 1]


def test_uyvy_is_byte_swapped_yuyv(oracle_lib):
    w, h = 64, 16
    src = synth.noise(11, w * h * 2)
    sw = src.reshape(-1, 2)[:, ::-1].copy().ravel()
    _, a = oracle_convert_to_i420(oracle_lib, src, w, h, FOURCC["YUYV"])
    _, b = oracle_convert_to_i420(oracle_lib, sw, w, h, FOURCC["UYVY"])
    assert np.array_equal(a, b)


def test_nv12_nv21_and_i422(oracle_lib):
    w, h = 32, 8
    ysz = w * h
    src = synth.noise(5, ysz * 3 // 2)
    _, a = oracle_convert_to_i420(oracle_lib, src, w, h, FOURCC["NV12"])
    _, b = oracle_convert_to_i420(oracle_lib, src, w, h, FOURCC["NV21"])
    assert np.array_equal(a[:ysz], src[:ysz])
    assert np.array_equal(a[ysz:ysz + ysz // 4], src[ysz::2]) and np.array_equal(a[ysz + ysz // 4:], src[ysz + 1::2])
    assert np.array_equal(b[ysz:ysz + ysz // 4], src[ysz + 1::2]) and np.array_equal(b[ysz + ysz // 4:], src[ysz::2])
    s422 = synth.noise(6, ysz * 2)
    _, c = oracle_convert_to_i420(oracle_lib, s422, w, h, FOURCC["I422"])
    U = s422[ysz:ysz + ysz // 2].reshape(h, w // 2).astype(np.int32)
    assert np.array_equal(c[ysz:ysz + ysz // 4].reshape(h // 2, w // 2), ((U[0::2] + U[1::2] + 1) >> 1).astype(np.uint8))


def test_rgb_family_channel_orders_and_known_values(oracle_lib):
    w, h = 2, 2
    # one uniform colour: R=200 G=100 B=50
    R, G, B, A = 200, 100, 50, 77
    Y = (66 * R + 129 * G + 25 * B + 0x1080) >> 8
    U = (112 * B - 74 * G - 38 * R + 0x8080) >> 8
    V = (112 * R - 94 * G - 18 * B + 0x8080) >> 8
    layouts = {"ARGB": [B, G, R, A], "BGRA": [A, R, G, B], "ABGR": [R, G, B, A], "RGBA": [A, B, G, R],
               "24BG": [B, G, R], "RAW": [R, G, B]}
    for name, px in layouts.items():
        src = np.array(px * 4, np.uint8)
        rc, out = oracle_convert_to_i420(oracle_lib, src, w, h, FOURCC[name])
        assert rc == 0 and out.tolist() == [Y] * 4 + [U, V], name
    # black and white land on the limited-range end points
    _, blk = oracle_convert_to_i420(oracle_lib, np.zeros(16, np.uint8), 2, 2, FOURCC["ARGB"])
    _, wht = oracle_convert_to_i420(oracle_lib, np.full(16, 255, np.uint8), 2, 2, FOURCC["ARGB"])
    assert blk.tolist() == [16] * 4 + [128, 128] and wht.tolist() == [235] * 4 + [128, 128]


def test_unsupported_fourcc_leaves_output_untouched(oracle_lib):
    src = np.zeros(64, np.uint8)
    for fcc in (2, FOURCC["MJPG"]):      # `2` is what the reference passes for DT_RGB24VIDEO
        rc, out = oracle_convert_to_i420(oracle_lib, src, 4, 4, fcc, fill=0xAA)
        assert rc == -1 and (out == 0xAA).all()
