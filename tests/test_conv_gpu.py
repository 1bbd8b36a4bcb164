"""GPU parity: conversion kernels (through the C ABI) vs the CPU oracle, bit-exact."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from kvazzup_b200 import convert, devmem, synth
from kvazzup_b200.capi import FOURCC
from tests.helpers import (RESOLUTIONS, all_uv_frame, edge_i420_frames, oracle_convert_to_i420,
                           oracle_i420_to_rgb32, ptr)

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden" / "conv_golden.json"


@pytest.mark.parametrize("wh", RESOLUTIONS + [(64, 16), (36, 20), (50, 22), (8, 2), (2, 2)])
def test_i420_to_rgb32_host_api_bit_exact(oracle_lib, wh):
    w, h = wh
    i420 = synth.noise(1234, w * h * 3 // 2)
    got = convert.yuv420_to_rgb32(i420, w, h)
    assert np.array_equal(got, oracle_i420_to_rgb32(oracle_lib, i420, w, h))
    assert (got[3::4] == 0).all()


def test_i420_to_rgb32_all_uv_and_edges(oracle_lib):
    f, w, h = all_uv_frame()
    assert np.array_equal(convert.yuv420_to_rgb32(f, w, h), oracle_i420_to_rgb32(oracle_lib, f, w, h))
    for name, fr in edge_i420_frames(64, 32).items():
        assert np.array_equal(convert.yuv420_to_rgb32(fr, 64, 32), oracle_i420_to_rgb32(oracle_lib, fr, 64, 32)), name


def test_i420_to_rgb32_golden_from_reference():
    g = json.loads(GOLDEN.read_text())
    for case in g["i420_to_rgb32"]:
        w, h = case["w"], case["h"]
        out = convert.yuv420_to_rgb32(synth.noise(case["seed"], w * h * 3 // 2), w, h)
        assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"], case
    f, w, h = all_uv_frame()
    assert hashlib.sha256(convert.yuv420_to_rgb32(f, w, h).tobytes()).hexdigest() == g["all_uv_sha256"]


def test_i420_to_rgb32_dev_batched(oracle_lib):
    w, h, n = 1280, 720, 5
    fb = w * h * 3 // 2
    src = synth.noise(99, fb * n)
    d_in = devmem.to_device(src)
    d_out = devmem.empty_u8(w * h * 4 * n)
    convert.i420_to_rgb32_dev(d_in, d_out, w, h, n, devmem.current_stream_ptr())
    got = d_out.cpu().numpy()
    for f in range(n):
        exp = oracle_i420_to_rgb32(oracle_lib, src[f * fb:(f + 1) * fb], w, h)
        assert np.array_equal(got[f * w * h * 4:(f + 1) * w * h * 4], exp), f


def test_i420_to_rgb32_full_size_property():
    """4K: the batched device path equals the per-frame host path (size-independent check)."""
    w, h, n = 3840, 2160, 3
    fb = w * h * 3 // 2
    src = synth.noise(5, fb * n)
    d_out = devmem.empty_u8(w * h * 4 * n)
    convert.i420_to_rgb32_dev(devmem.to_device(src), d_out, w, h, n, devmem.current_stream_ptr())
    got = d_out.cpu().numpy()
    one = convert.yuv420_to_rgb32(src[fb:2 * fb], w, h)
    assert np.array_equal(got[w * h * 4:2 * w * h * 4], one)


@pytest.mark.parametrize("wh", [(1280, 720), (64, 32), (50, 22), (6, 4)])
def test_half_and_flip(oracle_lib, wh):
    w, h = wh
    rgb = synth.noise(3, w * h * 4)
    exp = np.zeros((w // 2) * (h // 2) * 4, np.uint8)
    oracle_lib.oracle_half_rgb(ptr(rgb), ptr(exp), w, h)
    assert np.array_equal(convert.half_rgb(rgb, w, h), exp)
    for hor, ver in ((1, 0), (0, 1), (1, 1), (0, 0)):
        exp = np.full(w * h * 4, 9, np.uint8)
        got = np.full(w * h * 4, 9, np.uint8)
        oracle_lib.oracle_flip_rgb(ptr(rgb), ptr(exp), w, h, hor, ver)
        convert.flip_rgb(rgb, w, h, hor, ver, out=got)
        assert np.array_equal(got, exp), (hor, ver)
    # flip twice = identity
    once = convert.flip_rgb(rgb, w, h, True, True)
    assert np.array_equal(convert.flip_rgb(once, w, h, True, True), rgb)


BPP = {"YUYV": 2, "YUY2": 2, "UYVY": 2, "I422": 2, "NV12": 1.5, "NV21": 1.5, "I420": 1.5,
       "ARGB": 4, "BGRA": 4, "ABGR": 4, "RGBA": 4, "24BG": 3, "RAW": 3}


@pytest.mark.parametrize("name", sorted(BPP))
@pytest.mark.parametrize("wh", [(640, 480), (1920, 1080), (48, 18), (36, 20), (2, 2)])
def test_convert_to_i420_bit_exact(oracle_lib, name, wh):
    w, h = wh
    src = synth.noise(hash(name) % 1000 + 1, int(w * h * BPP[name]))
    rc_o, exp = oracle_convert_to_i420(oracle_lib, src, w, h, FOURCC[name])
    rc, got = convert.convert_to_i420(src, w, h, FOURCC[name])
    assert rc == rc_o == 0
    assert np.array_equal(got, exp)


def test_convert_unsupported_fourcc_leaves_output_untouched():
    src = np.zeros(64, np.uint8)
    for fcc in (2, FOURCC["MJPG"]):
        rc, out = convert.convert_to_i420(src, 4, 4, fcc, fill=0xAA)
        assert rc == -1 and (out == 0xAA).all()


def test_convert_dev_batched_yuyv(oracle_lib):
    w, h, n = 640, 480, 7
    fb = w * h * 2
    src = synth.noise(21, fb * n)
    d_out = devmem.empty_u8(w * h * 3 // 2 * n)
    convert.convert_to_i420_dev(devmem.to_device(src), d_out, w, h, FOURCC["YUYV"], n, devmem.current_stream_ptr())
    got = d_out.cpu().numpy()
    ob = w * h * 3 // 2
    for f in range(n):
        _, exp = oracle_convert_to_i420(oracle_lib, src[f * fb:(f + 1) * fb], w, h, FOURCC["YUYV"])
        assert np.array_equal(got[f * ob:(f + 1) * ob], exp), f


@pytest.mark.parametrize("wh", [(1280, 720), (640, 480), (72, 40)])
def test_fused_selfview_equals_the_three_stage_chain(oracle_lib, wh):
    """I420 -> RGB32 -> half -> mirror fused in one kernel == the oracle's (reference-pinned) stages in sequence."""
    w, h = wh
    i420 = synth.noise(17, w * h * 3 // 2)
    rgb = oracle_i420_to_rgb32(oracle_lib, i420, w, h)
    for half in (0, 1):
        if half:
            stage = np.zeros((w // 2) * (h // 2) * 4, np.uint8)
            oracle_lib.oracle_half_rgb(ptr(rgb), ptr(stage), w, h)
            ow, oh = w // 2, h // 2
        else:
            stage, ow, oh = rgb, w, h
        for hor, ver in ((0, 0), (1, 0), (0, 1), (1, 1)):
            exp = stage.copy()
            if hor or ver:
                oracle_lib.oracle_flip_rgb(ptr(stage), ptr(exp), ow, oh, hor, ver)
            got = convert.selfview(i420, w, h, bool(half), bool(hor), bool(ver))
            assert np.array_equal(got, exp), (half, hor, ver)


# ---- MJPG, the thirteenth camera format (libyuvconverter.cpp:94) ---------------------------------

def test_mjpg_committed_frames_equal_the_oracle_and_their_golden_hashes(oracle_lib):
    from tests import mjpg_util
    gold = json.loads((Path(__file__).parent / "golden" / "mjpg_golden.json").read_text())
    for name, e in gold["frames"].items():
        jpeg = np.frombuffer((Path(__file__).parent / "golden" / name).read_bytes(), np.uint8)
        rc, got = convert.convert_to_i420(jpeg, e["w"], e["h"], FOURCC["MJPG"])
        assert rc == 0, name
        assert np.array_equal(got, mjpg_util.oracle_mjpg_to_i420(oracle_lib, jpeg.tobytes(), e["w"], e["h"])[1]), name
        assert hashlib.sha256(got.tobytes()).hexdigest() == e["i420_sha256"], name


@pytest.mark.parametrize("w,h,q,sampling,restart,kind", [
    (640, 480, 75, "422", 0, "camera"), (640, 480, 50, "420", 0, "camera"), (200, 136, 95, "444", 0, "camera"),
    (640, 480, 30, "422", 4, "camera"), (72, 40, 100, "422", 0, "noise"), (1280, 720, 85, "422", 0, "camera"),
    (1920, 1080, 90, "420", 0, "camera"), (1920, 1080, 60, "422", 16, "noise"), (8, 8, 90, "444", 0, "noise"),
    (24, 16, 80, "420", 0, "noise"), (416, 240, 100, "444", 5, "noise"),
])
def test_mjpg_to_i420_bit_exact(oracle_lib, w, h, q, sampling, restart, kind):
    from tests import mjpg_util
    if not mjpg_util.have_cv2():
        pytest.skip("cv2 (JPEG encoder for the test frames) not present")
    jpeg = mjpg_util.make_jpeg(w, h, q, sampling, restart, kind)
    rc_o, exp = mjpg_util.oracle_mjpg_to_i420(oracle_lib, jpeg, w, h)
    rc, got = convert.convert_to_i420(np.frombuffer(jpeg, np.uint8), w, h, FOURCC["MJPG"])
    assert rc == rc_o == 0
    bad = np.flatnonzero(got != exp)
    assert bad.size == 0, f"{bad.size} samples differ, first at {bad[:8]}"


def test_mjpg_grey_no_dht_and_rejections(oracle_lib):
    from tests import mjpg_util
    if not mjpg_util.have_cv2():
        pytest.skip("cv2 not present")
    w, h = 160, 120
    for jpeg in (mjpg_util.make_jpeg(w, h, 80, grey=True), mjpg_util.strip_dht(mjpg_util.make_jpeg(w, h, 70, "422"))):
        rc, got = convert.convert_to_i420(np.frombuffer(jpeg, np.uint8), w, h, FOURCC["MJPG"])
        assert rc == 0 and np.array_equal(got, mjpg_util.oracle_mjpg_to_i420(oracle_lib, jpeg, w, h)[1])
    jpeg = mjpg_util.make_jpeg(w, h, 80, "422")
    for bad, bw in ((jpeg, 2 * w), (b"\xff\xd8\xff\xd9", w), (jpeg[:200], w)):       # wrong size, no scan, cut inside the headers
        rc, out = convert.convert_to_i420(np.frombuffer(bad, np.uint8), bw, h, FOURCC["MJPG"], fill=0xAA)
        assert rc == -1 and (out == 0xAA).all()
    # a frame cut inside the scan decodes (the rest is grey), as it does in the oracle: never a crash
    cut = jpeg[:len(jpeg) // 2]
    rc, got = convert.convert_to_i420(np.frombuffer(cut, np.uint8), w, h, FOURCC["MJPG"])
    rc_o, exp = mjpg_util.oracle_mjpg_to_i420(oracle_lib, cut, w, h)
    assert rc == rc_o and (rc != 0 or np.array_equal(got, exp))


def test_mjpg_device_resident_output_feeds_the_encoder(oracle_lib):
    from tests import mjpg_util
    if not mjpg_util.have_cv2():
        pytest.skip("cv2 not present")
    from kvazzup_b200.capi import lib
    import ctypes as C
    w, h = 640, 480
    jpeg = np.frombuffer(mjpg_util.make_jpeg(w, h, 85, "422"), np.uint8)
    d_out = devmem.empty_u8(w * h * 3 // 2)
    lib().b200_mjpg_to_i420_dev.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rc = lib().b200_mjpg_to_i420_dev(jpeg.ctypes.data, jpeg.size, d_out.data_ptr(), w, h, devmem.current_stream_ptr())
    assert rc == 0
    assert np.array_equal(d_out.cpu().numpy(), mjpg_util.oracle_mjpg_to_i420(oracle_lib, jpeg.tobytes(), w, h)[1])
