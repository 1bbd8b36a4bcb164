"""The MJPG leg of the conversion oracle (oracle/mjpg_oracle.c) against libjpeg-turbo itself -- the library
family libyuv::MJPGToI420 decodes with -- through PIL: in YCbCr mode PIL hands out libjpeg's planes without
a colour conversion, so luma (never subsampled) and the chroma of 4:4:4 frames must match bit for bit, and
the chroma of 4:2:2 frames must match after libjpeg's own (documented) triangle upsampling."""
import hashlib
import io
import json
from pathlib import Path

import numpy as np
import pytest

from tests import mjpg_util

try:
    from PIL import Image
    HAVE_PIL = True
except Exception:                                          # pragma: no cover
    HAVE_PIL = False

needs = pytest.mark.skipif(not (HAVE_PIL and mjpg_util.have_cv2()), reason="PIL / cv2 not present")
GOLDEN = Path(__file__).parent / "golden"


def pil_ycbcr(jpeg):
    im = Image.open(io.BytesIO(jpeg))
    im.draft("YCbCr", im.size)
    im.load()
    assert im.mode in ("YCbCr", "L")
    return np.asarray(im)


def fancy_h2v1(row):
    """libjpeg's h2v1 "fancy" upsampling of one chroma row (jdsample.c: weights 3/4, 1/4, alternating rounding)."""
    r = row.astype(np.int32)
    n = r.size
    out = np.empty(2 * n, np.int32)
    out[0] = r[0]
    out[1] = (3 * r[0] + r[1] + 2) >> 2
    out[2:2 * n - 2:2] = (3 * r[1:n - 1] + r[0:n - 2] + 1) >> 2
    out[3:2 * n - 1:2] = (3 * r[1:n - 1] + r[2:n] + 2) >> 2
    out[2 * n - 2] = (3 * r[n - 1] + r[n - 2] + 1) >> 2
    out[2 * n - 1] = r[n - 1]
    return out.astype(np.uint8)


@needs
@pytest.mark.parametrize("w,h,q,sampling,restart,kind", [
    (64, 48, 90, "422", 0, "camera"), (640, 480, 75, "422", 0, "camera"), (640, 480, 50, "420", 0, "camera"),
    (200, 136, 95, "444", 0, "camera"), (640, 480, 30, "422", 4, "camera"), (72, 40, 100, "422", 0, "noise"),
    (1280, 720, 85, "420", 0, "camera"), (416, 240, 60, "422", 8, "noise"), (200, 136, 100, "444", 3, "noise"),
])
def test_planes_equal_libjpeg(oracle_lib, w, h, q, sampling, restart, kind):
    jpeg = mjpg_util.make_jpeg(w, h, q, sampling, restart, kind)
    planes, ow, oh = mjpg_util.oracle_planes(oracle_lib, jpeg)
    ref = pil_ycbcr(jpeg)
    assert (ow, oh) == (w, h) and len(planes) == 3
    assert np.array_equal(planes[0][:h, :w], ref[:, :, 0])
    if sampling == "444":
        assert np.array_equal(planes[1][:h, :w], ref[:, :, 1]) and np.array_equal(planes[2][:h, :w], ref[:, :, 2])
    if sampling == "422":
        for c in (1, 2):
            up = np.stack([fancy_h2v1(planes[c][j, :w // 2]) for j in range(h)])
            assert np.array_equal(up, ref[:, :, c])


@needs
def test_grey_and_missing_huffman_tables(oracle_lib):
    w, h = 160, 120
    g = mjpg_util.make_jpeg(w, h, 80, grey=True)
    planes, _, _ = mjpg_util.oracle_planes(oracle_lib, g)
    assert len(planes) == 1 and np.array_equal(planes[0][:h, :w], pil_ycbcr(g))
    rc, i420 = mjpg_util.oracle_mjpg_to_i420(oracle_lib, g, w, h)
    assert rc == 0 and np.all(i420[w * h:] == 128)
    full = mjpg_util.make_jpeg(w, h, 70, "422")
    bare = mjpg_util.strip_dht(full)
    assert len(bare) < len(full)
    assert np.array_equal(mjpg_util.oracle_mjpg_to_i420(oracle_lib, bare, w, h)[1], mjpg_util.oracle_mjpg_to_i420(oracle_lib, full, w, h)[1])


@needs
def test_subsampling_conversion_rules(oracle_lib):
    """4:2:2 -> 4:2:0 averages row pairs with round-half-up; 4:4:4 -> 4:2:0 is the rounded 2x2 box; 4:2:0 copies."""
    w, h = 200, 136
    for sampling in ("422", "444", "420"):
        jpeg = mjpg_util.make_jpeg(w, h, 85, sampling, kind="noise")
        planes, _, _ = mjpg_util.oracle_planes(oracle_lib, jpeg)
        rc, i420 = mjpg_util.oracle_mjpg_to_i420(oracle_lib, jpeg, w, h)
        assert rc == 0
        assert np.array_equal(i420[:w * h].reshape(h, w), planes[0][:h, :w])
        for c in (1, 2):
            got = i420[w * h + (c - 1) * (w * h // 4):][:w * h // 4].reshape(h // 2, w // 2)
            p = planes[c].astype(np.int32)
            if sampling == "422":
                want = (p[0:h:2, :w // 2] + p[1:h:2, :w // 2] + 1) >> 1
            elif sampling == "444":
                want = (p[0:h:2, 0:w:2] + p[0:h:2, 1:w:2] + p[1:h:2, 0:w:2] + p[1:h:2, 1:w:2] + 2) >> 2
            else:
                want = p[:h // 2, :w // 2]
            assert np.array_equal(got, want)


def test_rejects_what_libyuv_rejects(oracle_lib):
    rc, _ = mjpg_util.oracle_mjpg_to_i420(oracle_lib, b"\xff\xd8\xff\xd9", 64, 48)
    assert rc == -1
    rc, _ = mjpg_util.oracle_mjpg_to_i420(oracle_lib, b"not a jpeg at all", 64, 48)
    assert rc == -1
    if mjpg_util.have_cv2():
        jpeg = mjpg_util.make_jpeg(64, 48, 80, "422")
        assert mjpg_util.oracle_mjpg_to_i420(oracle_lib, jpeg, 128, 48)[0] == -1          # size mismatch
        assert mjpg_util.oracle_mjpg_to_i420(oracle_lib, jpeg[:len(jpeg) // 3], 64, 48)[0] in (0, -1)   # truncated: never crashes


def test_committed_frames_match_their_golden_hashes(oracle_lib):
    """tests/golden/mjpg_*.jpg with the hashes of tests/golden/mjpg_golden.json (make_mjpg_golden.py wrote both
    after checking the planes against libjpeg-turbo): keeps the oracle pinned where PIL / cv2 are missing."""
    gold = json.loads((GOLDEN / "mjpg_golden.json").read_text())
    for name, e in gold["frames"].items():
        jpeg = (GOLDEN / name).read_bytes()
        rc, i420 = mjpg_util.oracle_mjpg_to_i420(oracle_lib, jpeg, e["w"], e["h"])
        assert rc == 0 and hashlib.sha256(i420.tobytes()).hexdigest() == e["i420_sha256"], name
