"""Writes tests/golden/mjpg_*.jpg and mjpg_golden.json: small JPEG frames from cv2's encoder and the SHA-256
of the I420 picture the oracle converts them to, after checking the oracle's luma plane against libjpeg-turbo
(PIL).  Run from the repository root: python tests/golden/make_mjpg_golden.py"""
import hashlib
import io
import json
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from tests import mjpg_util  # noqa: E402

lib = oracle.load()
out = {"generator": "tests/golden/make_mjpg_golden.py", "frames": {}}
for name, (w, h, q, sampling, restart, kind, bare) in {
        "mjpg_422_camera.jpg": (160, 120, 75, "422", 0, "camera", False),
        "mjpg_420_noise.jpg": (64, 48, 60, "420", 0, "noise", False),
        "mjpg_444_rst.jpg": (72, 40, 90, "444", 2, "camera", False),
        "mjpg_422_no_dht.jpg": (128, 72, 70, "422", 0, "camera", True)}.items():
    jpeg = mjpg_util.make_jpeg(w, h, q, sampling, restart, kind)
    if bare:
        jpeg = mjpg_util.strip_dht(jpeg)
    planes, _, _ = mjpg_util.oracle_planes(lib, jpeg)
    if not bare:
        im = Image.open(io.BytesIO(jpeg))
        im.draft("YCbCr", im.size)
        assert np.array_equal(planes[0][:h, :w], np.asarray(im)[:, :, 0]), name
    rc, i420 = mjpg_util.oracle_mjpg_to_i420(lib, jpeg, w, h)
    assert rc == 0
    (Path(__file__).parent / name).write_bytes(jpeg)
    out["frames"][name] = {"w": w, "h": h, "i420_sha256": hashlib.sha256(i420.tobytes()).hexdigest()}
(Path(__file__).parent / "mjpg_golden.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
