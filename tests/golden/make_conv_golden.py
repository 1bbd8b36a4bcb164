"""Generate tests/golden/conv_golden.json from the REFERENCE object code.

Run in the build container (needs /root/reference):  python tests/golden/make_conv_golden.py
The fixture pins I420->RGB32 to what /root/reference/src/media/processing/yuvconversions.cpp
(compiled unmodified into oracle/_ref/) produces, so that the oracle stays pinned on boxes
where the reference tree does not exist.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from kvazzup_b200 import synth  # noqa: E402
from tests.helpers import all_uv_frame, ptr  # noqa: E402


def main():
    ref = oracle.load_ref()
    assert ref.ref_has_avx2()
    g = {"generator": "tests/golden/make_conv_golden.py", "source": "reference yuv420_to_rgb_i_avx2 (oracle/_ref)",
         "i420_to_rgb32": []}
    for (w, h), seed in (((64, 16), 1), ((640, 480), 1234), ((1280, 720), 2), ((1920, 1080), 3)):
        i420 = synth.noise(seed, w * h * 3 // 2)
        out = np.zeros(w * h * 4, np.uint8)
        ref.ref_yuv420_to_rgb_i_avx2(ptr(i420), ptr(out), w, h)
        g["i420_to_rgb32"].append({"w": w, "h": h, "seed": seed, "sha256": hashlib.sha256(out.tobytes()).hexdigest()})
    f, w, h = all_uv_frame()
    out = np.zeros(w * h * 4, np.uint8)
    ref.ref_yuv420_to_rgb_i_avx2(ptr(f), ptr(out), w, h)
    g["all_uv_sha256"] = hashlib.sha256(out.tobytes()).hexdigest()
    # explicit pixels: a 16x2 frame per (Y,U,V) triple, first output pixel recorded
    yuv = [(0, 0, 0), (255, 255, 255), (128, 128, 128), (16, 240, 240), (235, 16, 16), (81, 90, 240),
           (145, 54, 34), (41, 240, 110), (200, 127, 129), (1, 129, 127)]
    bgr0 = []
    for Y, U, V in yuv:
        fr = np.array([Y] * 32 + [U] * 8 + [V] * 8, np.uint8)
        o = np.zeros(16 * 2 * 4, np.uint8)
        ref.ref_yuv420_to_rgb_i_avx2(ptr(fr), ptr(o), 16, 2)
        bgr0.append(o[:4].tolist())
    g["known_pixels"] = {"yuv": [list(t) for t in yuv], "bgr0": bgr0}
    out_path = Path(__file__).parent / "conv_golden.json"
    out_path.write_text(json.dumps(g, indent=1) + "\n")
    print("wrote", out_path)


if __name__ == "__main__":
    main()
