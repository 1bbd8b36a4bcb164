"""Writes tests/golden/hevc_golden.json: SHA-256 of the oracle encoder's access units and
reconstructions for small fixed inputs.  The streams behind these hashes were decoded bit-exactly by
FFmpeg's HEVC decoder when the file was generated (tests/test_oracle_hevc.py does that live wherever
the cv2 wheel is present); the hashes keep the oracle pinned on machines without it and catch
unintended changes of the oracle, which every GPU parity test leans on.

  python tests/golden/make_hevc_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from oracle.encoder import OracleEncoder, OracleTiledEncoder  # noqa: E402
from tests import ffhevc  # noqa: E402
from tests.test_oracle_hevc import crop_i420, frames_of, odd_size_frames, pad_i420, roi_pattern, vaq_frames  # noqa: E402

CASES = [
    {"name": "camera_192x136_qp32", "kind": "camera", "w": 192, "h": 136, "n": 4, "kw": {"qp": 32, "intra_period": 0}},
    {"name": "noise_128x72_qp10", "kind": "noise", "w": 128, "h": 72, "n": 3, "kw": {"qp": 10, "intra_period": 0}},
    {"name": "screen_416x240_qp32", "kind": "screen", "w": 416, "h": 240, "n": 5, "kw": {"qp": 32, "intra_period": 0}},
    {"name": "camera_200x200_idr3", "kind": "camera", "w": 200, "h": 200, "n": 7, "kw": {"qp": 30, "intra_period": 3}},
    {"name": "camera_416x240_roi", "kind": "camera", "w": 416, "h": 240, "n": 5, "kw": {"qp": 30, "intra_period": 0, "qp_delta": 1}, "roi": "random"},
    {"name": "camera_416x240_sao", "kind": "camera", "w": 416, "h": 240, "n": 4, "kw": {"qp": 37, "intra_period": 0, "sao": 1}},
    {"name": "camera_416x240_tiles3", "kind": "camera", "w": 416, "h": 240, "n": 4, "kw": {"qp": 30, "intra_period": 0}, "tiles": 3},
    # round 2: the presets' tools, the peer syntax of the decoder tests, a tile grid
    {"name": "sports_416x240_veryfast_tools", "kind": "sports", "w": 416, "h": 240, "n": 5,
     "kw": {"qp": 30, "intra_period": 0, "search_range": 6, "me_coarse": 16, "sao": 2, "intra_in_p": 1, "intra_satd": 1}},
    {"name": "camera_192x136_subme_satd", "kind": "camera", "w": 192, "h": 136, "n": 4, "kw": {"qp": 30, "intra_period": 0, "subme_satd": 1}},
    {"name": "sports_416x240_peer_syntax", "kind": "sports", "w": 416, "h": 240, "n": 5,
     "kw": {"qp": 30, "intra_period": 3, "tr_depth": 2, "tu4": 1, "intra_sizes": 7, "chroma_modes": 1, "sign_hiding": 1, "strong_intra": 1,
            "cb_qp_offset": 2, "cr_qp_offset": -2, "beta_offset_div2": 1, "tc_offset_div2": 1, "sao": 2, "intra_in_p": 1, "refs": 2,
            "tmvp": 1, "cabac_init": 1}},
    {"name": "camera_416x240_tiles2x2", "kind": "camera", "w": 416, "h": 240, "n": 4, "kw": {"qp": 30, "intra_period": 0, "tile_rows": 2}, "tiles": 2},
    # late round 2: variance adaptive quantisation, scaling lists (default / coded in the SPS / in the PPS), a source
    # size that is not a multiple of 8 ("src": the pictures are padded, the hashed reconstruction is the window), the
    # frame motion constraint
    {"name": "camera_416x240_vaq10", "kind": "camera", "w": 416, "h": 240, "n": 4, "kw": {"qp": 30, "intra_period": 0, "qp_delta": 1, "vaq": 10},
     "frames": "vaq"},
    {"name": "camera_416x240_scaling_default", "kind": "camera", "w": 416, "h": 240, "n": 4, "kw": {"qp": 22, "intra_period": 3, "scaling_list": 1}},
    {"name": "noise_256x136_scaling_sps", "kind": "noise", "w": 256, "h": 136, "n": 3,
     "kw": {"qp": 17, "intra_period": 0, "scaling_list": 2, "tr_depth": 2, "tu4": 1, "intra_sizes": 7}},
    {"name": "screen_640x200_scaling_pps", "kind": "screen", "w": 640, "h": 200, "n": 3, "kw": {"qp": 32, "intra_period": 2, "scaling_list": 3, "tr_depth": 1}},
    {"name": "camera_410x234_conformance_window", "kind": "camera", "w": 416, "h": 240, "n": 3,
     "kw": {"qp": 30, "intra_period": 0, "conf_right": 6, "conf_bottom": 6}, "src": [410, 234]},
    {"name": "sports_416x240_mv_frame", "kind": "sports", "w": 416, "h": 240, "n": 4,
     "kw": {"qp": 30, "intra_period": 0, "me_coarse": 16, "search_range": 4, "mv_edges": 15}},
]


def run_case(c):
    if c.get("src"):
        frames = [pad_i420(f, c["src"][0], c["src"][1], c["w"], c["h"]) for f in odd_size_frames(c["kind"], c["src"][0], c["src"][1], c["n"])]
    elif c.get("frames") == "vaq":
        frames = vaq_frames(c["kind"], c["w"], c["h"], c["n"])
    else:
        frames = frames_of(c["kind"], c["w"], c["h"], c["n"])
    if c.get("tiles"):
        enc = OracleTiledEncoder(c["w"], c["h"], c["tiles"], **c["kw"])
    else:
        enc = OracleEncoder(c["w"], c["h"], **c["kw"])
    aus, recs = [], []
    for t, f in enumerate(frames):
        if c.get("roi"):
            enc.set_ctu_dqp(roi_pattern(c["w"], c["h"], t, c["roi"]))
        aus.append(enc.encode(f))
        recs.append(crop_i420(enc.recon(), c["w"], c["h"], c["src"][0], c["src"][1]) if c.get("src") else enc.recon())
    enc.close()
    return aus, recs


def digest(aus, recs):
    return {"au_sha256": hashlib.sha256(b"".join(aus)).hexdigest(), "au_bytes": [len(a) for a in aus],
            "recon_sha256": hashlib.sha256(b"".join(r.tobytes() for r in recs)).hexdigest()}


if __name__ == "__main__":
    out = {}
    for c in CASES:
        aus, recs = run_case(c)
        ok = None
        if ffhevc.available():
            dec, errs = ffhevc.decode_stream(aus)
            ok = errs == 0 and len(dec) == len(recs) and all(np.array_equal(d[0], r) for d, r in zip(dec, recs))
            assert ok, c["name"]
        out[c["name"]] = digest(aus, recs) | {"ffmpeg_verified_when_generated": ok}
    (ROOT / "tests" / "golden" / "hevc_golden.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", len(out), "cases")
