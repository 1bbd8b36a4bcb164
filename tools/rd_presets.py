"""BD-rate of this round's encoder presets against the round-1 encoder (zero-centred +-12 search, no
SAO, no intra CUs in P pictures, SAD-only intra search), per synthetic sequence.  Computed with the
CPU oracle, whose streams the GPU encoder reproduces byte for byte (tests/test_enc_gpu.py), so the
numbers are the GPU encoder's.  Bjontegaard delta rate over PSNR-Y at QP 22/27/32/37.

  python tools/rd_presets.py [frames]      -> JSON lines (profiles/r02_rd_presets.jsonl)
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import oracle  # noqa: E402
from kvazzup_b200 import synth  # noqa: E402
from oracle.encoder import OracleEncoder  # noqa: E402
from tools.rd_oracle_options import bd_rate  # noqa: E402

ROUND1 = {"search_range": 12}
PRESETS = {          # what kvz_api.cu kPresets selects (b200_enc_params_from_preset)
    "ultrafast": {"search_range": 4, "me_coarse": 16, "sao": 0, "intra_in_p": 1, "intra_satd": 0},
    "veryfast": {"search_range": 6, "me_coarse": 16, "sao": 2, "intra_in_p": 1, "intra_satd": 1},
    "medium": {"search_range": 12, "me_coarse": 32, "sao": 2, "intra_in_p": 1, "intra_satd": 1},
}


def frame(kind, w, h, t):
    if kind == "screen":
        return synth.screen_i420(w, h, t * 5)
    if kind == "sports":
        return synth.sports_i420(w, h, t)
    return synth.camera_i420(w, h, t)


def rd_points(kind, w, h, n, period, **kw):
    pts = []
    for qp in (22, 27, 32, 37):
        enc = OracleEncoder(w, h, qp=qp, intra_period=period, **kw)
        bits, psnr = 0, 0.0
        for t in range(n):
            f = frame(kind, w, h, t)
            bits += 8 * len(enc.encode(f))
            psnr += synth.psnr(f[:w * h], enc.recon()[:w * h])
        enc.close()
        pts.append((bits / n * 30 / 1000, psnr / n))
    return pts


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    oracle.load().orc_set_threads(8)
    print(json.dumps({"note": "tools/rd_presets.py %d -- BD-rate (Bjontegaard, PSNR-Y, QP 22/27/32/37) against the round-1 encoder; "
                              "negative = fewer bits at equal quality; oracle streams = GPU streams byte for byte" % n}))
    for kind, w, h, period in (("camera", 416, 240, 0), ("camera", 1280, 720, 0), ("sports", 416, 240, 0), ("sports", 1280, 720, 0),
                               ("screen", 640, 360, 0), ("camera", 416, 240, 8)):
        base = rd_points(kind, w, h, n, period, **ROUND1)
        out = {"sequence": f"{kind} {w}x{h}, {n} pictures, intra period {period or 'first only'}",
               "round1_kbps_psnr": [(round(r, 1), round(p, 2)) for r, p in base]}
        for name, kw in PRESETS.items():
            pts = rd_points(kind, w, h, n, period, **kw)
            out["bd_rate_%_" + name] = round(bd_rate(base, pts), 2)
            if name == "veryfast":
                out["veryfast_kbps_psnr"] = [(round(r, 1), round(p, 2)) for r, p in pts]
        print(json.dumps(out), flush=True)
