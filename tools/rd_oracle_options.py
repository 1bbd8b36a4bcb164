"""Rate-distortion effect of the oracle-only tools (CPU, no GPU): BD-rate of SATD-based fractional
refinement and of SAO against the encoder the GPU implements, on the synthetic sequences.
Bjontegaard delta rate from a cubic fit of log-rate over PSNR-Y at QP 22/27/32/37.

  python tools/rd_oracle_options.py [frames]
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np  # noqa: E402

from kvazzup_b200 import synth  # noqa: E402
from oracle.encoder import OracleEncoder  # noqa: E402


def rd_points(kind, w, h, n, **kw):
    pts = []
    for qp in (22, 27, 32, 37):
        enc = OracleEncoder(w, h, qp=qp, intra_period=0, search_range=12, **kw)
        bits, psnr = 0, 0.0
        for t in range(n):
            f = synth.camera_i420(w, h, t) if kind == "camera" else synth.screen_i420(w, h, t * 5)
            bits += 8 * len(enc.encode(f))
            psnr += synth.psnr(f[:w * h], enc.recon()[:w * h])
        enc.close()
        pts.append((bits / n * 30 / 1000, psnr / n))
    return pts


def bd_rate(ref, test):
    """Average bitrate difference (%) of `test` against `ref` at equal PSNR."""
    lr1, p1 = np.log([r for r, _ in ref]), np.array([p for _, p in ref])
    lr2, p2 = np.log([r for r, _ in test]), np.array([p for _, p in test])
    f1, f2 = np.polyfit(p1, lr1, 3), np.polyfit(p2, lr2, 3)
    lo, hi = max(p1.min(), p2.min()), min(p1.max(), p2.max())
    i1, i2 = np.polyint(f1), np.polyint(f2)
    a1 = (np.polyval(i1, hi) - np.polyval(i1, lo)) / (hi - lo)
    a2 = (np.polyval(i2, hi) - np.polyval(i2, lo)) / (hi - lo)
    return float((np.exp(a2 - a1) - 1) * 100)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    for kind, w, h in (("camera", 416, 240), ("camera", 1280, 720), ("screen", 640, 360)):
        base = rd_points(kind, w, h, n)
        out = {"sequence": f"{kind} {w}x{h}, {n} pictures, IPPP", "base_kbps_psnr": [(round(r, 1), round(p, 2)) for r, p in base]}
        for name, kw in (("subme_satd", {"subme_satd": 1}), ("sao", {"sao": 1}), ("subme_satd+sao", {"subme_satd": 1, "sao": 1})):
            out["bd_rate_%_" + name] = round(bd_rate(base, rd_points(kind, w, h, n, **kw)), 2)
        print(json.dumps(out), flush=True)
