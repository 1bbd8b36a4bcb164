"""What the optional tools do to PSNR-based rate-distortion (they are not meant to improve it): BD-rate over PSNR-Y at
QP 22...37 against the veryfast preset, on the CPU oracle (= the GPU encoder, byte for byte).  VAQ and scaling lists are
psychovisual tools; the frame motion constraint trades efficiency at the picture edges for self-contained pictures.

  python tools/rd_tools.py  -> JSON lines (profiles/r02_rd_tools.jsonl)
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle
from kvazzup_b200 import synth
from oracle.encoder import OracleEncoder
from tools.rd_oracle_options import bd_rate
from tools.rd_presets import frame
from tests.test_oracle_hevc import vaq_frames
oracle.load().orc_set_threads(8)
VF = {"search_range": 6, "me_coarse": 16, "sao": 2, "intra_in_p": 1, "intra_satd": 1}
def rd(kind, w, h, n, mixed, **kw):
    pts = []
    fr = vaq_frames(kind, w, h, n) if mixed else [frame(kind, w, h, t) for t in range(n)]
    for qp in (22, 27, 32, 37):
        enc = OracleEncoder(w, h, qp=qp, intra_period=0, **(VF | kw))
        bits, psnr = 0, 0.0
        for f in fr:
            bits += 8 * len(enc.encode(f))
            psnr += synth.psnr(f[:w * h], enc.recon()[:w * h])
        enc.close()
        pts.append((bits / n * 30 / 1000, psnr / n))
    return pts
n = 13
for kind, w, h, mixed in (("camera", 416, 240, 0), ("camera", 416, 240, 1), ("camera", 1280, 720, 1), ("screen", 640, 360, 0)):
    base = rd(kind, w, h, n, mixed)
    out = {"sequence": f"{kind} {w}x{h}" + (" with flat and noisy thirds" if mixed else ""), "pictures": n}
    out["bd_rate_psnr_y_%_vaq10"] = round(bd_rate(base, rd(kind, w, h, n, mixed, qp_delta=1, vaq=10)), 2)
    out["bd_rate_psnr_y_%_scaling_list_default"] = round(bd_rate(base, rd(kind, w, h, n, mixed, scaling_list=1)), 2)
    out["bd_rate_psnr_y_%_mv_constraint_frame"] = round(bd_rate(base, rd(kind, w, h, n, mixed, mv_edges=15)), 2)
    print(json.dumps(out), flush=True)
