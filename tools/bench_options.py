"""What the optional encoder tools cost: 1080p veryfast QP 27, pictures resident in HBM, 32 in flight, GOP 64 --
the shape of bench.py's headline -- with variance adaptive quantisation, the default scaling lists, per-CTU QP alone
(ROI), the frame motion constraint, and a 1366x768 source (padding + conformance window).

  python tools/bench_options.py [pictures]  -> JSON lines (profiles/r02_bench_options.jsonl)
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kvazzup_b200 import synth  # noqa: E402
from kvazzup_b200.encoder import GpuEncoder, preset_options  # noqa: E402


def run(name, w, h, n, src=None, **opts):
    sw, sh = src or (w, h)
    pics = [torch.from_numpy(np.ascontiguousarray(synth.camera_i420(sw, sh, t))).cuda() for t in range(16)]
    kw = preset_options("veryfast") | opts
    if src:
        kw |= {"src_width": sw, "src_height": sh}
    enc = GpuEncoder(w, h, qp=27, intra_period=64, depth=32, **kw)
    total = 0
    for t in range(48):                                   # warm-up: fill the pipeline
        total += len(enc.encode_dev(pics[t % 16]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(n):
        total += len(enc.encode_dev(pics[(t + 48) % 16]))
    while enc.pending():
        total += len(enc.flush())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    enc.close()
    print(json.dumps({"case": name, "size": f"{sw}x{sh}", "pictures": n, "pictures_per_s": round(n / dt, 1),
                      "options": opts}), flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    print(json.dumps({"note": "tools/bench_options.py: wall-clock over n pictures after a 48-picture warm-up, depth 32, "
                              "device-resident pictures, one B200; includes draining the pipeline"}))
    run("veryfast (reference point)", 1920, 1080, n)
    run("+ cu_qp_delta, no offsets (ROI path)", 1920, 1080, n, qp_delta=1)
    run("+ vaq 10", 1920, 1080, n, qp_delta=1, vaq=10)
    run("+ scaling-list default", 1920, 1080, n, scaling_list=1)
    run("+ mv-constraint frame", 1920, 1080, n, mv_edges=15)
    run("1366x768 source (coded 1368x768, padded on the GPU)", 1368, 768, n, src=(1366, 768))
    run("1368x768 source (no padding)", 1368, 768, n)
