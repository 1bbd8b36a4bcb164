"""Quick timing of the encoder engine (device-resident frames)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kvazzup_b200 import synth  # noqa: E402
from kvazzup_b200.encoder import GpuEncoder  # noqa: E402

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
qp = int(sys.argv[4]) if len(sys.argv) > 4 else 27
sr = int(sys.argv[5]) if len(sys.argv) > 5 else 12
frames = [torch.from_numpy(synth.camera_i420(w, h, t)).cuda() for t in range(min(n, 16))]
torch.cuda.synchronize()
enc = GpuEncoder(w, h, qp=qp, intra_period=64, search_range=sr)
sizes = []
t0 = time.perf_counter()
for i in range(n):
    t1 = time.perf_counter()
    au = enc.encode_dev(frames[i % len(frames)])
    sizes.append((len(au), (time.perf_counter() - t1) * 1e3))
dt = time.perf_counter() - t0
print(f"{w}x{h} qp{qp} R{sr}: {n} frames in {dt*1e3:.1f} ms -> {n/dt:.1f} fps")
print("per-frame (bytes, ms):", [(s, round(ms, 2)) for s, ms in sizes[:12]])
print("bins last frame", enc.bins())
