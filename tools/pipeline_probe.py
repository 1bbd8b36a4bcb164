"""Where does the time of a pipelined encode go?  P-only vs GOP, depth sweep, host-side cost per call."""
import os
import sys
import time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", sys.argv[1] if len(sys.argv) > 1 else "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder

w, h = 1920, 1080
frames = [torch.from_numpy(synth.camera_i420(w, h, t)).cuda() for t in range(32)]
torch.cuda.synchronize()


def run(period, depth, n, sr=12, qp=27):
    e = GpuEncoder(w, h, qp=qp, intra_period=period, search_range=sr, depth=depth)
    for i in range(depth + 4):
        e.encode_dev(frames[i % 32])
    while e.pending():
        e.flush()
    torch.cuda.synchronize()
    e.set_profile(True)
    t0 = time.perf_counter()
    host = 0.0
    for i in range(n):
        t1 = time.perf_counter()
        e.encode_dev(frames[i % 32])
        host += time.perf_counter() - t1
    t_sub = time.perf_counter() - t0
    while e.pending():
        e.flush()
    dt = time.perf_counter() - t0
    prof = e.profile()
    ks = " ".join(f"{k}={v[0]/max(v[1],1)*1e3:.0f}us" for k, v in prof.items() if v[1])
    print(f"period={period:3d} depth={depth:3d} n={n}: {n/dt:8.1f} fps  ({dt/n*1e3:.3f} ms/frame; submit loop {t_sub/n*1e3:.3f} ms/frame) {ks}")
    e.close()


run(0, 48, 256)
run(0, 16, 256)
run(0, 4, 128)
run(0, 1, 64)
run(64, 48, 256)
