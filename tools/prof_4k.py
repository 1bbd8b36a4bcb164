"""Per-kernel device time of the untiled encoder at 2160p (event profile), depth 1 and 32."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder
W, H = 3840, 2160
frames = [torch.from_numpy(synth.camera_i420(W, H, t)).cuda() for t in range(8)]
torch.cuda.synchronize()
for depth in (1, 32):
    enc = GpuEncoder(W, H, qp=27, intra_period=64, search_range=12, depth=depth)
    enc.set_profile(True)
    t0 = time.perf_counter()
    n = 40
    for i in range(n):
        enc.encode_dev(frames[i % 8])
    while enc.pending():
        enc.flush()
    dt = time.perf_counter() - t0
    print(f"depth {depth}: {n/dt:.1f} fps", {k: (round(v[0] / max(v[1], 1), 3), v[1]) for k, v in enc.profile().items()}, flush=True)
    enc.close()
