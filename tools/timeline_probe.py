"""Print the device timeline of consecutive pictures (depth 16, P only)."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder
w, h = 1920, 1080
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
frames = [torch.from_numpy(synth.camera_i420(w, h, t)).cuda() for t in range(16)]
torch.cuda.synchronize()
e = GpuEncoder(w, h, qp=27, intra_period=0, search_range=12, depth=depth)
e.set_profile(True)
rows = []
for i in range(depth * 4):
    t0 = time.perf_counter()
    au = e.encode_dev(frames[i % 16])
    dt = (time.perf_counter() - t0) * 1e3
    if au:
        rows.append((i, dt, e.timeline()))
for i, dt, tl in rows[-24:]:
    print(f"call {i:3d} host {dt:6.3f} ms | me {tl['me'][0]:8.2f}-{tl['me'][1]:8.2f} recon -{tl['recon'][1]:8.2f} deblock {tl['deblock'][0]:8.2f}-{tl['deblock'][1]:8.2f} "
          f"| binarise {tl['binarise'][0]:8.2f}-{tl['binarise'][1]:8.2f} arith {tl['arith'][0]:8.2f}-{tl['arith'][1]:8.2f} pack -{tl['pack'][1]:8.2f}")
