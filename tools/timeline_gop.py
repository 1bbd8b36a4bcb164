"""Device timeline around an IDR (period 64, depth 96): where does the GOP time go?"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder
w, h = 1920, 1080
frames = [torch.from_numpy(synth.camera_i420(w, h, t)).cuda() for t in range(64)]
torch.cuda.synchronize()
depth = 96
e = GpuEncoder(w, h, qp=27, intra_period=64, search_range=12, depth=depth)
e.set_profile(True)
rows = []
for i in range(64 * 5):
    au = e.encode_dev(frames[i % 64])
    if au:
        rows.append((i - depth + 1, e.timeline()))
while e.pending():
    e.flush()
    rows.append((rows[-1][0] + 1, e.timeline()))
prev_end = None
for n, tl in rows:
    if n < 120 or n > 200:
        continue
    if n % 64 == 0:
        print(f"pic {n:4d} IDR  intra {tl['intra'][0]:8.2f}-{tl['intra'][1]:8.2f}  deblock -{tl['deblock'][1]:8.2f} arith {tl['arith'][0]:8.2f}-{tl['arith'][1]:8.2f}")
    elif n % 64 in (1, 2, 3, 62, 63) or n % 8 == 0:
        gap = tl['me'][0] - prev_end if prev_end is not None else 0
        print(f"pic {n:4d} P    me {tl['me'][0]:8.2f}-{tl['me'][1]:8.2f} recon -{tl['recon'][1]:8.2f} deblock -{tl['deblock'][1]:8.2f} (gap before me {gap:6.3f}) binarise {tl['binarise'][0]:8.2f}-{tl['binarise'][1]:8.2f} arith -{tl['arith'][1]:8.2f}")
    if n % 64 != 0:
        prev_end = tl['deblock'][1]
