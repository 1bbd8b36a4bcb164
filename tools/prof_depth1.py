"""Per-kernel device time at depth 1 (nothing else on the GPU): 1080p QP27, IDR + P pictures."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
frames = [torch.from_numpy(synth.camera_i420(W, H, t)).cuda() for t in range(12)]
torch.cuda.synchronize()
for period in (1, 0):
    enc = GpuEncoder(W, H, qp=27, intra_period=period, search_range=12, depth=1)
    for i in range(3): enc.encode_dev(frames[i])
    enc.set_profile(True)
    for i in range(3, 12): enc.encode_dev(frames[i])
    print("all-intra" if period == 1 else "P pictures", {k: round(v[0] / max(v[1], 1), 3) for k, v in enc.profile().items() if v[1]}, flush=True)
    enc.close()
