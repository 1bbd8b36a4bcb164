"""Encode a few 1080p pictures in the bench configuration (for ncu captures)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder, preset_options
w, h = 1920, 1080
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
enc = GpuEncoder(w, h, qp=27, intra_period=64, fps_num=30, fps_den=1, **preset_options("veryfast"))
for t in range(n):
    d = torch.from_numpy(synth.camera_i420(w, h, t)).cuda()
    torch.cuda.synchronize()
    print(t, len(enc.encode_dev(d)))
