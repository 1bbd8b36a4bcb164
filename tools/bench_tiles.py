"""BASELINE config 3: ONE 3840x2160 stream, tile columns split across the GPUs of the box (strong
scaling, no exchange between the GPUs -- SURVEY.md 8e option A).  One process drives all GPUs.

  python tools/bench_tiles.py [max_gpus] [tiles] [frames]
"""
import json
import os
import sys
import time
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kvazzup_b200 import synth  # noqa: E402
from kvazzup_b200.capi import lib  # noqa: E402
from kvazzup_b200.encoder import GpuEncoder, GpuTiledEncoder  # noqa: E402

W, H = 3840, 2160
max_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_frames = int(sys.argv[3]) if len(sys.argv) > 3 else 96
QP, GOP, ME = 27, 64, 12
DEPTH = int(os.environ.get("B200_TILE_DEPTH", "32"))
src = [synth.camera_i420(W, H, t) for t in range(16)]
# page-locked copies: the strips are uploaded straight from the caller's picture
pinned = [torch.from_numpy(f).pin_memory() for f in src]
frames = [p.numpy() for p in pinned]
frames = frames + frames[-2:0:-1]
for d in range(torch.cuda.device_count()):
    torch.cuda.synchronize(d)                # create every device context before anything is timed          # ping-pong order: no scene cut when the sequence wraps


def run(make):
    enc = make()
    for i in range(DEPTH + 8):
        enc.encode(frames[i % len(frames)])
    while enc.pending():
        enc.flush()
    t0 = time.perf_counter()
    nbytes = 0
    for i in range(n_frames):
        nbytes += len(enc.encode(frames[i % len(frames)]))
    while enc.pending():
        nbytes += len(enc.flush())
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    enc.close()
    return n_frames / dt, nbytes * 8 / n_frames * 30 / 1000


fps, kbps = run(lambda: GpuEncoder(W, H, qp=QP, intra_period=GOP, search_range=ME, depth=DEPTH))
print(json.dumps({"workload": "2160p QP27 GOP64, host pictures", "mode": "untiled, WPP, 1 GPU", "fps": round(fps, 1), "kbps_at_30fps": round(kbps, 1)}), flush=True)
print(json.dumps({"pictures_in_flight": DEPTH, "gop": GOP, "qp": QP, "me_range": ME}), flush=True)
for wpp in ((1,) if os.environ.get("B200_TILE_WPP_ONLY") else (1, 0)):
    g = 1
    while g <= max_gpus:
        if tiles % g == 0 or g == 1:
            fps, kbps = run(lambda: GpuTiledEncoder(W, H, tiles, qp=QP, intra_period=GOP, search_range=ME, depth=DEPTH, wpp=wpp,
                                                    devices=tuple(range(g))))
            print(json.dumps({"workload": "2160p QP27 GOP64, host pictures", "mode": f"{tiles} tile columns, {'WPP rows' if wpp else 'one substream'} per tile",
                              "gpus": g, "fps": round(fps, 1), "kbps_at_30fps": round(kbps, 1), "launches": int(lib().b200_launch_count())}), flush=True)
        g *= 2
