"""Where the host time of the kvz_api (e2e) path goes at 1080p: full mirror, mirror without the
three plane copies, and the raw C call with a pre-filled pinned picture."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ctypes as C
import numpy as np
import torch
from kvazzup_b200 import synth
from kvazzup_b200.kvazaar import KvazaarFilter
W, H, GOP, DEPTH = 1920, 1080, 64, 96
frames = [synth.camera_i420(W, H, t) for t in range(GOP)]
for mode in ("full", "nocopy", "memcpy_only"):
    f = KvazaarFilter({"video/ResolutionWidth": W, "video/ResolutionHeight": H, "video/Preset": "veryfast", "video/QP": 27,
                       "video/Intra": GOP, "video/OWF": DEPTH - 1})
    assert f.init()
    def step():
        n = 0
        for fr in frames:
            if mode == "full":
                n += sum(map(len, f.feed_input(fr, drain=False)))
            elif mode == "nocopy":
                pic = f.input_pics[f.next_input_pic]
                f.next_input_pic = (f.next_input_pic + 1) % len(f.input_pics)
                n += sum(map(len, f._drain(pic, False)))
            else:
                pic = f.input_pics[f.next_input_pic]
                f.next_input_pic = (f.next_input_pic + 1) % len(f.input_pics)
                C.memmove(pic.contents.y, fr.ctypes.data, W * H * 3 // 2)
        return n
    for _ in range(3): step()
    f.flush()
    t0 = time.perf_counter()
    steps = 10
    for _ in range(steps): step()
    f.flush(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{mode}: {steps*GOP/dt:.0f} pictures/s ({dt/(steps*GOP)*1e6:.0f} us per picture)", flush=True)
    f.close()
