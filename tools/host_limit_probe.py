"""Host-side cost per picture: encode a tiny picture so that GPU work is negligible."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder
for (w, h) in ((64, 64), (640, 480)):
    fr = [torch.from_numpy(synth.camera_i420(w, h, t)).cuda() for t in range(8)]
    torch.cuda.synchronize()
    for prof in (False, True):
        e = GpuEncoder(w, h, qp=27, intra_period=0, search_range=12, depth=48)
        e.set_profile(prof)
        for i in range(100):
            e.encode_dev(fr[i % 8])
        t0 = time.perf_counter()
        n = 1000
        for i in range(n):
            e.encode_dev(fr[i % 8])
        while e.pending():
            e.flush()
        dt = time.perf_counter() - t0
        print(f"{w}x{h} profile={prof}: {n/dt:.0f} calls/s ({dt/n*1e6:.0f} us per picture)")
        e.close()
