"""BASELINE config 4: N concurrent 720p30 participant streams on one GPU, each with its own encoder
(kvz_api engine) and decoder (libOpenHevc* ABI), one host thread per stream (the reference runs one
QThread per filter).  Reports aggregate pictures/s and the number of 30 fps streams that sustains.

  python tools/bench_conference.py [streams] [frames_per_stream] [encode_only]
"""
import os
import sys
import threading
import time
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import json  # noqa: E402

import numpy as np  # noqa: E402

from kvazzup_b200 import synth  # noqa: E402
from kvazzup_b200.encoder import GpuEncoder  # noqa: E402
from kvazzup_b200.openhevc import OpenHEVCFilter, split_nals  # noqa: E402

W, H = 1280, 720
n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 90
encode_only = len(sys.argv) > 3 and sys.argv[3] == "1"
frames = [synth.camera_i420(W, H, t) for t in range(30)]
barrier = threading.Barrier(n_streams + 1)
ok = [True] * n_streams


def worker(sid):
    enc = GpuEncoder(W, H, qp=32, intra_period=64, search_range=12, depth=4)
    dec = None
    if not encode_only:
        dec = OpenHEVCFilter()
        dec.init()
    # warm-up
    for t in range(6):
        au = enc.encode(frames[(t + sid) % 30])
        if au and dec:
            for nal in split_nals(au):
                dec.process(nal)
    barrier.wait()
    decoded = 0
    for t in range(n_frames):
        au = enc.encode(frames[(t + 6 + sid) % 30])
        if au and dec:
            for nal in split_nals(au):
                if dec.process(nal) is not None:
                    decoded += 1
    while enc.pending():
        au = enc.flush()
        if au and dec:
            for nal in split_nals(au):
                if dec.process(nal) is not None:
                    decoded += 1
    ok[sid] = encode_only or decoded == n_frames + 3      # + the depth-1 pictures that were in flight after the warm-up
    barrier.wait()
    enc.close()
    if dec:
        dec.close()


threads = [threading.Thread(target=worker, args=(i,)) for i in range(n_streams)]
for t in threads:
    t.start()
barrier.wait()
t0 = time.perf_counter()
barrier.wait()
dt = time.perf_counter() - t0
for t in threads:
    t.join()
fps = n_streams * n_frames / dt
print(json.dumps({"workload": "conference 720p30, QP32, encode" + ("" if encode_only else "+decode") + " per stream",
                  "streams": n_streams, "frames_per_stream": n_frames, "seconds": round(dt, 3),
                  "aggregate_fps": round(fps, 1), "fps_per_stream": round(fps / n_streams, 1),
                  "streams_sustained_at_30fps": int(fps // 30), "all_pictures_decoded": all(ok)}))
