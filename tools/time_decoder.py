"""Decode throughput of one stream (1080p / 720p), streams produced by the GPU encoder."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder
from kvazzup_b200.openhevc import OpenHEVCFilter, split_nals
for (w, h) in ((1920, 1080), (1280, 720)):
    n = 40
    enc = GpuEncoder(w, h, qp=27, intra_period=64, search_range=12)
    aus = [enc.encode(synth.camera_i420(w, h, t)) for t in range(n)]
    enc.close()
    nals = [split_nals(a) for a in aus]
    for threads in (8, 16, 32):
        dec = OpenHEVCFilter(threads, "Frame"); dec.init()
        reps = 4
        got = 0
        t0 = time.perf_counter()
        for r in range(reps):
            for ns in nals:
                for nal in ns:
                    got += dec.process(nal) is not None
        got += len(dec.drain())
        dt = time.perf_counter() - t0
        dec.close()
        assert got == reps * n, got
        print(f"{w}x{h}: frame threading {threads}: {reps*n/dt:.1f} pictures/s", flush=True)
    dec = OpenHEVCFilter(); dec.init()
    t_i = t_p = 0.0
    for i, ns in enumerate(nals):
        t0 = time.perf_counter()
        for nal in ns:
            dec.process(nal)
        dt = time.perf_counter() - t0
        if i == 0: t_i = dt
        elif i >= 4: t_p += dt
    dec.close()
    print(f"{w}x{h}: I picture {t_i*1e3:.2f} ms ({len(aus[0])} B), P pictures {t_p/(n-4)*1e3:.3f} ms each ({sum(map(len,aus[4:]))//(n-4)} B) -> {(n-4)/t_p:.1f} fps")
