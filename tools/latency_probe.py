"""Per-picture latency with nothing in flight (owf 0 / slice threading: what a live call sees):
encode call -> access unit, decode call -> picture, for untiled and tiled streams."""
import json, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder, GpuTiledEncoder
from kvazzup_b200.openhevc import OpenHEVCFilter, split_nals

for (w, h, qp) in ((1280, 720, 32), (1920, 1080, 27), (1920, 1080, 37)):
    frames = [synth.camera_i420(w, h, t) for t in range(24)]
    for tiles, wpp in ((1, 1), (4, 1)):
        enc = GpuEncoder(w, h, qp=qp, intra_period=12, search_range=12) if tiles == 1 else \
            GpuTiledEncoder(w, h, tiles, qp=qp, intra_period=12, search_range=12, wpp=wpp)
        dec = OpenHEVCFilter(); dec.init()
        te, td, sizes, idr = [], [], [], []
        for i, f in enumerate(frames):
            t0 = time.perf_counter(); au = enc.encode(f); t1 = time.perf_counter()
            nals = split_nals(au)
            t2 = time.perf_counter()
            for nal in nals:
                dec.process(nal)
            t3 = time.perf_counter()
            te.append((t1 - t0) * 1e3); td.append((t3 - t2) * 1e3); sizes.append(len(au)); idr.append(i % 12 == 0)
        enc.close(); dec.close()
        sel = lambda xs, flag: [x for x, k in zip(xs[12:], idr[12:]) if k == flag]
        print(json.dumps({"size": f"{w}x{h}", "qp": qp, "tile_columns": tiles,
                          "idr_bytes": int(np.mean(sel(sizes, True))), "p_bytes": int(np.mean(sel(sizes, False))),
                          "encode_ms": {"idr": round(float(np.mean(sel(te, True))), 2), "p": round(float(np.median(sel(te, False))), 2)},
                          "decode_ms": {"idr": round(float(np.mean(sel(td, True))), 2), "p": round(float(np.median(sel(td, False))), 2)}}), flush=True)
