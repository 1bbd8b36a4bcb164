"""BASELINE configs 1 and 5 on one GPU (the other configs: bench.py, tools/bench_tiles.py,
bench.py --workload conference).

  config 1  uvgComm loopback: 640x480 YUYV -> I420 -> ultrafast QP32 encode -> NALs -> decode -> RGB32,
            as the filter graph of tools/loopback_pipeline.cpp (every filter on its own thread, owf 0:
            one picture in flight, the shape of a live call), and the same chain on the CPU oracle +
            FFmpeg's decoder for a few pictures as the reported CPU figure.
  config 5  1440p screen share: veryfast QP27 encode of text-like content (pictures resident in HBM, 32
            in flight), decode of that stream, I420 -> RGB32 of the decoded pictures.

  python tools/bench_configs.py  -> JSON lines (profiles/r02_configs_1_5.jsonl)
"""
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import kvazzup_b200  # noqa: E402
from kvazzup_b200 import convert, synth  # noqa: E402
from kvazzup_b200.capi import FOURCC  # noqa: E402
from kvazzup_b200.encoder import GpuEncoder, preset_options  # noqa: E402
from kvazzup_b200.openhevc import OpenHEVCFilter, split_nals  # noqa: E402


def config1():
    w, h, nfile, frames = 640, 480, 32, 600
    tmp = Path(tempfile.mkdtemp())
    cams = [synth.i420_to_yuyv(synth.camera_i420(w, h, t), w, h) for t in range(nfile)]
    np.concatenate(cams).tofile(tmp / "cam.yuyv")
    exe = str(tmp / "loopback_pipeline")
    kvazzup_b200.load()
    lib_dir = str(ROOT / "kvazzup_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + str(ROOT / "include"), str(ROOT / "tools/loopback_pipeline.cpp"), "-o", exe,
                    "-lpthread", "-L" + lib_dir, "-lb200media", "-Wl,-rpath," + lib_dir], check=True)
    for pace in ("0", "30"):
        n = frames if pace == "0" else 90
        r = subprocess.run([exe, str(tmp / "cam.yuyv"), str(w), str(h), str(nfile), str(n), "ultrafast", "32", pace],
                           capture_output=True, text=True, timeout=600)
        line = json.loads(r.stdout.strip().splitlines()[-1])
        line["config"] = "BASELINE configs[0]" + (" unpaced" if pace == "0" else " paced at 30 fps")
        print(json.dumps(line), flush=True)
    # the same chain on the CPU: oracle conversion + oracle encoder + FFmpeg decode + oracle RGB conversion
    import oracle
    from oracle.encoder import OracleEncoder
    from tests import ffhevc
    from tests.helpers import oracle_convert_to_i420, oracle_i420_to_rgb32
    olib = oracle.load()
    olib.orc_set_threads(1)
    enc = OracleEncoder(w, h, qp=32, intra_period=64, fps_num=30, fps_den=1, **preset_options("ultrafast"))
    dec = ffhevc.HevcDecoder(quiet=True) if ffhevc.available() else None
    n = 16
    t0 = time.perf_counter()
    for t in range(n):
        _, i420 = oracle_convert_to_i420(olib, cams[t], w, h, FOURCC["YUYV"])
        au = enc.encode(i420)
        pics = dec.decode(au) if dec else [(enc.recon(), w, h)]
        for p, _, _ in pics:
            oracle_i420_to_rgb32(olib, p, w, h)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "BASELINE configs[0] on the CPU", "fps": round(n / dt, 2), "cores": 1,
                      "chain": "oracle conversion + oracle encoder (in-house port, not Kvazaar) + " +
                               ("FFmpeg hevc decoder" if dec else "no decoder") + " + oracle RGB conversion", "pictures": n}), flush=True)


def config5():
    w, h, n, depth = 2560, 1440, 64, 32
    opts = preset_options("veryfast")
    src = [synth.screen_i420(w, h, t * 5) for t in range(16)]
    d_frames = [torch.from_numpy(f).cuda() for f in src]
    enc = GpuEncoder(w, h, qp=27, intra_period=64, depth=depth, **opts)
    aus = []
    for i in range(depth + 4):                                 # warm-up
        enc.encode_dev(d_frames[i % 16])
    while enc.pending():
        enc.flush()
    enc.close()
    enc = GpuEncoder(w, h, qp=27, intra_period=64, depth=depth, **opts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        au = enc.encode_dev(d_frames[i % 16])
        if au:
            aus.append(au)
    while enc.pending():
        aus.append(enc.flush())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    enc.close()
    kbps = sum(map(len, aus)) * 8 / n * 30 / 1000
    print(json.dumps({"config": "BASELINE configs[4]: 1440p screen share, veryfast QP27 GOP64, pictures resident in HBM, %d in flight" % depth,
                      "encode_fps": round(n / dt, 1), "kbps_at_30fps": round(kbps, 1), **opts}), flush=True)
    # decode + display conversion of that stream
    for threads in (1, 8):
        dec = OpenHEVCFilter(threads=threads, parallelization="Frame" if threads > 1 else "Slice")
        assert dec.init()
        shown = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for au in aus:
            for nal in split_nals(au):
                pic = dec.process(nal)
                if pic is not None:
                    convert.yuv420_to_rgb32(pic[0], w, h)
                    shown += 1
        dt = time.perf_counter() - t0
        dec.close()
        print(json.dumps({"config": "BASELINE configs[4]: decode + I420 -> RGB32 of that stream (host pictures both ways)",
                          "decoder_threads": threads, "pictures": shown, "fps": round(shown / dt, 1)}), flush=True)


if __name__ == "__main__":
    config1()
    config5()
