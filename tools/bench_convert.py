"""Micro-benchmark of the conversion kernels (device-resident, batched, working set >> L2)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from kvazzup_b200 import convert  # noqa: E402
from kvazzup_b200.capi import FOURCC  # noqa: E402


def time_it(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    peak = 6453.7
    try:
        peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    except Exception:
        pass
    s = torch.cuda.current_stream().cuda_stream
    rows = []
    for (w, h) in ((1920, 1080), (1280, 720), (3840, 2160)):
        n = max(8, int(600e6 // (w * h * 5.5)))
        src = torch.randint(0, 256, (w * h * 3 // 2 * n,), dtype=torch.uint8, device="cuda")
        dst = torch.empty(w * h * 4 * n, dtype=torch.uint8, device="cuda")
        t = time_it(lambda: convert.i420_to_rgb32_dev(src, dst, w, h, n, s))
        gbs = 5.5 * w * h * n / t / 1e9
        rows.append(dict(kernel="i420_to_rgb32", w=w, h=h, n=n, us_per_frame=t / n * 1e6, gbs=gbs, frac=gbs / peak))
        del src, dst
    w, h = 1920, 1080
    for name, bpp in (("YUYV", 2), ("UYVY", 2), ("NV12", 1.5), ("I422", 2), ("ARGB", 4), ("24BG", 3)):
        n = 64
        src = torch.randint(0, 256, (int(w * h * bpp) * n,), dtype=torch.uint8, device="cuda")
        dst = torch.empty(w * h * 3 // 2 * n, dtype=torch.uint8, device="cuda")
        t = time_it(lambda: convert.convert_to_i420_dev(src, dst, w, h, FOURCC[name], n, s))
        gbs = (bpp + 1.5) * w * h * n / t / 1e9
        rows.append(dict(kernel=name + "_to_i420", w=w, h=h, n=n, us_per_frame=t / n * 1e6, gbs=gbs, frac=gbs / peak))
    n = 48
    src = torch.randint(0, 256, (w * h * 4 * n,), dtype=torch.uint8, device="cuda")
    dst = torch.empty(w * h * 4 * n, dtype=torch.uint8, device="cuda")
    t = time_it(lambda: convert.half_rgb_dev(src, dst, w, h, n, s))
    rows.append(dict(kernel="half_rgb", w=w, h=h, n=n, us_per_frame=t / n * 1e6, gbs=2 * w * h * n / t / 1e9,
                     frac=2 * w * h * n / t / 1e9 / peak))
    t = time_it(lambda: convert.flip_rgb_dev(src, dst, w, h, True, False, n, s))
    rows.append(dict(kernel="flip_rgb_h", w=w, h=h, n=n, us_per_frame=t / n * 1e6, gbs=8 * w * h * n / t / 1e9,
                     frac=8 * w * h * n / t / 1e9 / peak))
    # torch copy as a same-box reference for the achievable copy bandwidth
    a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    t = time_it(lambda: b.copy_(a))
    rows.append(dict(kernel="torch_copy_1GiB", gbs=2 * (1 << 30) / t / 1e9, frac=2 * (1 << 30) / t / 1e9 / peak))
    # MJPG: one frame per call (host JPEG -> device I420): Huffman decoding on the host, IDCT + resampling on the GPU
    try:
        import ctypes as C
        import time as _t

        import numpy as np

        import oracle
        from kvazzup_b200.capi import lib
        from tests import mjpg_util
        fn = lib().b200_mjpg_to_i420_dev
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        for (mw, mh, q) in ((640, 480, 80), (1280, 720, 80), (1920, 1080, 80)):
            jpeg = np.frombuffer(mjpg_util.make_jpeg(mw, mh, q, "422"), np.uint8)
            dst = torch.empty(mw * mh * 3 // 2, dtype=torch.uint8, device="cuda")
            for _ in range(5):
                fn(jpeg.ctypes.data, jpeg.size, dst.data_ptr(), mw, mh, s)
            torch.cuda.synchronize()
            t0 = _t.perf_counter()
            iters = 100
            for _ in range(iters):
                fn(jpeg.ctypes.data, jpeg.size, dst.data_ptr(), mw, mh, s)
            torch.cuda.synchronize()
            dt = (_t.perf_counter() - t0) / iters
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(jpeg.ctypes.data, jpeg.size, dst.data_ptr(), mw, mh, s)
            torch.cuda.synchronize()
            # the oracle (one host core) on the same frame
            t1 = _t.perf_counter()
            for _ in range(5):
                mjpg_util.oracle_mjpg_to_i420(oracle.load(), jpeg.tobytes(), mw, mh)
            cpu = (_t.perf_counter() - t1) / 5
            rows.append(dict(kernel="MJPG_to_i420 (4:2:2, quality %d, %d kB)" % (q, jpeg.size // 1000), w=mw, h=mh,
                             us_per_frame=dt * 1e6, frames_per_s=1 / dt, cpu_oracle_us_per_frame=cpu * 1e6))
    except Exception as e:                                  # cv2 (the test frames' encoder) missing
        rows.append(dict(kernel="MJPG_to_i420", skipped=str(e)))
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
