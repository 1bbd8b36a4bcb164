#!/usr/bin/env python
"""Benchmark of the B200 HEVC media path on BASELINE.json's metric.

  python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm, see below)

metric   1080p HEVC encode fps (BASELINE.json: "1080p HEVC encode fps/GPU"), whole job, all ranks
config   configs[1]: 1080p30 8-bit, veryfast preset, low-delay P (each picture references the
         previous reconstruction), IDR period 64, constant QP 27 (headline; the QP 22/27/32/37
         sweep is reported in `qp_sweep`), one stream per GPU, synthetic `camera` sequence
step     one GOP: 64 consecutive pictures (1 IDR + 63 P) of that stream
value    pictures/s with the 64 source pictures already resident in HBM (encoder engine, pipelined)
e2e      pictures/s through the reference-facing C ABI (kvz_api: picture_alloc / encoder_encode /
         chunk list) with HOST I420 buffers: the host->device copy of every picture and the
         device->host read of every access unit are inside the timed region
roofline per-kernel device time measured live with CUDA events on the launching streams; the entry
         is the kernel with the largest share of device time
cpu_baseline / --impl reference
         Kvazaar is not in the reference tree nor on this image, so the CPU arm is
         the in-house oracle port of the same encoder (oracle/, kind "port"), OpenMP over CTUs on
         the box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# one hardware work queue per stream: must be set before the CUDA context exists (see runtime.cu)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

W, H = 1920, 1080
GOP = 64
QP = 27
PRESET = "veryfast"
ME_RANGE = 12                      # what "veryfast" maps to (kvz_api.cu kPresets)
DEPTH = 96                         # pictures in flight (owf = 95): an IDR's serial entropy coding (~30 ms) overlaps the next GOP's prediction chain
KERNELS = ("intra", "me", "recon", "modes", "deblock", "binarise", "arith", "pack")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def peaks():
    try:
        d = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        # NVML first: a query takes microseconds and the first sample lands at once (the timed region
        # of a default run is well under a second); nvidia-smi -lms as the fallback of the recipe.
        if self._run_nvml():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            if self.stop_flag.is_set():
                break
            parts = [p.strip() for p in line.split(",")]
            try:
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                continue

    def _run_nvml(self) -> bool:
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu < len(ids) and ids[self.gpu].isdigit():
                    idx = int(ids[self.gpu])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            while not self.stop_flag.is_set():
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for b, n in bits.items():
                    if r & b:
                        self.reasons.add(n)
                time.sleep(0.02)
            return True
        except Exception:
            return bool(self.samples)

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_source(n_frames: int, rank: int):
    from kvazzup_b200 import synth
    # every rank encodes its own participant stream: same generator, different time origin
    return [synth.camera_i420(W, H, t + 1000 * rank) for t in range(n_frames)]


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist

    import kvazzup_b200
    from kvazzup_b200.encoder import GpuEncoder
    from kvazzup_b200.kvazaar import KvazaarFilter

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 media path has no CPU fallback")
    torch.cuda.set_device(local)
    lib = kvazzup_b200.lib()
    lib.b200_set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frames = make_source(GOP, rank)
    d_frames = [torch.from_numpy(f).cuda() for f in frames]
    torch.cuda.synchronize()
    frame_bytes = W * H * 3 // 2

    # ---- value: engine, source pictures resident in HBM ----
    enc = GpuEncoder(W, H, qp=QP, intra_period=GOP, search_range=ME_RANGE, depth=DEPTH)
    out_bytes = [0]

    # A step submits one GOP (64 pictures); the pipeline stays full across steps (like a live
    # stream) and is drained once, inside the timed region, after the last step.
    def step_engine(e):
        n = 0
        for d in d_frames:
            au = e.encode_dev(d)
            n += len(au)
        return n

    def drain(e):
        n = 0
        while e.pending():
            n += len(e.flush())
        return n

    for _ in range(args.warmup):
        step_engine(enc)
    drain(enc)
    enc.set_profile(True)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.b200_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    produced = 0
    for _ in range(args.steps):
        produced += step_engine(enc)
    produced += drain(enc)
    out_bytes[0] = produced / args.steps
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_s = e0.elapsed_time(e1) * 1e-3
    launches = lib.b200_launch_count() - launches0
    prof = enc.profile()
    enc.set_profile(False)
    clocks = sampler.finish()
    elapsed = max(dev_s, 1e-9)
    if world > 1:
        t = torch.tensor([elapsed], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    total_frames = GOP * args.steps * world
    value = total_frames / elapsed
    bitrate_kbps = out_bytes[0] * 8 / GOP * 30 / 1000
    enc.close()

    # ---- e2e: kvz_api with host buffers ----
    filt = KvazaarFilter({"video/ResolutionWidth": W, "video/ResolutionHeight": H, "video/Preset": PRESET, "video/QP": QP,
                          "video/Intra": GOP, "video/OWF": DEPTH - 1, "video/FramerateNumerator": 30})
    if not filt.init():
        raise SystemExit("KvazaarFilter.init failed: " + lib.b200_last_error().decode())
    d2h = [0]

    def step_e2e():
        n = 0
        for f in frames:
            for au in filt.feed_input(f, drain=False):
                n += len(au)
        return n

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    filt.flush()
    barrier()
    e2e_steps = max(1, args.steps)
    t0 = time.perf_counter()
    nb = 0
    for _ in range(e2e_steps):
        nb += step_e2e()
    nb += sum(len(a) for a in filt.flush())
    d2h[0] = nb // e2e_steps
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = GOP * e2e_steps * world / e2e_s
    filt.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    hbm_peak, peak_src = peaks()
    px = W * H
    # algorithmic bytes per launch (DESIGN.md "Kernels"): every datum moved once
    alg = {
        "intra": px * 1.5 + px * 1.5 + px * 3.0,              # source read, reconstruction written, levels written
        "me": px * 1.0 + px * 1.0 + (px / 64) * 12,           # luma source + luma reference read, cu map written
        "recon": px * 1.5 * 2 + px * 1.5 + px * 3.0 + (px / 64) * 12,   # src+ref read, recon + levels written, cu map
        "modes": (px / 64) * 12 * 2,
        "deblock": 2 * (px * 1.0 * 2),                        # two passes, luma read + written
        "binarise": px * 3.0 + (px / 64) * 12 + 4.0 * 8 * out_bytes[0] / GOP,   # levels + cu map read, ~1 record (4 B) per bin written
        "arith": 12.0 * 8 * out_bytes[0] / GOP + out_bytes[0] / GOP,             # records: phase A reads + rewrites, phase B reads; bitstream written
        "pack": 2.0 * out_bytes[0] / GOP,
    }
    total_ms = sum(v[0] for v in prof.values()) or 1.0
    kernels = {}
    for k in KERNELS:
        ms, cnt = prof[k]
        if not cnt:
            continue
        avg = ms / cnt
        ach = alg[k] / (avg * 1e-3) / 1e9
        kernels[k] = {"launches": cnt, "avg_us": round(avg * 1e3, 2), "share": round(ms / total_ms, 4),
                      "achieved_gbs": round(ach, 2), "frac": round(ach / hbm_peak, 5)}
    top = max(kernels, key=lambda k: kernels[k]["share"])
    traffic = ncu_traffic()
    for k in kernels:
        kernels[k]["ncu_dram_bytes"] = traffic.get(k)
    roofline = {"kernel": top, "bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": hbm_peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": kernels[top]["frac"], "traffic": traffic.get(top),
                "traffic_source": "profiles/r01_ncu_launch_summary.csv (dram__bytes_read+write per launch, same command under ncu)",
                "algorithmic_bytes_per_launch": int(alg[top]), "avg_launch_us": kernels[top]["avg_us"],
                "note": "integer / latency bound kernels: see profiles/ for the ncu pipe utilisation; fraction of HBM "
                        "roofline is reported because the contract asks for it",
                "kernels": kernels}

    cpu = cpu_baseline_sample(frames, max_seconds=25.0, threads=1)

    line = {
        "metric": "1080p HEVC encode fps", "value": round(value, 2), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(elapsed / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "1080p30 veryfast low-delay-P QP27, GOP 64, one stream per GPU (BASELINE configs[1])",
                   "width": W, "height": H, "qp": QP, "preset": PRESET, "me_range": ME_RANGE, "gop": GOP,
                   "frames_per_step": GOP, "pictures_in_flight": DEPTH, "streams": world,
                   "l2_policy": "inputs larger than L2: 64 distinct 3.1 MB pictures (199 MB) cycled per step",
                   "bitrate_kbps_at_30fps": round(bitrate_kbps, 1)},
        "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "h2d_bytes_per_step": frame_bytes * GOP,
                "d2h_bytes_per_step": d2h[0], "api": "kvz_api (picture_alloc/encoder_encode/chunk_free), host I420 buffers"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wall_s": round(wall, 3),
    }
    if args.sweep:
        line["qp_sweep"] = qp_sweep(d_frames)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic():
    """DRAM bytes per launch of each kernel from the committed ncu launch summary (None if absent)."""
    names = {"k_intra_frame<0>": "intra", "k_intra_modes": "intra_modes", "k_me_ctu": "me", "k_inter_recon<0>": "recon",
             "k_inter_modes": "modes", "k_deblock": "deblock", "k_binarise": "binarise", "k_arith_rows": "arith",
             "k_pack_rows": "pack"}
    out = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_launch_summary.csv")) as f:
            next(f)
            for ln in f:
                c = ln.strip().split(",")
                if c[0] in names:
                    out[names[c[0]]] = int((float(c[5]) + float(c[6])) * 1e6)
    except (OSError, ValueError, IndexError):
        pass
    return out


def qp_sweep(d_frames):
    """fps / bitrate / PSNR-Y at the four QPs of BASELINE configs[1] (outside the timed region)."""
    import torch

    from kvazzup_b200 import synth
    from kvazzup_b200.encoder import GpuEncoder
    out = {}
    for qp in (22, 27, 32, 37):
        e = GpuEncoder(W, H, qp=qp, intra_period=GOP, search_range=ME_RANGE, depth=DEPTH)
        for d in d_frames[:8]:
            e.encode_dev(d)
        while e.pending():
            e.flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nbytes = 0
        for d in d_frames:
            nbytes += len(e.encode_dev(d))
        while e.pending():
            nbytes += len(e.flush())
        dt = time.perf_counter() - t0
        rec = e.recon()
        src = d_frames[-1].cpu().numpy()
        out[str(qp)] = {"fps": round(len(d_frames) / dt, 1), "kbps_at_30fps": round(nbytes * 8 / len(d_frames) * 30 / 1000, 1),
                        "psnr_y_last": round(synth.psnr(src[:W * H], rec[:W * H]), 2)}
        e.close()
    return out


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the checker, timed -- never the product)
# ---------------------------------------------------------------------------------------------------

def cpu_baseline_sample(frames, max_seconds, threads):
    """Encode a bounded sample (1 IDR + a few P pictures) of the same stream with the CPU oracle."""
    import oracle
    from oracle.encoder import OracleEncoder
    lib = oracle.load()
    lib.orc_set_threads(threads)
    enc = OracleEncoder(W, H, qp=QP, intra_period=GOP, search_range=ME_RANGE)
    t0 = time.perf_counter()
    n = 0
    for f in frames:
        enc.encode(f)
        n += 1
        if time.perf_counter() - t0 > max_seconds or n >= 8:
            break
    dt = time.perf_counter() - t0
    enc.close()
    lib.orc_set_threads(1)
    return {"value": round(n / dt, 4), "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"first {n} pictures (1 IDR + {n - 1} P) of the same 1080p stream, in-house oracle encoder "
                      f"(not Kvazaar: absent from the reference tree and this image)"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import oracle
    lib = oracle.load()
    threads = max(1, min(lib.orc_max_threads(), os.cpu_count() or 1))
    frames = make_source(6, 0)
    from oracle.encoder import OracleEncoder
    lib.orc_set_threads(threads)
    per_step = 3                                    # pictures per step: bounded sample of the GOP workload
    enc = OracleEncoder(W, H, qp=QP, intra_period=GOP, search_range=ME_RANGE)
    idx = 0

    def step():
        nonlocal idx
        for _ in range(per_step):
            enc.encode(frames[idx % len(frames)])
            idx += 1

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    enc.close()
    line = {
        "impl": "reference", "metric": "1080p HEVC encode fps", "value": round(value, 4), "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "1080p30 veryfast low-delay-P QP27, GOP 64, one stream (BASELINE configs[1])",
                   "width": W, "height": H, "qp": QP, "me_range": ME_RANGE, "gop": GOP, "frames_per_step": per_step},
        "cpu_baseline": {"value": round(value, 4), "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} pictures per step of the same 1080p stream; in-house oracle encoder, OpenMP over "
                                   f"CTUs, {threads} threads (Kvazaar itself is not available: SURVEY.md 8c)"},
        "e2e": {"value": round(value, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweep", action="store_true", help="also report the QP 22/27/32/37 sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
