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
PRESET = os.environ.get("B200_BENCH_PRESET", "veryfast")      # the headline is veryfast (BASELINE configs[1]); others for comparison
DEPTH = 96                         # pictures in flight (owf = 95): an IDR's serial entropy coding overlaps the next GOP's prediction chain
KERNELS = ("intra", "me", "recon", "modes", "deblock", "sao", "binarise", "arith", "pack")
CHAIN = ("me", "recon", "modes", "deblock", "sao")     # the P-picture prediction chain: what one picture must wait for
WAVEFRONT_STEPS = 62               # (cols - 1) + 2 (rows - 1) + 1 at 1080p: 1000 fps <=> 16.1 us per step


def engine_options():
    """What kvz_api's "veryfast" stands for (two-level motion search, SAO, intra CUs in P): the bare
    engine of `value` runs exactly the encoder `e2e` reaches through kvz_api."""
    from kvazzup_b200.encoder import preset_options
    return preset_options(PRESET)


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
    opts = engine_options()

    # ---- value: engine, source pictures resident in HBM ----
    enc = GpuEncoder(W, H, qp=QP, intra_period=GOP, depth=DEPTH, fps_num=30, fps_den=1, **opts)
    out_bytes = [0]
    last_gop = []

    # A step submits one GOP (64 pictures); the pipeline stays full across steps (like a live
    # stream) and is drained once, inside the timed region, after the last step.
    def step_engine(e, keep=None):
        n = 0
        for d in d_frames:
            au = e.encode_dev(d)
            n += len(au)
            if keep is not None and au:
                keep.append(au)
        return n

    def drain(e, keep=None):
        n = 0
        while e.pending():
            au = e.flush()
            n += len(au)
            if keep is not None and au:
                keep.append(au)
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
        produced += step_engine(enc, last_gop)
    produced += drain(enc, last_gop)
    out_bytes[0] = produced / args.steps
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_s = e0.elapsed_time(e1) * 1e-3
    launches = lib.b200_launch_count() - launches0
    prof = enc.profile()
    me_stats = enc.me_stats()
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
    last_gop = last_gop[-GOP:]                     # the access units of the last timed step (its GOP starts with an IDR)

    # ---- e2e: kvz_api with host buffers ----
    def kvz_run(owf, stock_drain, steps, warm, pinned_ring=False):
        """pictures/s through kvz_api.  stock_drain: the reference's own loop, which after every access
        unit keeps calling encoder_encode(NULL) until nothing comes back (kvazaarfilter.cpp:440-449) and
        so empties the pipeline; else the one-line patch of INTEGRATION.md section 1 (poll once).
        pinned_ring: the source pictures already sit in picture_alloc (page-locked) pictures -- a producer
        writing straight into the encoder's ring -- instead of being copied there plane by plane per
        picture as the reference's filter does (kvazaarfilter.cpp:410-418)."""
        f = KvazaarFilter({"video/ResolutionWidth": W, "video/ResolutionHeight": H, "video/Preset": PRESET, "video/QP": QP,
                           "video/Intra": GOP, "video/OWF": owf, "video/FramerateNumerator": 30})
        if not f.init():
            raise SystemExit("KvazaarFilter.init failed: " + lib.b200_last_error().decode())

        pics = f.alloc_pictures(frames) if pinned_ring else None

        def step():
            n = 0
            if pinned_ring:
                for pic in pics:
                    for au in f.feed_picture(pic, drain=stock_drain):
                        n += len(au)
            else:
                for fr in frames:
                    for au in f.feed_input(fr, drain=stock_drain):
                        n += len(au)
            return n

        for _ in range(warm):
            step()
        f.flush()
        barrier()
        t0 = time.perf_counter()
        nb = 0
        for _ in range(steps):
            nb += step()
        nb += sum(len(a) for a in f.flush())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if pics:
            f.free_pictures(pics)
        f.close()
        return GOP * steps * world / dt, nb // steps

    e2e_steps = max(1, args.steps)
    e2e_value, d2h = kvz_run(DEPTH - 1, False, e2e_steps, max(1, args.warmup // 2))
    short = max(1, min(args.steps, 4))
    e2e_stock_owf2, _ = kvz_run(2, True, short, 1)          # the reference's own loop at its largest default owf
    e2e_stock_deep, _ = kvz_run(DEPTH - 1, True, short, 1)  # the reference's own loop, deep pipeline
    e2e_pinned, _ = kvz_run(DEPTH - 1, False, e2e_steps, 1, pinned_ring=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    parity = parity_check(frames, last_gop, opts)

    # ---- roofline of the critical-path kernel ----
    hbm_peak, peak_src = peaks()
    int_peak = {name: lib.b200_int_peak(k) for k, name in enumerate(("vabsdiff4", "dp4a", "dp2a"))}
    px = W * H
    # algorithmic bytes per launch (DESIGN.md "Kernels"): every datum moved once
    alg = {
        "intra": px * 1.5 + px * 1.5 + px * 3.0,              # source read, reconstruction written, levels written
        "me": px * 1.0 + px * 1.0 + (px / 64) * 12 + 2 * px / 16,   # luma source + reference read, cu map written, quarter-res planes
        "recon": px * 1.5 * 2 + px * 1.5 + px * 3.0 + (px / 64) * 12,   # src+ref read, recon + levels written, cu map
        "modes": (px / 64) * 12 * 2,
        "deblock": 2 * (px * 1.0 * 2),                        # two passes, luma read + written
        "sao": px * 1.5 * 3,                                  # source + deblocked read, output written
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
        kernels[k] = {"launches": cnt, "avg_us": round(avg * 1e3, 2), "share_of_summed_device_time": round(ms / total_ms, 4),
                      "achieved_gbs": round(ach, 2), "frac_of_hbm": round(ach / hbm_peak, 5)}
    traffic, traffic_src = ncu_traffic()
    for k in kernels:
        kernels[k]["ncu_dram_bytes"] = traffic.get(k)
    top = "me"
    # counted integer instructions of one k_me_ctu launch: the useful arithmetic
    # only -- VABSDIFF4 of the SADs, DP4A / DP2A of the interpolation -- from the kernel's own work
    # counters; all three issue at the same measured rate (half-rate integer pipe)
    R, Rc = opts["search_range"], opts["me_coarse"]
    n_me = max(prof["me"][1], 1)
    ctus = me_stats["ctus"] / n_me
    side, cside = 2 * R + 1, 2 * Rc + 1
    instr = {
        "coarse_vabsdiff4": ctus * (cside * cside * 4 * 16 if Rc else 0),
        "fine_vabsdiff4": (ctus * 4 + me_stats["second_set_quadrants"] / n_me) * 16 * side * side * 16,
        "refine_dp4a": ctus * 64 * 16 * 4 * 60,
        "refine_dp2a": ctus * 64 * 16 * 4 * 64,
        "refine_vabsdiff4": ctus * 64 * 16 * 4 * 8,
    }
    instr_total = sum(instr.values())
    me_s = kernels["me"]["avg_us"] * 1e-6
    ipk = int_peak["vabsdiff4"]
    chain_us = sum(kernels[k]["avg_us"] for k in CHAIN if k in kernels)
    roofline = {"kernel": "k_me_ctu", "why": "largest kernel of the P-picture prediction chain (the critical path); the entropy kernels "
                                             "sum to more device time but run concurrently on other streams",
                "bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": hbm_peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": kernels[top]["frac_of_hbm"], "traffic": traffic.get(top),
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": int(alg[top]), "avg_launch_us": kernels[top]["avg_us"],
                "int_pipe": {"counted_instr_per_launch": {k: int(v) for k, v in instr.items()}, "total": int(instr_total),
                             "achieved_ginstr_s": round(instr_total / me_s / 1e9, 1), "peak_ginstr_s": round(ipk / 1e9, 1),
                             "frac": round(instr_total / me_s / ipk, 4),
                             "peak_source": "b200_int_peak, measured live (vabsdiff4 / dp4a / dp2a: %s Ginstr/s)"
                                            % "/".join(str(round(v / 1e9)) for v in int_peak.values()),
                             "me_work": {k: round(v / n_me, 1) for k, v in me_stats.items()},
                             "note": "useful SAD / interpolation instructions only; address arithmetic, shifts and shuffles excluded",
                             "ncu_pipe_utilisation": ncu_pipes("k_me_ctu")},
                "critical_path": {"kernels": list(CHAIN), "us_per_picture": round(chain_us, 1),
                                  "us_per_wavefront_step": round(chain_us / WAVEFRONT_STEPS, 2), "floor_us_per_step_at_1000fps": 16.1,
                                  "note": "sum of the average launch times of the P-picture chain; per-picture kernels cover all "
                                          "CTUs at once, so this is the dependency floor between consecutive pictures"},
                "note": "share_of_summed_device_time adds up concurrent kernels (entropy coding overlaps the chain of later "
                        "pictures): it is not a share of the critical path",
                "kernels": kernels}

    cpu = cpu_baseline_sample(frames, opts, max_seconds=25.0, threads=1)

    line = {
        "metric": "1080p HEVC encode fps", "value": round(value, 2), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(elapsed / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"1080p30 {PRESET} low-delay-P QP27, GOP 64, one stream per GPU (BASELINE configs[1])",
                   "width": W, "height": H, "qp": QP, "preset": PRESET, "gop": GOP, **opts,
                   "frames_per_step": GOP, "pictures_in_flight": DEPTH, "streams": world,
                   "l2_policy": "inputs larger than L2: 64 distinct 3.1 MB pictures (199 MB) cycled per step",
                   "bitrate_kbps_at_30fps": round(bitrate_kbps, 1)},
        "e2e": {"value": round(e2e_pinned, 2), "unit": "frames/s", "h2d_bytes_per_step": frame_bytes * GOP,
                "d2h_bytes_per_step": d2h,
                "api": "kvz_api (picture_alloc / encoder_encode / chunk_free): every picture is uploaded from the page-locked "
                       "host picture picture_alloc returned (the ring of owf + 1 pictures the reference's filter keeps, "
                       "kvazaarfilter.cpp:299) inside the timed region, every access unit comes back to host memory; owf 95, "
                       "feedInput polling once per picture (INTEGRATION.md section 1)",
                "with_filter_plane_copies": {"value": round(e2e_value, 2), "unit": "frames/s",
                                             "note": "the same plus what the reference's filter does before the call: three plane "
                                                     "memcpys from a pageable Data buffer into the ring picture "
                                                     "(kvazaarfilter.cpp:410-418) -- host work outside the C ABI, bound by the "
                                                     "box's host memory bandwidth when 8 ranks share one socket"},
                "stock_drain_loop": {"owf_2": round(e2e_stock_owf2, 2), "owf_95": round(e2e_stock_deep, 2), "unit": "frames/s",
                                     "note": "the reference's unmodified feedInput loop (kvazaarfilter.cpp:440-449), which empties "
                                             "the pipeline after every access unit; owf 2 is the largest value the reference's "
                                             "defaults ever set (defaultsettings.cpp:200-236)"}},
        "gpu_launches": int(launches),
        "parity_checked": parity["ok"], "parity": parity,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wall_s": round(wall, 3),
    }
    if args.sweep:
        line["qp_sweep"] = qp_sweep(d_frames, opts)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def parity_check(frames, timed_aus, opts):
    """The configuration that was just timed (1080p, GOP 64, 96 pictures in flight, IDR overlap) against
    a synchronous depth-1 encoder on the same pictures: the access units of the last timed step must be
    byte-identical, and FFmpeg's decoder must reproduce the encoder's reconstruction from them."""
    import hashlib

    from kvazzup_b200.encoder import GpuEncoder
    out = {"ok": False, "gop_pictures": len(timed_aus)}
    if len(timed_aus) != GOP:
        out["error"] = "timed run returned %d access units for the last step" % len(timed_aus)
        return out
    e = GpuEncoder(W, H, qp=QP, intra_period=GOP, depth=1, fps_num=30, fps_den=1, **opts)
    sync_aus = [e.encode(f) for f in frames]
    rec = e.recon()
    e.close()
    out["sha256_timed"] = hashlib.sha256(b"".join(timed_aus)).hexdigest()[:16]
    out["sha256_depth1"] = hashlib.sha256(b"".join(sync_aus)).hexdigest()[:16]
    out["depth96_equals_depth1"] = timed_aus == sync_aus
    try:
        sys.path.insert(0, str(ROOT))
        from tests import ffhevc
        if ffhevc.available():
            dec, errs = ffhevc.decode_stream(timed_aus)
            out["ffmpeg_decodes"] = errs == 0 and len(dec) == GOP
            out["ffmpeg_equals_recon"] = bool(out["ffmpeg_decodes"] and np.array_equal(dec[-1][0], rec))
        else:
            out["ffmpeg_decodes"] = None
    except Exception as ex:                      # the check is reported, never silently passed
        out["ffmpeg_error"] = repr(ex)
    out["ok"] = bool(out["depth96_equals_depth1"] and out.get("ffmpeg_equals_recon", out.get("ffmpeg_decodes") is None))
    return out


def ncu_pipes(kernel):
    """ALU pipe / issue utilisation of a kernel from the committed --set full capture (tools/summarise_ncu.py full):
    what the hardware counters say about the integer pipe, next to the counted-instruction fraction."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_full_chain.csv")
    try:
        with open(path) as f:
            hdr = next(f).strip().split(",")
            col = {h.split("[")[0]: i for i, h in enumerate(hdr)}
            rows = [ln.strip().split(",") for ln in f if ln.startswith(kernel)]
        rows = [r for r in rows if float(r[col["time"]]) > 0.05]          # the launches that did a picture's work
        if not rows:
            return None
        mean = lambda k: round(sum(float(r[col[k]]) for r in rows) / len(rows), 1)
        return {"alu_pipe_pct": mean("alu_%"), "issue_slots_pct": mean("issue_%"), "occupancy_pct": mean("occ_%"),
                "launches": len(rows), "source": "profiles/r02_ncu_full_chain.csv (ncu --set full, cold cache, serialised)"}
    except (OSError, ValueError, KeyError, StopIteration):
        return None


def ncu_traffic():
    """DRAM bytes per launch of each kernel from the newest committed ncu launch summary, and its name."""
    names = {"k_intra_frame<0>": "intra", "k_intra_frame": "intra", "k_intra_modes": "intra_modes", "k_me_ctu": "me",
             "k_me_ctu<0>": "me", "k_me_ctu<1>": "me", "k_inter_recon<0>": "recon",
             "k_inter_modes": "modes", "k_deblock": "deblock", "k_sao_ctu<1>": "sao", "k_binarise": "binarise",
             "k_ctx_rows": "ctx", "k_arith_rows": "arith", "k_entropy_rows": "arith", "k_pack_rows": "pack"}
    out = {}
    for name in ("r02_ncu_launch_summary.csv", "r01_ncu_launch_summary.csv"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            with open(path) as f:
                next(f)
                for ln in f:
                    c = ln.strip().split(",")
                    if c[0] in names:
                        out[names[c[0]]] = int((float(c[5]) + float(c[6])) * 1e6)
        except (OSError, ValueError, IndexError):
            continue
        if out:
            note = "dram__bytes_read+write per launch under ncu"
            if name.startswith("r01"):
                note += "; CAPTURED IN ROUND 1, before this round's kernels changed -- indicative only"
            return out, "profiles/%s (%s)" % (name, note)
    return out, None


def qp_sweep(d_frames, opts):
    """fps / bitrate / PSNR-Y at the four QPs of BASELINE configs[1] (outside the timed region)."""
    import torch

    from kvazzup_b200 import synth
    from kvazzup_b200.encoder import GpuEncoder
    out = {}
    for qp in (22, 27, 32, 37):
        e = GpuEncoder(W, H, qp=qp, intra_period=GOP, depth=DEPTH, **opts)
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

def oracle_options(opts):
    """The oracle port runs the same algorithm as the engine options of the preset."""
    return {"search_range": opts["search_range"], "me_coarse": opts["me_coarse"], "sao": opts["sao"],
            "intra_in_p": opts["intra_in_p"], "intra_satd": opts.get("intra_satd", 0),
            "subme_satd": opts.get("subme_satd", 0), "fps_num": 30, "fps_den": 1}


def cpu_baseline_sample(frames, opts, max_seconds, threads):
    """Encode a bounded sample (1 IDR + a few P pictures) of the same stream with the CPU oracle."""
    import oracle
    from oracle.encoder import OracleEncoder
    lib = oracle.load()
    lib.orc_set_threads(threads)
    enc = OracleEncoder(W, H, qp=QP, intra_period=GOP, **oracle_options(opts))
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
            "sample": f"first {n} pictures (1 IDR + {n - 1} P) of the same 1080p stream, in-house oracle encoder, gcc -O3 -mavx2 "
                      f"(not Kvazaar: absent from the reference tree and this image)"}


def find_kvazaar():
    """A Kvazaar binary, should the box have one (baseline/_ref, PATH, pkg-config prefix)."""
    import shutil
    cands = [ROOT / "baseline" / "_ref" / "bin" / "kvazaar", ROOT / "baseline" / "_ref" / "kvazaar"]
    w = shutil.which("kvazaar")
    if w:
        cands.append(Path(w))
    try:
        r = subprocess.run(["pkg-config", "--variable=prefix", "kvazaar"], capture_output=True, text=True, timeout=10)
        if r.returncode == 0 and r.stdout.strip():
            cands.append(Path(r.stdout.strip()) / "bin" / "kvazaar")
    except Exception:
        pass
    for c in cands:
        if c.is_file() and os.access(c, os.X_OK):
            return c
    return None


def run_kvazaar(binary, frames, threads):
    """Real Kvazaar (AVX2 strategies are selected by its own cpuid dispatch) on the same pictures."""
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        yuv = Path(td) / "in.yuv"
        with open(yuv, "wb") as f:
            for fr in frames:
                f.write(fr.tobytes())
        cmd = [str(binary), "-i", str(yuv), "--input-res", f"{W}x{H}", "-n", str(len(frames)), "--preset", PRESET, "-q", str(QP),
               "--period", str(GOP), "--gop", "lp-g4d3t1", "--threads", str(threads), "--owf", "auto", "--no-psnr", "--no-info",
               "-o", str(Path(td) / "out.hevc")]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError("kvazaar failed: " + r.stderr[-300:])
        return len(frames) / dt, (Path(td) / "out.hevc").stat().st_size


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import oracle
    lib = oracle.load()
    threads = max(1, min(lib.orc_max_threads(), os.cpu_count() or 1))
    opts = engine_options()
    per_step = 3                                    # pictures per step: bounded sample of the GOP workload
    # consecutive pictures of the same stream; the sample wraps at the IDR period, where the encoder
    # codes an IDR anyway, so the wrap is not an artificial scene cut
    n_frames = min(GOP, per_step * (args.steps + args.warmup))
    frames = make_source(n_frames, 0)
    kvz = find_kvazaar()
    if kvz is not None:
        try:
            fps, nbytes = run_kvazaar(kvz, frames[:min(len(frames), per_step * args.steps)], threads)
            line = {
                "impl": "reference", "metric": "1080p HEVC encode fps", "value": round(fps, 4), "unit": "frames/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(per_step / fps * 1e3, 2),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "1080p30 veryfast low-delay-P QP27, GOP 64, one stream (BASELINE configs[1])",
                           "width": W, "height": H, "qp": QP, "preset": PRESET, "gop": GOP, "frames_per_step": per_step},
                "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": threads, "kind": "reference",
                                 "sample": f"{kvz} (Kvazaar CLI, its own cpuid-selected strategies), --threads {threads}, "
                                           f"{len(frames)} pictures of the same 1080p stream, {nbytes} bytes out"},
                "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
            }
            print(json.dumps(line))
            return
        except Exception as ex:
            print(f"bench.py: Kvazaar at {kvz} could not be used ({ex}); falling back to the oracle port", file=sys.stderr)
    from oracle.encoder import OracleEncoder
    lib.orc_set_threads(threads)
    enc = OracleEncoder(W, H, qp=QP, intra_period=GOP, **oracle_options(opts))
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
                   "width": W, "height": H, "qp": QP, "preset": PRESET, "gop": GOP, **opts, "frames_per_step": per_step,
                   "same_algorithm_as_gpu_arm": True},
        "cpu_baseline": {"value": round(value, 4), "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} consecutive pictures per step of the same 1080p stream; in-house oracle encoder "
                                   f"(gcc -O3 -mavx2, OpenMP over CTUs, {threads} threads).  No Kvazaar binary on this box "
                                   f"(probed baseline/_ref, PATH, pkg-config): this is the GPU encoder's own algorithm on the "
                                   f"CPU, not Kvazaar's AVX2 strategies"},
        "e2e": {"value": round(value, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_conference(args):
    """BASELINE config 4: concurrent 720p30 call streams (encode + decode each), sharded by participant
    over the GPUs of the box (stream s -> GPU s mod N, no data shared, no collective), every stream paced
    at 30 fps with nothing in flight -- the C++ harness over the C ABI (tools/conference_bench.cpp), one
    host thread per stream as in the reference's Filter model.  Rank 0 drives all GPUs."""
    if env_int("RANK", 0) != 0:
        return
    import numpy as np
    from kvazzup_b200 import synth
    exe = "/tmp/conference_bench"
    lib_dir = ROOT / "kvazzup_b200"
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-I" + str(ROOT / "include"), str(ROOT / "tools" / "conference_bench.cpp"), "-o", exe,
                        "-L" + str(lib_dir), "-lb200media", "-Wl,-rpath," + str(lib_dir), "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit("cannot build tools/conference_bench.cpp: " + r.stderr[-400:])
    cw, ch, nfile = 1280, 720, 30
    yuv = "/tmp/conf_720p.yuv"
    np.concatenate([synth.camera_i420(cw, ch, t) for t in range(nfile)]).tofile(yuv)
    frames = max(30, min(300, 30 * args.steps))
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    results, best = [], None
    for per_gpu in (20, 30, 40):
        streams = per_gpu * args.gpus
        p = subprocess.run([exe, yuv, str(cw), str(ch), str(nfile), str(streams), str(frames), "0", "1", "30", str(args.gpus)],
                           capture_output=True, text=True, env=env, timeout=900)
        try:
            d = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception:
            d = {"streams": streams, "error": (p.stderr or p.stdout)[-300:]}
        results.append(d)
        ok = d.get("all_pictures_decoded") and d.get("achieved_fps_per_stream", 0) >= 29.5
        if ok:
            best = d
        else:
            break
    value = best["streams"] if best else 0
    line = {
        "metric": "concurrent 720p30 encode+decode streams", "value": value, "unit": "streams", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000.0 / 30, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "multi-party conference: 720p30 veryfast QP32 encode + decode per stream, paced at 30 fps, nothing in "
                               "flight, streams sharded by participant over the GPUs (BASELINE configs[3])",
                   "frames_per_stream": frames, "streams_tried_per_gpu": [20, 30, 40],
                   "sustained_means": "every picture decoded and >= 29.5 fps per stream"},
        "latency_ms": best["latency_ms"] if best else None,
        "streams_hash": best.get("streams_hash") if best else None,
        "runs": results,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweep", action="store_true", help="also report the QP 22/27/32/37 sweep")
    ap.add_argument("--workload", default="encode", choices=["encode", "conference"],
                    help="encode: the headline 1080p metric (default); conference: concurrent 720p30 call streams over --gpus GPUs")
    args = ap.parse_args()
    if args.workload == "conference":
        run_conference(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
