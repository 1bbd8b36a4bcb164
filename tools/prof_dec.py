"""Decode a few 1080p pictures (the GPU encoder's veryfast stream), convert an MJPG frame and a 24-bit RGB
picture (for ncu captures of the decoder and conversion kernels)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from kvazzup_b200 import convert, synth
from kvazzup_b200.capi import FOURCC
from kvazzup_b200.encoder import GpuEncoder, preset_options
from kvazzup_b200.openhevc import OpenHEVCFilter, split_nals
w, h = 1920, 1080
enc = GpuEncoder(w, h, qp=27, intra_period=64, **preset_options("veryfast"))
aus = [enc.encode(synth.camera_i420(w, h, t)) for t in range(3)]
enc.close()
dec = OpenHEVCFilter(); dec.init()
for a in aus:
    for nal in split_nals(a):
        dec.process(nal)
dec.close()
convert.convert_to_i420(synth.noise(3, w * h * 3), w, h, FOURCC["24BG"])
convert.convert_to_i420(synth.noise(4, w * h * 4), w, h, FOURCC["ARGB"])
try:
    from tests import mjpg_util
    convert.convert_to_i420(np.frombuffer(mjpg_util.make_jpeg(w, h, 85, "422"), np.uint8), w, h, FOURCC["MJPG"])
except Exception as e:
    print("mjpg skipped:", e)
