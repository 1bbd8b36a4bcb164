import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder
from kvazzup_b200.openhevc import OpenHEVCFilter, split_nals
w, h = 1920, 1080
enc = GpuEncoder(w, h, qp=27, intra_period=64, search_range=12)
aus = [enc.encode(synth.camera_i420(w, h, t)) for t in range(3)]
enc.close()
dec = OpenHEVCFilter(); dec.init()
for a in aus:
    for nal in split_nals(a):
        dec.process(nal)
dec.close()
