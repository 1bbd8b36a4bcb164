"""A few pictures through the encoder with vaq on and with a padded source size: what
`ncu --kernel-name regex:k_vaq|k_pad` lists (profiles/r02_ncu_new_kernels.csv)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from kvazzup_b200 import synth  # noqa: E402
from kvazzup_b200.encoder import GpuEncoder, preset_options  # noqa: E402

vf = preset_options("veryfast")
e = GpuEncoder(1920, 1080, qp=27, qp_delta=1, vaq=10, **vf)
for t in range(4):
    e.encode(synth.camera_i420(1920, 1080, t))
e.close()
e = GpuEncoder(1368, 768, qp=27, src_width=1366, src_height=768, **vf)
for t in range(4):
    e.encode(synth.camera_i420(1366, 768, t))
e.close()
