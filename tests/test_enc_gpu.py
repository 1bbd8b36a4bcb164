"""GPU parity: the CUDA HEVC encoder vs the CPU oracle, stage by stage and bit for bit.

Integer work, so the bar is equality: cu map (partition, motion vectors, modes, cbf, merge / skip /
mvp decisions), quantised levels, reconstruction before and after deblocking, and the access unit
bytes.  The oracle itself is pinned by an independent decoder (tests/test_oracle_hevc.py); the GPU
streams are additionally decoded by that decoder here.
"""
import numpy as np
import pytest

from kvazzup_b200 import synth
from kvazzup_b200.encoder import GpuEncoder
from oracle.encoder import OracleEncoder
from tests import ffhevc
from tests.test_oracle_hevc import frames_of

pytestmark = pytest.mark.gpu

CU_FIELDS = ("log2_size", "pred_mode", "intra_mode", "mvx", "mvy", "cbf", "merge_idx", "skip", "mvp_idx")


def compare_frame(tag, g, o, w, h):
    gc, oc = g.cu_map(), o.cu_map()
    for f in CU_FIELDS:
        bad = np.flatnonzero(gc[f] != oc[f])
        assert bad.size == 0, f"{tag}: cu map field {f} differs at units {bad[:8]} gpu={gc[f][bad[:8]]} oracle={oc[f][bad[:8]]}"
    gl, ol = g.levels(), o.levels()
    bad = np.flatnonzero(gl != ol)
    assert bad.size == 0, f"{tag}: levels differ at {bad[:8]} (of {bad.size})"
    gp, op = g.recon_predeblock(), o.recon_predeblock()
    bad = np.flatnonzero(gp != op)
    assert bad.size == 0, f"{tag}: reconstruction before deblocking differs at {bad[:8]} (of {bad.size})"
    gr, orr = g.recon(), o.recon()
    bad = np.flatnonzero(gr != orr)
    assert bad.size == 0, f"{tag}: reconstruction after deblocking differs at {bad[:8]} (of {bad.size})"


CASES = [
    ("camera", 64, 64, 2, 32, {}),
    ("camera", 192, 136, 4, 32, {}),
    ("camera", 72, 200, 3, 27, {"deblock": 0}),
    ("noise", 128, 72, 3, 0, {}),
    ("noise", 128, 72, 3, 10, {}),
    ("noise", 128, 72, 3, 45, {}),
    ("camera", 128, 72, 3, 51, {}),
    ("camera", 416, 240, 5, 27, {}),
    ("screen", 416, 240, 5, 32, {}),
    ("camera", 200, 200, 7, 30, {"intra_period": 3}),
    ("camera", 256, 128, 4, 22, {"search_range": 16}),
    ("camera", 64, 8, 3, 37, {}),
    ("camera", 640, 480, 3, 32, {}),
]


@pytest.mark.parametrize("kind,w,h,n,qp,kw", CASES)
def test_gpu_encoder_matches_oracle_stage_by_stage(kind, w, h, n, qp, kw):
    frames = frames_of(kind, w, h, n)
    args = {"intra_period": 0} | kw
    g = GpuEncoder(w, h, qp=qp, debug=1, **args)
    o = OracleEncoder(w, h, qp=qp, **args)
    aus = []
    for i, f in enumerate(frames):
        ga, oa = g.encode(f), o.encode(f)
        compare_frame(f"{kind} {w}x{h} qp{qp} frame {i}", g, o, w, h)
        assert ga == oa, f"frame {i}: access unit differs (gpu {len(ga)} B, oracle {len(oa)} B)"
        assert g.bins() == o.bins()
        aus.append(ga)
    if ffhevc.available():
        dec, errs = ffhevc.decode_stream(aus)
        assert errs == 0 and len(dec) == n
        assert np.array_equal(dec[-1][0], g.recon())
    g.close()
    o.close()


def test_device_resident_input_gives_the_same_stream():
    import torch
    w, h = 192, 136
    frames = frames_of("camera", w, h, 3)
    a = GpuEncoder(w, h, qp=30, intra_period=0)
    b = GpuEncoder(w, h, qp=30, intra_period=0)
    for f in frames:
        d = torch.from_numpy(f).cuda()
        torch.cuda.synchronize()
        assert a.encode(f) == b.encode_dev(d)


def test_full_hd_two_frames_match_oracle():
    """BASELINE config 2 size (1080p, partial bottom CTU row): I + P picture, bit-identical stream."""
    w, h = 1920, 1080
    frames = frames_of("camera", w, h, 2)
    g = GpuEncoder(w, h, qp=27, intra_period=0, debug=1)
    o = OracleEncoder(w, h, qp=27, intra_period=0)
    for i, f in enumerate(frames):
        ga, oa = g.encode(f), o.encode(f)
        compare_frame(f"1080p frame {i}", g, o, w, h)
        assert ga == oa


def test_encoder_rejects_bad_configuration():
    from kvazzup_b200.capi import B200Error
    for bad in ((100, 64, 30), (64, 64, 52), (0, 0, 30)):
        with pytest.raises(B200Error):
            GpuEncoder(bad[0], bad[1], qp=bad[2])
