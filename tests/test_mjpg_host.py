"""Host side of the MJPG conversion (header parser + Huffman decoder of kvazzup_b200/csrc/mjpg.cu) through
b200_mjpg_probe, which needs no GPU: committed frames are accepted with the right geometry, and corrupted or
truncated frames never crash the parser (it runs on whatever a camera or a capture pipeline hands over)."""
import ctypes as C
import json
from pathlib import Path

import numpy as np

import kvazzup_b200

GOLDEN = Path(__file__).parent / "golden"


def probe(lib, data: bytes):
    a = np.frombuffer(data + b"\0", np.uint8)
    w, h, s = C.c_int(), C.c_int(), C.c_int()
    rc = lib.b200_mjpg_probe(a.ctypes.data, len(data), C.byref(w), C.byref(h), C.byref(s))
    return rc, w.value, h.value, s.value


def test_probe_accepts_the_committed_frames_and_survives_corruption():
    lib = kvazzup_b200.lib()
    lib.b200_mjpg_probe.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    gold = json.loads((GOLDEN / "mjpg_golden.json").read_text())
    want = {"mjpg_422_camera.jpg": 422, "mjpg_420_noise.jpg": 420, "mjpg_444_rst.jpg": 444, "mjpg_422_no_dht.jpg": 422}
    rng = np.random.default_rng(11)
    for name, e in gold["frames"].items():
        jpeg = (GOLDEN / name).read_bytes()
        assert probe(lib, jpeg) == (0, e["w"], e["h"], want[name]), name
        for trial in range(150):
            bad = bytearray(jpeg)
            kind = trial % 3
            if kind == 0:                                   # flipped bytes anywhere (headers included)
                for _ in range(int(rng.integers(1, 8))):
                    bad[int(rng.integers(2, len(bad)))] ^= int(rng.integers(1, 256))
            elif kind == 1:                                 # truncated
                bad = bad[:int(rng.integers(2, len(bad)))]
            else:                                           # a run of 0xff (markers in the middle of the scan)
                k = int(rng.integers(2, len(bad) - 8))
                bad[k:k + 4] = b"\xff\xd9\xff\xc4"
            rc, w, h, s = probe(lib, bytes(bad))
            assert rc in (0, -1, -2), (name, trial, rc)      # accepted or refused; what matters is getting here
    assert probe(lib, b"")[0] < 0 and probe(lib, b"\xff\xd8")[0] < 0
