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


def test_host_pass_survives_mutated_frames_under_asan_ubsan(tmp_path):
    """Marker parsing, Huffman table construction and the entropy-coded scan (the host half of the MJPG path) under
    AddressSanitizer + UndefinedBehaviorSanitizer on mutated golden frames (tools/fuzz/mjpg_harness.cpp)."""
    import os
    import shutil
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    exe = tmp_path / "mjpg_harness"
    b = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-g", "-std=c++17", "-Xcompiler=-fsanitize=address",
                        "-Xcompiler=-fsanitize=undefined", "-Xcompiler=-fno-sanitize-recover=undefined", "-I", str(root / "include"),
                        str(root / "tools/fuzz/mjpg_harness.cpp"), str(root / "kvazzup_b200/csrc/mjpg.cu"), str(root / "kvazzup_b200/csrc/runtime.cu"),
                        "-o", str(exe)], capture_output=True, text=True)
    if b.returncode != 0 and "sanitize" in b.stderr + b.stdout:
        pytest.skip("sanitizer runtime not available")
    assert b.returncode == 0, (b.stdout + b.stderr)[-2000:]
    cases = tmp_path / "cases.bin"
    subprocess.run([sys.executable, str(root / "tools/fuzz/gen_mjpg_cases.py"), str(cases), "4000", "9"], check=True)
    out = subprocess.run([str(exe), str(cases)], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, ASAN_OPTIONS="protect_shadow_gap=0"))
    assert out.returncode == 0, (out.stdout + out.stderr)[-2000:]
    assert out.stdout.startswith("cases 4000")
