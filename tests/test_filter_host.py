"""The Qt-free Filter / Data runtime (kvazzup_b200/host/filter.h) and
BASELINE config 1 as a filter graph over the C ABI (tools/loopback_pipeline.cpp).

CPU: queue caps, drop policy and fan-out deep copy against the rules of the reference's filter.cpp.
GPU: 640x480 YUYV -> I420 -> HEVC ultrafast QP32 -> NALs -> decode -> RGB32 through the threaded graph
gives, picture for picture, what the oracle chain gives (CPU encoder's reconstruction converted by the
reference-pinned conversion oracle), and the self-view branch (the deep-copied I420 frame) the
conversion of the camera picture."""
import json
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def build(src, exe, link_lib):
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I" + str(ROOT / "include"), str(ROOT / src), "-o", exe, "-lpthread"]
    if link_lib:
        import kvazzup_b200
        kvazzup_b200.load()
        lib_dir = str(ROOT / "kvazzup_b200")
        cmd += ["-L" + lib_dir, "-lb200media", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_filter_runtime_semantics(tmp_path):
    exe = str(tmp_path / "filter_semantics")
    build("tests/cpp/filter_semantics.cpp", exe, link_lib=False)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_filter_runtime_semantics_under_sanitizers(tmp_path, sanitizer):
    """The same program under ThreadSanitizer (the runtime is one thread per filter, bounded queues with a drop
    policy, fan-out copies) and under AddressSanitizer + UndefinedBehaviorSanitizer: no report, same verdict."""
    exe = str(tmp_path / "filter_semantics_san")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=" + sanitizer, "-fno-sanitize-recover=all", "-I" + str(ROOT / "include"),
           str(ROOT / "tests/cpp/filter_semantics.cpp"), "-o", exe, "-lpthread"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip("sanitizer runtime not available")
    assert b.returncode == 0, b.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK") and "Sanitizer" not in r.stderr, r.stdout + r.stderr[-2000:]


def test_loopback_program_builds_against_the_library(tmp_path):
    build("tools/loopback_pipeline.cpp", str(tmp_path / "loopback_pipeline"), link_lib=True)


@pytest.mark.gpu
def test_config1_loopback_graph_equals_the_oracle_chain(tmp_path, oracle_lib):
    from kvazzup_b200 import synth
    from kvazzup_b200.capi import FOURCC
    from kvazzup_b200.encoder import preset_options
    from oracle.encoder import OracleEncoder
    from tests.helpers import oracle_convert_to_i420, oracle_i420_to_rgb32
    w, h, n = 640, 480, 6
    cams = [synth.i420_to_yuyv(synth.camera_i420(w, h, t), w, h) for t in range(n)]
    yuyv = tmp_path / "cam.yuyv"
    np.concatenate(cams).tofile(yuyv)
    exe = str(tmp_path / "loopback_pipeline")
    build("tools/loopback_pipeline.cpp", exe, link_lib=True)
    env = dict(os.environ, B200_LOOPBACK_DUMP=str(tmp_path / "out"))
    r = subprocess.run([exe, str(yuyv), str(w), str(h), str(n), str(n), "ultrafast", "32"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["displayed"] == n and line["selfview"] == n and line["dropped"] == 0
    shown = np.fromfile(str(tmp_path / "out.display.rgb"), np.uint8).reshape(n, w * h * 4)
    selfv = np.fromfile(str(tmp_path / "out.selfview.rgb"), np.uint8).reshape(n, w * h * 4)
    enc = OracleEncoder(w, h, qp=32, intra_period=64, fps_num=30, fps_den=1, **preset_options("ultrafast"))
    for t in range(n):
        rc, i420 = oracle_convert_to_i420(oracle_lib, cams[t], w, h, FOURCC["YUYV"])
        assert rc == 0
        assert np.array_equal(selfv[t], oracle_i420_to_rgb32(oracle_lib, i420, w, h)), f"self view {t}"
        enc.encode(i420)
        assert np.array_equal(shown[t], oracle_i420_to_rgb32(oracle_lib, enc.recon(), w, h)), f"display {t}"
    enc.close()
