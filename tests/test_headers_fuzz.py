"""The host-side parsers of network input (VPS / SPS / PPS / slice segment headers, hevc_headers.cpp) under
AddressSanitizer + UndefinedBehaviorSanitizer on mutated oracle streams: any finding aborts the harness."""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_header_parsers_survive_mutated_streams_under_asan_ubsan(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = tmp_path / "harness"
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        "-I", str(ROOT / "kvazzup_b200/csrc"), str(ROOT / "tools/fuzz/headers_harness.cpp"),
                        str(ROOT / "kvazzup_b200/csrc/hevc_headers.cpp"), "-o", str(exe)], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("sanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr
    cases = tmp_path / "cases.bin"
    subprocess.run([sys.executable, str(ROOT / "tools/fuzz/gen_header_cases.py"), str(cases), "6000", "3"], check=True)
    r = subprocess.run([str(exe), str(cases)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "cases 6000" in r.stdout
