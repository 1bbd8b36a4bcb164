"""CPU-only: libb200media.so builds, loads and exports every symbol include/*.h declares."""
import ctypes as C
import re
from pathlib import Path

import kvazzup_b200
from kvazzup_b200 import capi as libmod

ROOT = Path(__file__).resolve().parent.parent

DECL = re.compile(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\(", re.M)


def declared_symbols():
    names = set()
    for hdr in sorted((ROOT / "include").glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", hdr.read_text(), flags=re.S)
        text = re.sub(r"//.*", "", text)
        text = re.sub(r"^\s*#.*(?:\\\n.*)*", "", text, flags=re.M)
        text = re.sub(r"typedef\s+struct\s+\w*\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
        text = re.sub(r"struct\s+\w+\s*\{.*?\}\s*;", "", text, flags=re.S)
        for m in re.finditer(r"\b(\w+)\s*\([^;{}]*\)\s*;", text):
            name = m.group(1)
            if name not in ("defined", "sizeof"):
                names.add(name)
    return names


def test_library_builds_and_exports_every_declared_symbol():
    l = kvazzup_b200.load()
    syms = declared_symbols()
    assert "b200_yuv420_to_rgb32" in syms and "b200_ConvertToI420" in syms
    missing = [s for s in sorted(syms) if not hasattr(l, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_runtime_queries_do_not_need_a_gpu():
    l = kvazzup_b200.lib()
    assert l.b200_device_count() >= 0
    assert b"sm_100a" in l.b200_version()
    assert l.b200_frame_bytes(libmod.FOURCC["YUYV"], 640, 480) == 640 * 480 * 2
    assert l.b200_frame_bytes(2, 640, 480) == 0


def test_product_never_imports_the_oracle():
    for py in (ROOT / "kvazzup_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
    for cu in (ROOT / "kvazzup_b200" / "csrc").glob("*"):
        if cu.is_file():
            code = re.sub(r"/\*.*?\*/", "", cu.read_text(), flags=re.S)
            code = re.sub(r"//.*", "", code)
            assert "oracle" not in code, cu          # no include, link or path into oracle/


def test_compute_fails_loudly_without_gpu():
    l = kvazzup_b200.lib()
    if l.b200_device_count() > 0:
        return
    import numpy as np
    a = np.zeros(2 * 2 * 3 // 2, np.uint8)
    o = np.zeros(16, np.uint8)
    rc = l.b200_yuv420_to_rgb32(C.c_void_p(a.ctypes.data), C.c_void_p(o.ctypes.data), 2, 2)
    assert rc < 0 and b"no CPU fallback" in l.b200_last_error()


def test_headers_are_plain_c_and_self_contained(tmp_path):
    """The drop-in boundary is a C ABI: every header must compile as C99 on its own (no C++ / CUDA /
    torch types in the signatures), and a C program must link against the library."""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else shutil.which("gcc")
    assert gcc
    for hdr in sorted((ROOT / "include").glob("*.h")):
        src = tmp_path / f"only_{hdr.stem}.c"
        src.write_text(f'#include "{hdr.name}"\nint main(void) {{ return 0; }}\n')
        r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", str(ROOT / "include"), str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, f"{hdr.name}: {r.stderr}"
    prog = tmp_path / "link.c"
    prog.write_text('''
#include <stdio.h>
#include "b200media.h"
#include "b200_kvazaar.h"
#include "b200_openhevc.h"
#include "b200_rtp.h"
#include "b200_hevc.h"
int main(void) {
  const kvz_api *api = kvz_api_get(8);
  b200_nal_span span[4];
  const unsigned char au[] = {0, 0, 0, 1, 0x40, 1, 0xaa, 0, 0, 1, 0x42, 1, 0xbb};
  int n = b200_annexb_split(au, sizeof au, span, 4);
  printf("%s %d %d %d\\n", b200_version(), api != NULL, kvz_api_get(10) == NULL, n);
  return (api && n == 2) ? 0 : 1;
}
''')
    exe = tmp_path / "link_test"
    libdir = ROOT / "kvazzup_b200"
    kvazzup_b200.load()
    r = subprocess.run([gcc, "-std=c99", "-I", str(ROOT / "include"), str(prog), "-o", str(exe), "-L", str(libdir), "-lb200media",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "b200media" in r.stdout, r.stdout + r.stderr


def test_kvz_config_parse_follows_the_reference_contract():
    """config_parse returns 1 on success and the reference only warns otherwise
    (kvazaarfilter.cpp:363-369): known Kvazaar options are applied or knowingly ignored, unknown names
    and unsupported values return 0.  No GPU needed: the encoder is not opened."""
    from kvazzup_b200.kvazaar import kvz_api_get
    api = kvz_api_get(8)
    assert api is not None
    cfg = api.config_alloc()
    assert api.config_init(cfg) == 1
    ok = lambda n, v: api.config_parse(cfg, n.encode(), v.encode() if v is not None else None)
    # a preset selects the search (window around each centre, coarse-level range) and SAO
    for preset, rng, coarse, sao in (("ultrafast", 4, 16, 0), ("veryfast", 6, 16, 3), ("medium", 12, 32, 3), ("placebo", 16, 32, 3)):
        assert ok("preset", preset) == 1
        assert (cfg.contents.me_range, cfg.contents.me_coarse, cfg.contents.sao_type) == (rng, coarse, sao)
    assert ok("sao", "band") == 1 and cfg.contents.sao_type == 2 and ok("sao", "full") == 1 and ok("sao", "purple") == 0
    assert ok("b200-me-coarse", "8") == 1 and cfg.contents.me_coarse == 8 and ok("b200-me-coarse", "7") == 0
    assert ok("preset", "warp-speed") == 0
    assert ok("qp", "27") == 1 and cfg.contents.qp == 27
    assert ok("qp", "52") == 0 and ok("qp", "abc") == 0
    assert ok("period", "64") == 1 and ok("vps-period", "1") == 1 and ok("owf", "3") == 1 and cfg.contents.owf == 3
    assert ok("gop", "lp-g4d3t1") == 1 and ok("intra-bits", "") == 1 and ok("rd", "0") == 1 and ok("sao", "off") == 1
    assert ok("tiles", "3x1") == 1 and cfg.contents.tiles_width_count == 3
    assert ok("tiles", "2x2") == 1 and (cfg.contents.tiles_width_count, cfg.contents.tiles_height_count) == (2, 2)   # the reference's default
    assert ok("tiles", "9x9") == 0 and cfg.contents.tiles_width_count == 2      # more than 64 tiles: refused, state unchanged
    assert ok("tiles", "banana") == 0 and ok("slices", "tiles") == 0
    assert ok("b200-roi", "1") == 1 and cfg.contents.roi_enable == 1
    assert ok("set-qp-in-cu", "1") == 1 and cfg.contents.set_qp_in_cu == 1
    assert ok("bitrate", "1500000") == 1 and cfg.contents.target_bitrate == 1500000
    assert ok("rc-algorithm", "lambda") == 1 and ok("rc-algorithm", "no-rc") == 1 and ok("rc-algorithm", "oba") == 0   # oba is not built
    assert ok("no-such-option", "1") == 0
    api.config_destroy(cfg)
    assert kvz_api_get(10) is None                      # only 8-bit, like the reference's kvz_api_get(8) (:145)
