"""ctypes loader for oracle/_build/liboracle.so and oracle/_ref/*.so (test infrastructure)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"
REF = HERE / "_ref" / "libref_yuvconversions.so"
_lib = None
_ref = None


def build(force: bool = False) -> None:
    """Run the committed recipe (oracle/Makefile): oracle restatement + oracle/_ref when
    /root/reference is present."""
    srcs = list(HERE.glob("*.c")) + list(HERE.glob("*.h")) + [HERE / "Makefile"]
    stale = (not LIB.exists()) or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(HERE), "_build/liboracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if force or not REF.exists():
        subprocess.run(["make", "-C", str(HERE), "ref"], check=False,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB), mode=C.RTLD_LOCAL)
        from . import sigs
        sigs.bind(_lib)
        try:
            _lib.orc_set_threads(1)        # deterministic default; bench.py raises it explicitly
        except AttributeError:
            pass
    return _lib


def ref_available() -> bool:
    if not REF.exists():
        build()
    return REF.exists()


def load_ref():
    """The reference's own yuvconversions.cpp object code (built unmodified)."""
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref not built and /root/reference absent")
        _ref = C.CDLL(str(REF), mode=C.RTLD_LOCAL)
        v, u16, u8 = C.c_void_p, C.c_uint16, C.c_uint8
        _ref.ref_yuv420_to_rgb_i_avx2_mt.argtypes = [v, v, u16, u16, u8]
        _ref.ref_yuv420_to_rgb_i_avx2.argtypes = [v, v, u16, u16]
        _ref.ref_yuv420_to_rgb_i_sse41.argtypes = [v, v, u16, u16]
        _ref.ref_yuv420_to_rgb_i_c.argtypes = [v, v, u16, u16]
        _ref.ref_half_rgb.argtypes = [v, v, u16, u16]
        _ref.ref_flip_rgb.argtypes = [v, v, u16, u16, C.c_int, C.c_int]
    return _ref
