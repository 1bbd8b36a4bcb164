"""ctypes signatures for liboracle.so (test infrastructure)."""
import ctypes as C

v = C.c_void_p
i = C.c_int

SIGS = {
    "oracle_i420_to_rgb32": (None, [v, v, i, i]),
    "oracle_half_rgb": (None, [v, v, i, i]),
    "oracle_flip_rgb": (None, [v, v, i, i, i, i]),
    "oracle_convert_to_i420": (i, [v, C.c_size_t, v, i, v, i, v, i, i, i, C.c_uint32]),
}


def bind(lib) -> None:
    for name, (res, args) in SIGS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue
        fn.restype = res
        fn.argtypes = args
