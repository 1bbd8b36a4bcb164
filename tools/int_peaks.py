"""Measured integer-pipe peaks of the box's GPU -> profiles/r02_int_peaks.json.

Run on the GPU box:  python tools/int_peaks.py [out.json]
Register-only microbenchmarks inside libb200media.so (csrc/int_peaks.cu); thread-level
instructions per second over the whole GPU, best of 5 launches, CUDA events.
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import kvazzup_b200  # noqa: E402

KINDS = ["vabsdiff4_add", "dp4a", "dp2a", "imad", "shf", "iadd"]


def measure() -> dict:
    l = kvazzup_b200.lib()
    out = {}
    for k, name in enumerate(KINDS):
        v = l.b200_int_peak(k)
        if v < 0:
            raise RuntimeError(l.b200_last_error().decode())
        out[name] = {"ginstr_per_s": round(v / 1e9, 1)}
    out["vabsdiff4_add"]["gbyte_ops_per_s"] = round(4 * out["vabsdiff4_add"]["ginstr_per_s"], 1)
    out["dp4a"]["gmac_per_s"] = round(4 * out["dp4a"]["ginstr_per_s"], 1)
    out["dp2a"]["gmac_per_s"] = round(2 * out["dp2a"]["ginstr_per_s"], 1)
    return out


if __name__ == "__main__":
    res = {"how": "csrc/int_peaks.cu: 8 independent chains x 4096 iterations per thread, 8 CTAs x 256 threads per SM, "
                  "best of 5, CUDA events; thread-level instructions per second", "peaks": measure()}
    text = json.dumps(res, indent=1)
    print(text)
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(text + "\n")
