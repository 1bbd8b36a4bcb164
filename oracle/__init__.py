"""TEST INFRASTRUCTURE ONLY: Python handle on the CPU oracle.

`oracle/` is the checker, never the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package.
kvazzup_b200/ must never import it.
"""
from .binding import build, load, load_ref, ref_available  # noqa: F401
