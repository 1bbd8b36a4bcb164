"""Mutated JPEG frames for tools/fuzz/mjpg_harness.cpp: records {u32 length, bytes}.  Starts from the committed
golden frames (tests/golden/mjpg_*.jpg); bit flips, byte replacements, deletions, insertions, truncations, with
half of the mutations aimed at the marker segments at the front (SOF / DHT / DQT / SOS / DRI).

  python tools/fuzz/gen_mjpg_cases.py cases.bin 20000 [seed]
"""
import random
import struct
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
random.seed(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
frames = [p.read_bytes() for p in sorted((ROOT / "tests" / "golden").glob("mjpg_*.jpg"))]
with open(sys.argv[1], "wb") as f:
    for it in range(int(sys.argv[2])):
        s = bytearray(random.choice(frames))
        for _ in range(random.randint(1, 6)):
            mode = random.random()
            k = random.randrange(min(len(s), 700)) if random.random() < 0.5 else random.randrange(len(s))
            if mode < 0.55: s[k] ^= 1 << random.randrange(8)
            elif mode < 0.75: s[k] = random.choice((0, 0xff, random.randrange(256)))
            elif mode < 0.85: del s[k:k + random.randint(1, 16)]
            elif mode < 0.95: s[k:k] = bytes(random.randrange(256) for _ in range(random.randint(1, 8)))
            else: del s[k:]
            if not s: s = bytearray(b"\xff\xd8")
        f.write(struct.pack("<I", len(s)))
        f.write(s)
