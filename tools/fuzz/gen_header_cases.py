"""Mutated oracle streams for tools/fuzz/headers_harness.cpp: records {u32 length, bytes}.  Bit flips, byte
replacements, deletions, insertions and runs of escaped zero bytes (huge Exp-Golomb values) in and around the
parameter sets and slice headers.

  python tools/fuzz/gen_header_cases.py cases.bin 60000 [seed]
"""
import random
import struct
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle.encoder import OracleEncoder, OracleTiledEncoder
from tests.test_oracle_hevc import frames_of
random.seed(int(sys.argv[3]) if len(sys.argv) > 3 else 7)
streams = []
for kw in ({}, {"refs": 3, "tmvp": 1, "sao": 2, "qp_delta": 1, "scaling_list": 2}, {"scaling_list": 3, "conf_right": 4, "fps_num": 30, "fps_den": 1, "cabac_init": 1},
           {"refs": 4, "tr_depth": 2, "sign_hiding": 1, "cb_qp_offset": 3, "beta_offset_div2": 2, "tc_offset_div2": -1, "no_wpp": 1}):
    enc = OracleEncoder(192, 136, qp=30, intra_period=0, **kw)
    streams.append(b"".join(enc.encode(f)[:700] for f in frames_of("camera", 192, 136, 5)))
    enc.close()
enc = OracleTiledEncoder(640, 256, 3, tile_rows=2, qp=30, intra_period=0, wpp=1)
streams.append(b"".join(enc.encode(f)[:900] for f in frames_of("camera", 640, 256, 2)))
enc.close()
with open(sys.argv[1], "wb") as f:
    for it in range(int(sys.argv[2])):
        s = bytearray(random.choice(streams))
        for _ in range(random.randint(1, 8)):
            mode = random.random(); k = random.randrange(len(s))
            if random.random() < 0.7: k = random.randrange(min(len(s), 300))
            if mode < 0.6: s[k] ^= 1 << random.randrange(8)
            elif mode < 0.8: s[k] = random.randrange(256)
            elif mode < 0.9: del s[k:k + random.randint(1, 8)]
            elif mode < 0.95: s[k:k] = bytes(random.randrange(256) for _ in range(random.randint(1, 6)))
            else: s[k:k] = b"\x00\x00\x03" * random.randint(1, 3) + bytes([random.randrange(256), random.randrange(256)])
        f.write(struct.pack("<I", len(s))); f.write(s)
