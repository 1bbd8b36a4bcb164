"""CPU-only: Annex-B splitting and the RFC 7798 packetiser / depacketiser (include/b200_rtp.h), the
glue uvgRTP provides between KvazaarFilter and OpenHEVCFilter in the reference
(uvgrtpsender.cpp:89-118, uvgrtpreceiver.cpp:54-116).  Byte shuffling only -- no GPU."""
import struct

import numpy as np
import pytest

from kvazzup_b200 import rtp
from kvazzup_b200.openhevc import split_nals


def fake_nal(nal_type, size, seed, tid=1):
    rng = np.random.default_rng(seed)
    body = rng.integers(1, 256, size=size, dtype=np.uint8).tobytes()      # no zero bytes: no accidental start codes
    return bytes([nal_type << 1, tid]) + body


def fake_au(sizes, seed=0, idr=True):
    types = ([32, 33, 34, 19] if idr else [1])
    au = b""
    for k, n in enumerate(sizes):
        t = types[k] if k < len(types) else 1
        au += (b"\0\0\0\1" if k % 2 == 0 else b"\0\0\1") + fake_nal(t, n, seed + k)
    return au


def test_annexb_split_matches_the_python_splitter_and_handles_both_start_codes():
    au = fake_au([20, 40, 7, 5000])
    nals = rtp.annexb_split(au)
    assert [n[0] >> 1 for n in nals] == [32, 33, 34, 19]
    assert [b"\0\0\0\1" + n for n in nals] == split_nals(au)
    assert rtp.annexb_split(b"") == [] and rtp.annexb_split(b"\1\2\3\4\5") == []
    assert rtp.annexb_split(b"\0\0\1") == []
    assert rtp.annexb_split(b"junk\0\0\1\x40\x01\xaa\0\0\0\0\1\x42\x01\xbb") == [b"\x40\x01\xaa", b"\x42\x01\xbb"]


def test_intra_inter_probes_follow_filter_cpp():
    assert rtp.is_hevc_intra(b"\0\0\0\1" + fake_nal(19, 10, 1)) and not rtp.is_hevc_inter(b"\0\0\0\1" + fake_nal(19, 10, 1))
    assert rtp.is_hevc_inter(b"\0\0\0\1" + fake_nal(1, 10, 1))
    assert not rtp.is_hevc_intra(b"\0\0\1" + fake_nal(19, 10, 1))         # 3-byte start code: the reference checks 4 bytes
    assert not rtp.is_hevc_intra(b"\0\0")


@pytest.mark.parametrize("max_payload", [1460, 100, 9, 4])
def test_packets_follow_rfc7798_and_reassemble(max_payload):
    s, r = rtp.RtpSender(0xCAFE, 96, max_payload), rtp.RtpReceiver(0xCAFE)
    got, seq_expect = [], 0
    for f in range(4):
        au = fake_au([23, 41, 8, 3000 + 977 * f] if f % 2 == 0 else [max_payload, max_payload + 1, 1, 2 * max_payload], seed=10 * f, idr=f % 2 == 0)
        pkts = s.push_frame(au, 3000 * f)
        for k, p in enumerate(pkts):
            b0, b1, seq, ts, ssrc = struct.unpack(">BBHII", p[:12])
            assert b0 == 0x80 and (b1 & 0x7f) == 96 and seq == seq_expect & 0xffff and ts == 3000 * f and ssrc == 0xCAFE
            assert (b1 >> 7) == (1 if k == len(pkts) - 1 else 0)           # marker on the last packet of the AU
            assert len(p) - 12 <= max_payload
            ptype = (p[12] >> 1) & 63
            if ptype == 49:                                                # FU: never both S and E, FuType is a real NAL type
                assert len(p) - 12 > 3 and (p[14] >> 6) != 3 and (p[14] & 63) < 48
            seq_expect += 1
            for nal, nts, mk in r.receive(p):
                got.append((f, nal, nts, mk))
        nals = split_nals(au)
        mine = [g for g in got if g[0] == f]
        assert [g[1] for g in mine] == nals
        assert all(g[2] == 3000 * f for g in mine) and [g[3] for g in mine] == [0] * (len(nals) - 1) + [1]
    assert r.lost == 0
    s.close(); r.close()


def test_sequence_numbers_wrap_and_large_access_units_fit_the_bounds():
    s, r = rtp.RtpSender(1, 97, 1200), rtp.RtpReceiver(1)
    n = 0
    for f in range(40):                                                    # > 65536 packets in total
        au = fake_au([30, 30, 30, 2_000_000], seed=f)
        pkts = s.push_frame(au, f)
        n += len(pkts)
        out = [x for p in pkts for x in r.receive(p)]
        assert [o[0] for o in out] == split_nals(au)
    assert n > 65536 and r.lost == 0


def test_lost_fragment_drops_that_nal_only_and_wrong_ssrc_is_ignored():
    s, r = rtp.RtpSender(7, 96, 200), rtp.RtpReceiver(7)
    au = fake_au([20, 30, 10, 1500])
    pkts = s.push_frame(au, 0)
    assert len(pkts) == 3 + 8
    out = []
    for k, p in enumerate(pkts):
        if k == 6:
            continue                                                       # lose a middle fragment of the slice NAL
        out += r.receive(p)
    assert [o[0] for o in out] == split_nals(au)[:3] and r.lost == 1
    # the next access unit decodes normally
    au2 = fake_au([900], seed=5, idr=False)
    out2 = [x for p in s.push_frame(au2, 3000) for x in r.receive(p)]
    assert [o[0] for o in out2] == split_nals(au2)
    other = rtp.RtpSender(8, 96, 200)
    assert r.receive(other.push_frame(au2, 0)[0]) is None                  # wrong SSRC (uvgrtpreceiver.cpp:68-76)
    assert r.receive(b"\x80\x60") is None                                  # truncated
    # lose the FIRST fragment: the tail must not be mistaken for a NAL
    pk = s.push_frame(au2, 6000)
    out3 = [x for p in pk[1:] for x in r.receive(p)]
    assert out3 == [] and r.lost == 2


def test_aggregation_packets_from_other_senders_are_unpacked():
    r = rtp.RtpReceiver(9)
    a, b = fake_nal(33, 12, 1), fake_nal(34, 5, 2)
    payload = bytes([48 << 1, 1]) + struct.pack(">H", len(a)) + a + struct.pack(">H", len(b)) + b
    pkt = struct.pack(">BBHII", 0x80, 0x80 | 96, 5, 777, 9) + payload
    out = r.receive(pkt)
    assert [o[0] for o in out] == [b"\0\0\0\1" + a, b"\0\0\0\1" + b] and [o[2] for o in out] == [0, 1] and out[0][1] == 777
    assert r.receive(pkt[:-1]) is None                                     # truncated aggregation unit


def test_small_output_buffer_is_reported_and_consumes_nothing():
    import ctypes as C
    l = rtp._l()
    h = l.b200_rtp_sender_new(1, 96, 100)
    au = fake_au([500], idr=False)
    out, lens = (C.c_ubyte * 64)(), (C.c_uint32 * 64)()
    assert l.b200_rtp_push_frame(h, au, len(au), 0, out, 64, lens, 64) == -2
    big = (C.c_ubyte * 4096)()
    n = l.b200_rtp_push_frame(h, au, len(au), 0, big, 4096, lens, 64)
    assert n == 6 and bytes(big[2:4]) == b"\0\0"                           # first sequence number is still 0
    l.b200_rtp_sender_free(h)


def test_receiver_survives_garbage_and_truncations():
    """Packets come from the network: random bytes, truncated and bit-flipped packets must be
    rejected or parsed, never crash, and a clean packet afterwards still comes through."""
    rng = np.random.default_rng(2024)
    s, r = rtp.RtpSender(3, 96, 120), rtp.RtpReceiver(3)
    good = s.push_frame(fake_au([20, 30, 10, 700]), 0)
    for k in range(3000):
        kind = k % 3
        if kind == 0:
            pkt = rng.integers(0, 256, size=int(rng.integers(0, 200)), dtype=np.uint8).tobytes()
        elif kind == 1:
            p = good[int(rng.integers(0, len(good)))]
            pkt = p[:int(rng.integers(0, len(p) + 1))]
        else:
            p = bytearray(good[int(rng.integers(0, len(good)))])
            for _ in range(3):
                p[int(rng.integers(0, len(p)))] ^= 1 << int(rng.integers(0, 8))
            pkt = bytes(p)
        out = r.receive(pkt)
        assert out is None or all(n[0][:4] == b"\0\0\0\1" and len(n[0]) >= 6 for n in out)
    au = fake_au([900], seed=9, idr=False)
    out = [x for p in s.push_frame(au, 9000) for x in (r.receive(p) or [])]
    assert [o[0] for o in out][-1:] == split_nals(au)
    assert rtp.annexb_split(bytes(rng.integers(0, 2, size=5000, dtype=np.uint8))) is not None     # dense start codes


def test_csrc_list_and_padding_longer_than_the_packet_are_refused():
    """Found by tools/fuzz/rtp_harness.cpp under AddressSanitizer: a 17-byte packet that announces four CSRC entries
    (28-byte header) and a padding count larger than the packet made the payload offset wrap and the NAL header be
    read beyond the packet.  Such packets are malformed: -1, state untouched."""
    r = rtp.RtpReceiver(3)
    pkt = bytearray(17)
    pkt[0] = 0x80 | 0x20 | 4                        # version 2, padding, CC = 4
    pkt[8:12] = (3).to_bytes(4, "big")              # the expected SSRC
    pkt[16] = 200                                   # padding count
    assert r.receive(bytes(pkt)) is None
    pkt[0] = 0x80 | 0x10 | 15                       # extension flag with a CSRC list that already ends beyond the packet
    assert r.receive(bytes(pkt)) is None
    assert r.lost == 0


def test_rtp_shim_survives_the_sanitizer_fuzz(tmp_path):
    """The whole shim (packetiser, depacketiser, splitter, intra / inter probes) under AddressSanitizer +
    UndefinedBehaviorSanitizer: packets lost, truncated, bit-flipped, reordered, duplicated, relabelled as aggregation
    or fragmentation units, random datagrams, output buffers of random sizes."""
    import shutil
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = tmp_path / "rtp_harness"
    b = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        "-I", str(root / "include"), str(root / "tools/fuzz/rtp_harness.cpp"), str(root / "kvazzup_b200/csrc/rtp_shim.cpp"),
                        "-o", str(exe)], capture_output=True, text=True)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip("sanitizer runtime not available")
    assert b.returncode == 0, b.stderr
    out = subprocess.run([str(exe), "5", "250"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.stdout + out.stderr)[-2000:]
    assert out.stdout.startswith("rounds 250")
