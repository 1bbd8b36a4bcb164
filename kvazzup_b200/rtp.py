"""ctypes binding of include/b200_rtp.h: Annex-B splitting and the RFC 7798 packetiser /
depacketiser that stand in for uvgRTP between the encoder and decoder filters
(/root/reference/src/media/delivery/uvgrtpsender.cpp:89-118 push_frame,
uvgrtpreceiver.cpp:54-116 receiveHook: one NAL per buffer, 4-byte start code prepended)."""
from __future__ import annotations

import ctypes as C

from .capi import B200Error, lib

v, i, sz, u32 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint32


class NalSpan(C.Structure):
    _fields_ = [("offset", u32), ("length", u32)]


_SIGS = {
    "b200_annexb_split": (i, [v, sz, v, i]),
    "b200_is_hevc_intra": (i, [v, sz]),
    "b200_is_hevc_inter": (i, [v, sz]),
    "b200_rtp_sender_new": (v, [u32, i, i]),
    "b200_rtp_sender_free": (None, [v]),
    "b200_rtp_push_frame": (i, [v, v, sz, u32, v, sz, v, i]),
    "b200_rtp_bound_bytes": (sz, [v, sz]),
    "b200_rtp_bound_packets": (i, [v, sz]),
    "b200_rtp_receiver_new": (v, [u32]),
    "b200_rtp_receiver_free": (None, [v]),
    "b200_rtp_receive": (i, [v, v, sz]),
    "b200_rtp_next_nal": (i, [v, v, sz, v, v]),
    "b200_rtp_receiver_lost": (C.c_uint, [v]),
}
_bound = False


def _l():
    global _bound
    l = lib()
    if not _bound:
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return l


def annexb_split(au: bytes):
    """NAL units of an Annex-B buffer, without start codes."""
    l = _l()
    n = l.b200_annexb_split(au, len(au), None, 0)
    if n < 0:
        raise B200Error("b200_annexb_split: bad arguments")
    spans = (NalSpan * max(n, 1))()
    l.b200_annexb_split(au, len(au), spans, n)
    return [au[s.offset:s.offset + s.length] for s in spans[:n]]


def is_hevc_intra(buf: bytes) -> bool:
    return bool(_l().b200_is_hevc_intra(buf, len(buf)))


def is_hevc_inter(buf: bytes) -> bool:
    return bool(_l().b200_is_hevc_inter(buf, len(buf)))


class RtpSender:
    def __init__(self, ssrc: int = 0x1234, payload_type: int = 96, max_payload: int = 1460):
        self.l = _l()
        self.h = self.l.b200_rtp_sender_new(ssrc, payload_type, max_payload)
        if not self.h:
            raise B200Error("b200_rtp_sender_new failed")

    def push_frame(self, au: bytes, rtp_timestamp: int):
        """One access unit -> list of RTP packets (bytes)."""
        cap = self.l.b200_rtp_bound_bytes(self.h, len(au))
        maxp = self.l.b200_rtp_bound_packets(self.h, len(au))
        out = (C.c_ubyte * cap)()
        lens = (u32 * maxp)()
        n = self.l.b200_rtp_push_frame(self.h, au, len(au), rtp_timestamp & 0xffffffff, out, cap, lens, maxp)
        if n < 0:
            raise B200Error(f"b200_rtp_push_frame failed ({n})")
        raw, pkts, off = bytes(out), [], 0
        for k in range(n):
            pkts.append(raw[off:off + lens[k]])
            off += lens[k]
        return pkts

    def close(self):
        if self.h:
            self.l.b200_rtp_sender_free(self.h)
            self.h = None


class RtpReceiver:
    def __init__(self, ssrc: int = 0x1234):
        self.l = _l()
        self.h = self.l.b200_rtp_receiver_new(ssrc)
        self.buf = (C.c_ubyte * (1 << 22))()

    def receive(self, pkt: bytes):
        """Feeds one packet; returns the NAL units (4-byte start code + NAL, rtp timestamp, marker) completed by it."""
        rc = self.l.b200_rtp_receive(self.h, pkt, len(pkt))
        if rc < 0:
            return None
        out = []
        ts, mk = u32(), i()
        while True:
            n = self.l.b200_rtp_next_nal(self.h, self.buf, len(self.buf), C.byref(ts), C.byref(mk))
            if n == -2:                      # NAL larger than the buffer: ts holds the size needed
                self.buf = (C.c_ubyte * max(ts.value, 2 * len(self.buf)))()
                continue
            if n <= 0:
                break
            out.append((bytes(self.buf[:n]), ts.value, mk.value))
        return out

    @property
    def lost(self) -> int:
        return self.l.b200_rtp_receiver_lost(self.h)

    def close(self):
        if self.h:
            self.l.b200_rtp_receiver_free(self.h)
            self.h = None
