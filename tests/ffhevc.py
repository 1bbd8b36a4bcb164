"""Independent conformance decoder for the tests: FFmpeg's native HEVC decoder, driven through
ctypes from the libavcodec that ships inside the cv2 wheel of this image.

Test infrastructure only.  HEVC decoding is normatively bit-exact, so equality between this
decoder's output and an encoder's own reconstruction pins every normative kernel of that encoder.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_state = {}


def _find(name: str) -> str | None:
    try:
        import cv2
    except Exception:
        return None
    base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    hits = sorted(glob.glob(os.path.join(base, f"lib{name}-*.so*")))
    return hits[0] if hits else None


def available() -> bool:
    return _find("avcodec") is not None and _find("avutil") is not None


def required() -> bool:
    """For the `-m gpu` tests: the independent decoder must take part.  A missing cv2 wheel FAILS the
    test (it would otherwise silently skip the one check that does not come from this repository)
    unless B200_ALLOW_NO_FFMPEG=1 says the box is known not to have it."""
    if available():
        return True
    if os.environ.get("B200_ALLOW_NO_FFMPEG") == "1":
        import warnings
        warnings.warn("FFmpeg (cv2 wheel) absent: independent-decoder check SKIPPED by B200_ALLOW_NO_FFMPEG=1")
        return False
    raise AssertionError("FFmpeg libavcodec (cv2 wheel) not found: the independent-decoder check cannot run "
                         "(set B200_ALLOW_NO_FFMPEG=1 to skip it knowingly)")


def _libs():
    if _state:
        return _state["avc"], _state["avu"]
    avu = C.CDLL(_find("avutil"), mode=C.RTLD_GLOBAL)
    avc = C.CDLL(_find("avcodec"), mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    avc.avcodec_find_decoder_by_name.restype = vp
    avc.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    avc.avcodec_alloc_context3.restype = vp
    avc.avcodec_alloc_context3.argtypes = [vp]
    avc.avcodec_open2.argtypes = [vp, vp, vp]
    avc.avcodec_send_packet.argtypes = [vp, vp]
    avc.avcodec_receive_frame.argtypes = [vp, vp]
    avc.av_packet_alloc.restype = vp
    avc.av_packet_free.argtypes = [C.POINTER(vp)]
    avc.avcodec_free_context.argtypes = [C.POINTER(vp)]
    avu.av_frame_alloc.restype = vp
    avu.av_frame_free.argtypes = [C.POINTER(vp)]
    avu.av_frame_unref.argtypes = [vp]
    avu.av_opt_set.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    avu.av_log_set_level.argtypes = [C.c_int]
    _state.update(avc=avc, avu=avu)
    return avc, avu


class _AVFrameHead(C.Structure):      # stable leading fields of AVFrame
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


class _AVPacketHead(C.Structure):     # stable leading fields of AVPacket
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p),
                ("size", C.c_int), ("stream_index", C.c_int)]


class HevcDecoder:
    """Feed Annex-B access units, get packed I420 frames (np.uint8) back."""

    def __init__(self, quiet: bool = False):
        avc, avu = _libs()
        avu.av_log_set_level(-8 if quiet else 16)     # AV_LOG_QUIET / AV_LOG_ERROR
        codec = avc.avcodec_find_decoder_by_name(b"hevc")
        if not codec:
            raise RuntimeError("FFmpeg hevc decoder not present")
        self.ctx = C.c_void_p(avc.avcodec_alloc_context3(codec))
        avu.av_opt_set(self.ctx, b"err_detect", b"crccheck+bitstream+buffer+explode", 0)
        avu.av_opt_set(self.ctx, b"threads", b"1", 0)
        if avc.avcodec_open2(self.ctx, codec, None) < 0:
            raise RuntimeError("avcodec_open2 failed")
        self.pkt = C.c_void_p(avc.av_packet_alloc())
        self.frame = C.c_void_p(avu.av_frame_alloc())
        self.errors = 0

    def _drain(self, out):
        avc, avu = _libs()
        while True:
            rc = avc.avcodec_receive_frame(self.ctx, self.frame)
            if rc < 0:
                break
            fr = _AVFrameHead.from_address(self.frame.value)
            w, h = fr.width, fr.height
            if fr.format != 0:                         # AV_PIX_FMT_YUV420P
                raise RuntimeError(f"unexpected pixel format {fr.format}")
            buf = np.empty(w * h * 3 // 2, np.uint8)
            off = 0
            for c, (pw, ph) in enumerate(((w, h), (w // 2, h // 2), (w // 2, h // 2))):
                ls = fr.linesize[c]
                src = np.ctypeslib.as_array(C.cast(fr.data[c], C.POINTER(C.c_uint8)), shape=(ph * ls,))
                buf[off:off + pw * ph] = src.reshape(ph, ls)[:, :pw].ravel()
                off += pw * ph
            out.append((buf, w, h))
            avu.av_frame_unref(self.frame)

    def decode(self, au: bytes):
        """Returns a list of (i420, w, h) frames that became available."""
        avc, _ = _libs()
        data = np.frombuffer(au + b"\0" * 64, np.uint8).copy()
        pk = _AVPacketHead.from_address(self.pkt.value)
        pk.data = data.ctypes.data
        pk.size = len(au)
        out = []
        rc = avc.avcodec_send_packet(self.ctx, self.pkt)
        if rc < 0:
            self.errors += 1
        self._drain(out)
        pk.data = None
        pk.size = 0
        return out

    def flush(self):
        avc, _ = _libs()
        out = []
        avc.avcodec_send_packet(self.ctx, None)
        self._drain(out)
        return out

    def close(self):
        avc, avu = _libs()
        if self.ctx:
            avu.av_frame_free(C.byref(self.frame))
            avc.av_packet_free(C.byref(self.pkt))
            avc.avcodec_free_context(C.byref(self.ctx))
            self.ctx = None


def decode_stream(aus, quiet: bool = False):
    """Decode a list of access units; returns (frames, errors)."""
    d = HevcDecoder(quiet=quiet)
    frames = []
    for au in aus:
        frames += d.decode(bytes(au))
    frames += d.flush()
    errs = d.errors
    d.close()
    return frames, errs
