"""Python mirror of the reference's OpenHEVCFilter on top of the libOpenHevc* C ABI
(include/b200_openhevc.h).

Follows /root/reference/src/media/processing/openhevcfilter.cpp: init() (:28-74), process() (:103-189,
one NAL per input buffer with a 4-byte start code, VCL NALs discarded until VPS+SPS+PPS were seen)
and sendDecodedOutput() (:192-239, strided planes repacked to packed I420).  `split_nals` stands in
for the uvgRTP receiver (src/media/delivery/uvgrtpreceiver.cpp:87-111), which hands the filter one
NAL per buffer.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import B200Error, lib

VPS_NUT, SPS_NUT, PPS_NUT = 32, 33, 34


class OpenHevcRational(C.Structure):
    _fields_ = [("num", C.c_int), ("den", C.c_int)]


class OpenHevcFrameInfo(C.Structure):
    _fields_ = [("nYPitch", C.c_int), ("nUPitch", C.c_int), ("nVPitch", C.c_int), ("nBitDepth", C.c_int),
                ("nWidth", C.c_int), ("nHeight", C.c_int), ("chromat_format", C.c_int),
                ("sample_aspect_ratio", OpenHevcRational), ("frameRate", OpenHevcRational),
                ("display_picture_number", C.c_int), ("flag", C.c_int), ("nTimeStamp", C.c_int64)]


class OpenHevcFrame(C.Structure):
    _fields_ = [("pvY", C.c_void_p), ("pvU", C.c_void_p), ("pvV", C.c_void_p), ("frameInfo", OpenHevcFrameInfo)]


def split_nals(au: bytes):
    """Annex-B access unit -> list of NAL units, each re-prefixed with a 4-byte start code."""
    out, i, n = [], 0, len(au)
    starts = []
    while True:
        j = au.find(b"\0\0\1", i)
        if j < 0:
            break
        starts.append(j + 3)
        i = j + 3
    for k, s in enumerate(starts):
        e = starts[k + 1] - 3 if k + 1 < len(starts) else n
        while e > s and au[e - 1] == 0:
            e -= 1
        out.append(b"\0\0\0\1" + au[s:e])
    return out


class StreamInfo(C.Structure):
    """b200_stream_info (include/b200_openhevc.h)."""
    _fields_ = [(n, C.c_int) for n in ("struct_size", "width", "height", "coded_width", "coded_height", "crop_left", "crop_top",
                                       "fps_num", "fps_den", "tile_cols", "tile_rows", "wpp", "sao", "sign_hiding", "qp_delta",
                                       "tmvp", "strong_intra", "cabac_init_present", "scaling_list", "max_tr_depth_inter",
                                       "max_tr_depth_intra", "max_dec_pic_buffering", "decodable")] + [("reason", C.c_char * 96)] + [
                    (n, C.c_int) for n in ("slices", "slice_type", "slice_qp", "entry_points", "num_ref_idx_l0", "rps_pictures")]


def probe(annexb: bytes, with_scaling_table: bool = False):
    """Host-only look at the parameter sets of a stream (b200_dec_probe): dict of the b200_stream_info fields, plus
    "scaling_table" (1552 bytes: 4 x 6 x 64 factors in raster order, 2 x 6 DC factors, padding) on request."""
    info = StreamInfo()
    info.struct_size = C.sizeof(StreamInfo)
    table = np.zeros(1552, np.uint8)
    buf = np.frombuffer(annexb, np.uint8)
    rc = lib().b200_dec_probe(C.c_void_p(buf.ctypes.data), buf.size, C.byref(info), C.c_void_p(table.ctypes.data) if with_scaling_table else None)
    if rc != 0:
        raise B200Error("b200_dec_probe failed: " + lib().b200_last_error().decode())
    out = {n: getattr(info, n) for n, _ in StreamInfo._fields_ if n not in ("struct_size", "reason")}
    out["reason"] = info.reason.decode()
    if with_scaling_table:
        out["scaling_table"] = table
    return out


class OpenHEVCFilter:
    def __init__(self, threads: int = 1, parallelization: str = "Slice"):
        self.l = lib()
        self.threads, self.mode = threads, parallelization
        self.handle = None
        self.vps = self.sps = self.pps = False
        self.discarded = 0

    def init(self) -> bool:
        ttype = {"Slice": 2, "Frame": 1}.get(self.mode, 3)      # OHThreadType, openhevcfilter.cpp:11
        self.handle = self.l.libOpenHevcInit(self.threads, ttype)
        if self.l.libOpenHevcStartDecoder(self.handle) == -1:
            return False
        self.l.libOpenHevcSetTemporalLayer_id(self.handle, 0)
        self.l.libOpenHevcSetActiveDecoders(self.handle, 0)
        self.l.libOpenHevcSetViewLayers(self.handle, 0)
        self.vps = self.sps = self.pps = False
        return True

    def version(self) -> str:
        return self.l.libOpenHevcVersion(self.handle).decode()

    def process(self, nal: bytes, pts: int = 0):
        """One NAL (4-byte start code) in; returns a decoded (i420, w, h) or None."""
        nal_type = nal[4] >> 1
        self.vps |= nal_type == VPS_NUT
        self.sps |= nal_type == SPS_NUT
        self.pps |= nal_type == PPS_NUT
        vcl = nal_type <= 31
        if not ((self.vps and self.sps and self.pps) or not vcl):
            self.discarded += 1
            return None
        buf = (C.c_ubyte * len(nal)).from_buffer_copy(nal)
        got = self.l.libOpenHevcDecode(self.handle, buf, len(nal), pts)
        if got <= -1:
            raise B200Error("libOpenHevcDecode failed: " + self.l.b200_last_error().decode())
        if got == 0:
            return None
        return self._send_decoded_output(got)

    def frame_rate(self):
        """(num, den) the decoder reports for the stream (OpenHevc_FrameInfo::frameRate)."""
        info = OpenHevcFrameInfo()
        self.l.libOpenHevcGetPictureInfo(self.handle, C.byref(info))
        return info.frameRate.num, info.frameRate.den

    def missing_refs(self) -> int:
        """P pictures decoded although their reference picture had been lost (concealed, drifting)."""
        return int(self.l.b200_dec_missing_refs(self.handle))

    def set_host_output(self, on: bool):
        """False: decoded pictures stay on the GPU; read them with output_dev()."""
        self.l.b200_dec_set_host_output(self.handle, int(on))

    def output_dev(self) -> int:
        """Device pointer (packed I420) of the picture the last successful process() call completed."""
        return self.l.b200_dec_output_dev(self.handle) or 0

    def process_dev(self, nal: bytes, pts: int = 0) -> int:
        """process() for a GPU consumer: returns the device pointer of the completed picture or 0."""
        buf = (C.c_ubyte * len(nal)).from_buffer_copy(nal)
        got = self.l.libOpenHevcDecode(self.handle, buf, len(nal), pts)
        if got <= -1:
            raise B200Error("libOpenHevcDecode failed: " + self.l.b200_last_error().decode())
        return self.output_dev() if got > 0 else 0

    def drain(self):
        """Held-back pictures of frame threading (not in the reference filter, which never drains:
        it flushes and closes, openhevcfilter.cpp:81-82)."""
        out = []
        while True:
            got = self.l.libOpenHevcDecode(self.handle, None, 0, 0)
            if got <= -1:
                raise B200Error("libOpenHevcDecode failed: " + self.l.b200_last_error().decode())
            if got == 0:
                return out
            out.append(self._send_decoded_output(got))

    def _send_decoded_output(self, got):
        fr = OpenHevcFrame()
        if self.l.libOpenHevcGetOutput(self.handle, got, C.byref(fr)) <= 0:
            return None
        self.l.libOpenHevcGetPictureInfo(self.handle, C.byref(fr.frameInfo))
        w, h = fr.frameInfo.nWidth, fr.frameInfo.nHeight
        out = np.empty(w * h * 3 // 2, np.uint8)
        s_stride, qs_stride = fr.frameInfo.nYPitch, fr.frameInfo.nUPitch // 2
        y = np.ctypeslib.as_array(C.cast(fr.pvY, C.POINTER(C.c_uint8)), shape=(h * s_stride,))
        u = np.ctypeslib.as_array(C.cast(fr.pvU, C.POINTER(C.c_uint8)), shape=(h * qs_stride,))
        v = np.ctypeslib.as_array(C.cast(fr.pvV, C.POINTER(C.c_uint8)), shape=(h * qs_stride,))
        out[:w * h] = y.reshape(h, s_stride)[:, :w].ravel()
        # the reference walks even luma rows i and reads chroma at i * (nUPitch / 2)  (:213-227)
        out[w * h:w * h + w * h // 4] = u.reshape(h // 2, 2 * qs_stride)[:, :w // 2].ravel()
        out[w * h + w * h // 4:] = v.reshape(h // 2, 2 * qs_stride)[:, :w // 2].ravel()
        return out, w, h

    def close(self):
        if self.handle:
            self.l.libOpenHevcFlush(self.handle)
            self.l.libOpenHevcClose(self.handle)
            self.handle = None
