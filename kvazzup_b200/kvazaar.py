"""Python mirror of the reference's KvazaarFilter on top of the kvz_api C ABI (include/b200_kvazaar.h).

Follows /root/reference/src/media/processing/kvazaarfilter.cpp step by step: init() (:122-311) maps
the uvgComm.ini settings keys to config_parse calls and direct kvz_config field writes, feedInput()
(:374-450) copies the I420 planes into a kvz_picture, calls encoder_encode and drains it with
pic=NULL, parseEncodedFrame() (:453-484) flattens the chunk list and frees chunks and recon.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import B200Error, lib

KVZ_DATA_CHUNK_SIZE = 4096


class KvzDataChunk(C.Structure):
    pass


KvzDataChunk._fields_ = [("data", C.c_uint8 * KVZ_DATA_CHUNK_SIZE), ("len", C.c_uint32),
                         ("next", C.POINTER(KvzDataChunk))]


class KvzConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "framerate_num", "framerate_denom", "qp", "intra_period", "vps_period", "wpp", "owf",
        "threads", "target_bitrate", "rc_algorithm", "lossless", "mv_constraint", "set_qp_in_cu", "hash",
        "deblock_enable", "sao_type", "tiles_width_count", "tiles_height_count", "slices", "vaq", "scaling_list",
        "gop_lowdelay", "gop_len", "me_range", "me_coarse", "subme_satd", "intra_satd", "return_recon", "device", "roi_enable")] + [("preset", C.c_char * 16)]


class KvzRoi(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("roi_array", C.POINTER(C.c_int8))]


class KvzPicture(C.Structure):
    pass


KvzPicture._fields_ = [("fulldata_buf", C.POINTER(C.c_uint8)), ("fulldata", C.POINTER(C.c_uint8)),
                       ("y", C.POINTER(C.c_uint8)), ("u", C.POINTER(C.c_uint8)), ("v", C.POINTER(C.c_uint8)),
                       ("data", C.POINTER(C.c_uint8) * 3), ("width", C.c_int32), ("height", C.c_int32),
                       ("stride", C.c_int32), ("base_image", C.POINTER(KvzPicture)), ("refcount", C.c_int32),
                       ("pts", C.c_int64), ("dts", C.c_int64), ("chroma_format", C.c_int), ("roi", KvzRoi)]


class KvzFrameInfo(C.Structure):
    _fields_ = [("poc", C.c_int32), ("qp", C.c_int8), ("nal_unit_type", C.c_int), ("slice_type", C.c_int),
                ("ref_list", (C.c_int * 16) * 2), ("ref_list_len", C.c_int * 2)]


_cfgp, _picp, _chunkp, _encp = C.POINTER(KvzConfig), C.POINTER(KvzPicture), C.POINTER(KvzDataChunk), C.c_void_p


class KvzApi(C.Structure):
    _fields_ = [
        ("config_alloc", C.CFUNCTYPE(_cfgp)),
        ("config_destroy", C.CFUNCTYPE(C.c_int, _cfgp)),
        ("config_init", C.CFUNCTYPE(C.c_int, _cfgp)),
        ("config_parse", C.CFUNCTYPE(C.c_int, _cfgp, C.c_char_p, C.c_char_p)),
        ("picture_alloc", C.CFUNCTYPE(_picp, C.c_int32, C.c_int32)),
        ("picture_free", C.CFUNCTYPE(None, _picp)),
        ("chunk_free", C.CFUNCTYPE(None, _chunkp)),
        ("encoder_open", C.CFUNCTYPE(_encp, _cfgp)),
        ("encoder_close", C.CFUNCTYPE(None, _encp)),
        ("encoder_headers", C.CFUNCTYPE(C.c_int, _encp, C.POINTER(_chunkp), C.POINTER(C.c_uint32))),
        ("encoder_encode", C.CFUNCTYPE(C.c_int, _encp, _picp, C.POINTER(_chunkp), C.POINTER(C.c_uint32),
                                       C.POINTER(_picp), C.POINTER(_picp), C.POINTER(KvzFrameInfo))),
        ("picture_alloc_csp", C.CFUNCTYPE(_picp, C.c_int, C.c_int32, C.c_int32)),
    ]


def kvz_api_get(bit_depth: int = 8):
    l = lib()
    l.kvz_api_get.restype = C.POINTER(KvzApi)
    l.kvz_api_get.argtypes = [C.c_int]
    p = l.kvz_api_get(bit_depth)
    return p.contents if p else None


# uvgComm.ini defaults (reference src/ui/settings/defaultsettings.cpp:266-281)
DEFAULT_SETTINGS = {
    "video/Preset": "veryfast", "video/ResolutionWidth": 1280, "video/ResolutionHeight": 720,
    "video/FramerateNumerator": 30, "video/FramerateDenominator": 1, "video/kvzThreads": "auto",
    "video/OWF": 0, "video/WPP": 1, "video/Tiles": 0, "video/tileDimensions": "2x2", "video/Slices": 0,
    "video/QP": 32, "video/Intra": 64, "video/VPS": 1, "video/bitrate": 0, "video/rcAlgorithm": "lambda",
    "video/scalingList": 0, "video/lossless": 0, "video/mvConstraint": "none", "video/qpInCU": 0, "video/vaq": 0,
    "parameters": [],
}


class KvazaarFilter:
    """Same life cycle as the reference filter: init() -> feed_input()* -> close()."""

    def __init__(self, settings: dict | None = None):
        self.settings = dict(DEFAULT_SETTINGS)
        if settings:
            self.settings.update(settings)
        self.api = None
        self.config = None
        self.enc = None
        self.input_pics = []
        self.next_input_pic = -1
        self.pts = 0
        self.warnings = []

    # kvazaarfilter.cpp:122-311
    def init(self) -> bool:
        s = self.settings
        if not (s["video/ResolutionWidth"] and s["video/ResolutionHeight"] and s["video/FramerateNumerator"]
                and s["video/FramerateDenominator"]):
            return False
        self.api = kvz_api_get(8)
        if self.api is None:
            return False
        cfg = self.api.config_alloc()
        if not cfg:
            return False
        self.config = cfg
        api = self.api
        api.config_init(cfg)

        def parse(name, value):
            rc = api.config_parse(cfg, name.encode(), str(value).encode())
            if rc != 1:
                self.warnings.append((name, str(value)))
            return rc

        parse("preset", s["video/Preset"])
        parse("input-res", f'{s["video/ResolutionWidth"]}x{s["video/ResolutionHeight"]}')
        parse("input-fps", f'{int(s["video/FramerateNumerator"])}/{int(s["video/FramerateDenominator"])}')
        threads = s["video/kvzThreads"]
        parse("threads", {"auto": "8", "Main": "0"}.get(str(threads), threads))
        parse("owf", s["video/OWF"])
        parse("wpp", s["video/WPP"])
        if s["video/Tiles"]:
            parse("tiles", s["video/tileDimensions"])
        if int(s["video/Slices"]) == 1:
            parse("slices", "wpp" if cfg.contents.wpp else "tiles")
        parse("qp", s["video/QP"])
        parse("period", s["video/Intra"])
        parse("vps-period", s["video/VPS"])
        cfg.contents.target_bitrate = int(s["video/bitrate"])
        if cfg.contents.target_bitrate != 0:
            parse("rc-algorithm", s["video/rcAlgorithm"])
        parse("intra-bits", "")
        parse("gop", "lp-g4d3t1")
        parse("scaling-list", "off" if int(s["video/scalingList"]) == 0 else "default")
        cfg.contents.lossless = int(s["video/lossless"])
        constraint = s["video/mvConstraint"]
        parse("mv-constraint", "" if constraint in ("frame", "frametile", "frametilemargin") else "none")
        cfg.contents.mv_constraint = {"frame": 1, "tile": 2, "frametile": 3, "frametilemargin": 4}.get(constraint, 0)
        cfg.contents.set_qp_in_cu = int(s["video/qpInCU"])
        if 0 < int(s["video/vaq"]) <= 20:
            parse("vaq", s["video/vaq"])
        for name, value in s["parameters"]:
            parse(name, value)
        cfg.contents.hash = 0
        self.enc = api.encoder_open(cfg)
        if not self.enc:
            return False
        self.input_pics = [api.picture_alloc(cfg.contents.width, cfg.contents.height) for _ in range(cfg.contents.owf + 1)]
        self.next_input_pic = 0
        return all(bool(p) for p in self.input_pics)

    def close(self):
        if self.api:
            if self.enc:
                self.api.encoder_close(self.enc)
            for p in self.input_pics:
                self.api.picture_free(p)
            if self.config:
                self.api.config_destroy(self.config)
        self.enc = None
        self.config = None
        self.input_pics = []
        self.api = None

    def _drain(self, pic, drain=True):
        api = self.api
        out = []
        data_out = _chunkp()
        len_out = C.c_uint32(0)
        recon = _picp()
        info = KvzFrameInfo()
        if api.encoder_encode(self.enc, pic, C.byref(data_out), C.byref(len_out), C.byref(recon), None, C.byref(info)) != 1:
            raise B200Error("encoder_encode failed: " + lib().b200_last_error().decode())
        while data_out:
            out.append(self._parse_encoded_frame(data_out, len_out.value, recon))
            if not drain:
                break                     # poll, do not drain, to keep owf pictures in flight
            data_out = _chunkp()
            recon = _picp()
            if api.encoder_encode(self.enc, None, C.byref(data_out), C.byref(len_out), C.byref(recon), None, C.byref(info)) != 1:
                raise B200Error("encoder_encode failed: " + lib().b200_last_error().decode())
        return out

    # kvazaarfilter.cpp:374-450
    def feed_input(self, i420: np.ndarray, drain: bool = True, roi: np.ndarray | None = None):
        """Returns the list of access units that became available (0 or more).  drain=True is the
        reference's loop (encoder_encode(pic=NULL) after every output, kvazaarfilter.cpp:440-449);
        drain=False is the one-line variant of INTEGRATION.md that keeps the pipeline full.
        roi: int8 delta-QP map (rows x cols, any resolution; the ROI filters make it per pixel),
        attached like vInfo->roi when the bitrate is 0 (:423-431)."""
        c = self.config.contents
        w, h = c.width, c.height
        assert i420.size == w * h * 3 // 2
        pic = self.input_pics[self.next_input_pic]
        self.next_input_pic = (self.next_input_pic + 1) % len(self.input_pics)
        src = np.ascontiguousarray(i420)
        C.memmove(pic.contents.y, src.ctypes.data, w * h)
        C.memmove(pic.contents.u, src.ctypes.data + w * h, w * h // 4)
        C.memmove(pic.contents.v, src.ctypes.data + w * h + w * h // 4, w * h // 4)
        pic.contents.pts = self.pts
        self.pts += 1
        if c.target_bitrate == 0 and roi is not None:
            self._roi = np.ascontiguousarray(roi, dtype=np.int8)          # caller-owned until the output arrives
            pic.contents.roi.width, pic.contents.roi.height = self._roi.shape[1], self._roi.shape[0]
            pic.contents.roi.roi_array = self._roi.ctypes.data_as(C.POINTER(C.c_int8))
        else:
            pic.contents.roi.width = pic.contents.roi.height = 0
            pic.contents.roi.roi_array = None
        return self._drain(pic, drain)

    def alloc_pictures(self, frames):
        """Page-locked pictures from picture_alloc holding `frames` (what a producer that writes its
        output straight into the encoder's input ring leaves behind): feed them with feed_picture()."""
        c = self.config.contents
        w, h = c.width, c.height
        pics = []
        for f in frames:
            pic = self.api.picture_alloc(w, h)
            if not pic:
                raise B200Error("picture_alloc failed")
            src = np.ascontiguousarray(f)
            C.memmove(pic.contents.y, src.ctypes.data, w * h * 3 // 2)       # planes are contiguous (b200_kvazaar.h)
            pics.append(pic)
        return pics

    def free_pictures(self, pics):
        for p in pics:
            self.api.picture_free(p)

    def feed_picture(self, pic, drain: bool = True):
        """encoder_encode on a picture that already sits in picture_alloc memory: no copy on the way in.
        The picture must stay untouched until its access unit has been returned (b200_kvazaar.h)."""
        pic.contents.pts = self.pts
        self.pts += 1
        pic.contents.roi.width = pic.contents.roi.height = 0
        pic.contents.roi.roi_array = None
        return self._drain(pic, drain)

    def flush(self):
        """Drain the frames still in flight (owf > 0): encoder_encode(pic = NULL) until empty."""
        out = []
        while True:
            got = self._drain(None)
            if not got:
                return out
            out += got

    # kvazaarfilter.cpp:453-484
    def _parse_encoded_frame(self, data_out, len_out, recon):
        buf = bytearray(len_out)
        off = 0
        chunk = data_out
        while chunk:
            n = chunk.contents.len
            buf[off:off + n] = bytes(chunk.contents.data[:n]) if n < 64 else C.string_at(chunk.contents.data, n)
            off += n
            chunk = chunk.contents.next
        self.api.chunk_free(data_out)
        self.api.picture_free(recon)
        assert off == len_out
        return bytes(buf)
