"""CPU-only: pin the HEVC oracle.

* normative parts: the oracle encoder's streams are decoded by FFmpeg's native HEVC decoder
  (independent implementation) and must equal the oracle's own reconstruction bit for bit;
  the MD5 picture-hash SEI must verify under err_detect=crccheck+explode.
* primitives: known-answer tests (DC block, impulse, spec worked values).
"""
import ctypes as C

import numpy as np
import pytest

from kvazzup_b200 import synth
from oracle.encoder import OracleEncoder
from tests import ffhevc
from tests.helpers import ptr

needs_ff = pytest.mark.skipif(not ffhevc.available(), reason="FFmpeg libavcodec (cv2 wheel) not present")


def frames_of(kind, w, h, n):
    out = []
    for t in range(n):
        if kind == "noise":
            out.append(synth.noise(100 + t, w * h * 3 // 2))
        elif kind == "screen":
            out.append(synth.screen_i420(w, h, t * 5))
        elif kind == "sports":
            out.append(synth.sports_i420(w, h, t))
        else:
            out.append(synth.camera_i420(w, h, t))
    return out


def encode_all(frames, w, h, **kw):
    enc = OracleEncoder(w, h, **kw)
    aus, recs = [], []
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
    enc.close()
    return aus, recs


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("camera", 64, 64, 1, 32, {}),                       # one CTU, I picture
    ("camera", 192, 136, 4, 32, {}),                     # partial CTUs (8-high bottom row), WPP, P pictures
    ("camera", 72, 200, 2, 27, {"deblock": 0}),          # deblocking disabled via PPS
    ("noise", 128, 72, 3, 0, {}),                        # QP 0: escape-coded levels
    ("noise", 128, 72, 3, 10, {}),
    ("noise", 128, 72, 3, 45, {}),
    ("camera", 128, 72, 3, 51, {}),                      # QP 51: mostly skip
    ("camera", 416, 240, 5, 27, {"hash_sei": 1}),        # MD5 SEI verified by the decoder
    ("screen", 416, 240, 5, 32, {}),                     # static content: merge / skip paths
    ("camera", 200, 200, 7, 30, {"intra_period": 3}),    # periodic IDR, POC reset
    ("camera", 256, 128, 4, 22, {"search_range": 16}),
    ("camera", 64, 8, 3, 37, {}),                        # single row of 8x8 CUs
])
def test_oracle_streams_decode_bit_exactly_in_ffmpeg(kind, w, h, n, qp, kw):
    frames = frames_of(kind, w, h, n)
    aus, recs = encode_all(frames, w, h, qp=qp, **({"intra_period": 0} | kw))
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i, (fr, fw, fh) in enumerate(dec):
        assert (fw, fh) == (w, h)
        assert np.array_equal(fr, recs[i]), f"frame {i}: decoder output differs from encoder reconstruction"


def roi_pattern(w, h, t, kind):
    """Per-CTU QP offsets: a moving low-QP window on a high-QP background, or extreme random values."""
    cols, rows = (w + 63) // 64, (h + 63) // 64
    if kind == "random":
        rng = np.random.default_rng(7 + t)
        return rng.integers(-30, 31, size=cols * rows).astype(np.int8)       # clipped to 0..51 by the encoder
    d = np.full((rows, cols), 6, np.int8)
    d[(t // 2) % rows, :] = -8
    d[:, (t + 1) % cols] = -3
    return d.ravel()


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,roi,kw", [
    ("camera", 192, 136, 4, 32, "window", {}),
    ("camera", 416, 240, 6, 27, "window", {"hash_sei": 1, "intra_period": 4}),
    ("camera", 416, 240, 5, 30, "random", {"hash_sei": 1}),            # deltas beyond +-26 wrap modulo 52
    ("noise", 256, 136, 3, 25, "random", {}),
    ("screen", 416, 240, 5, 40, "random", {}),                          # many CTUs with no coded residual: QP prediction chain
    ("camera", 128, 72, 3, 32, None, {}),                               # flag on, no offsets: delta 0 everywhere
])
def test_per_ctu_qp_streams_decode_bit_exactly_in_ffmpeg(kind, w, h, n, qp, roi, kw):
    """cu_qp_delta (ROI) path: QP prediction, delta binarisation, per-CU QP in dequantisation and
    deblocking are all normative -- FFmpeg's reconstruction must equal the oracle's."""
    frames = frames_of(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, qp_delta=1, **({"intra_period": 0} | kw))
    aus, recs, qps = [], [], []
    for t, f in enumerate(frames):
        if roi:
            enc.set_ctu_dqp(roi_pattern(w, h, t, roi))
        aus.append(enc.encode(f))
        recs.append(enc.recon())
        qps.append(enc.cu_map()["qp"].copy())
    enc.close()
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i, (fr, fw, fh) in enumerate(dec):
        assert np.array_equal(fr, recs[i]), f"frame {i}: decoder output differs from encoder reconstruction"
    if roi:
        assert len(np.unique(np.concatenate(qps))) > 2           # the offsets really reached the CUs
    else:
        assert all((q == qp).all() for q in qps)


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("camera", 416, 240, 4, 22, {"scaling_list": 1, "hash_sei": 1, "intra_period": 3}),
    ("noise", 256, 136, 3, 12, {"scaling_list": 1, "tr_depth": 2, "tu4": 1, "intra_sizes": 7}),       # every block size 4..32
    ("sports", 640, 480, 4, 30, {"scaling_list": 1, "sao": 2, "intra_in_p": 1, "sign_hiding": 1, "tr_depth": 1, "me_coarse": 16, "search_range": 4}),
    ("camera", 416, 240, 4, 27, {"scaling_list": 2, "hash_sei": 1, "intra_period": 3, "intra_sizes": 3}),     # lists coded in the SPS
    ("noise", 256, 136, 3, 17, {"scaling_list": 2, "tr_depth": 2, "tu4": 1, "intra_sizes": 7, "qp_delta": 1}),
    ("screen", 640, 200, 4, 32, {"scaling_list": 3, "tr_depth": 1, "intra_period": 2}),                       # ... in the PPS
    ("noise", 512, 136, 3, 22, {"scaling_list": 3, "tr_depth": 2, "tu4": 1, "intra_sizes": 7, "cb_qp_offset": 3, "cr_qp_offset": -4}),
])
def test_scaling_list_streams_decode_bit_exactly_in_ffmpeg(kind, w, h, n, qp, kw):
    """Scaling lists (7.3.4, 8.6.4.2): the default lists (Kvazaar --scaling-list default) and lists carried in
    the SPS / PPS (coded, copied from another list, inferred default; DC coefficients of 16x16 / 32x32) --
    FFmpeg's reconstruction equals the oracle's, which pins the tables, the list expansion and the
    per-coefficient dequantisation; the lists really change the stream."""
    frames = frames_of(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, **({"intra_period": 0} | kw))
    flat = OracleEncoder(w, h, qp=qp, **({"intra_period": 0} | kw | {"scaling_list": 0}))
    aus, recs = [], []
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
        flat.encode(f)
    assert not np.array_equal(flat.recon(), recs[-1])
    enc.close()
    flat.close()
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i, (fr, fw, fh) in enumerate(dec):
        assert np.array_equal(fr, recs[i]), f"frame {i}: decoder output differs from encoder reconstruction"


def pad_i420(f, w, h, W, H):
    """w x h I420 picture -> W x H by repeating the last column / row of every plane."""
    planes = [f[:w * h].reshape(h, w), f[w * h:w * h * 5 // 4].reshape(h // 2, w // 2), f[w * h * 5 // 4:].reshape(h // 2, w // 2)]
    out = [np.pad(p, ((0, (H >> (i > 0)) - p.shape[0]), (0, (W >> (i > 0)) - p.shape[1])), mode="edge") for i, p in enumerate(planes)]
    return np.concatenate([p.ravel() for p in out])


def crop_i420(f, W, H, w, h):
    """top-left w x h window of a W x H I420 picture"""
    planes = [f[:W * H].reshape(H, W)[:h, :w], f[W * H:W * H * 5 // 4].reshape(H // 2, W // 2)[:h // 2, :w // 2],
              f[W * H * 5 // 4:].reshape(H // 2, W // 2)[:h // 2, :w // 2]]
    return np.concatenate([p.ravel() for p in planes])


def odd_size_frames(kind, w, h, n):
    """n pictures of w x h (even, not multiples of 8) cut out of the generator's next multiple-of-8 size"""
    W, H = (w + 7) & ~7, (h + 7) & ~7
    return [crop_i420(f, W, H, w, h) for f in frames_of(kind, W, H, n)]


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("camera", 410, 234, 3, 30, {"hash_sei": 1}),
    ("screen", 638, 200, 3, 32, {"sao": 2}),
    ("sports", 416, 238, 4, 27, {"me_coarse": 16, "search_range": 4, "intra_in_p": 1, "intra_period": 3}),
    ("camera", 1366, 768, 2, 32, {"hash_sei": 1, "search_range": 6}),          # a laptop screen
])
def test_conformance_window_streams_decode_cropped_in_ffmpeg(kind, w, h, n, qp, kw):
    """Source sizes that are not multiples of 8: the encoder codes the picture padded by edge replication and
    the SPS carries a conformance window; FFmpeg returns w x h pictures equal to the cropped reconstruction."""
    W, H = (w + 7) & ~7, (h + 7) & ~7
    frames = odd_size_frames(kind, w, h, n)
    enc = OracleEncoder(W, H, qp=qp, conf_right=W - w, conf_bottom=H - h, **({"intra_period": 0} | kw))
    aus, recs = [], []
    for f in frames:
        aus.append(enc.encode(pad_i420(f, w, h, W, H)))
        recs.append(crop_i420(enc.recon(), W, H, w, h))
    enc.close()
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i, (fr, fw, fh) in enumerate(dec):
        assert (fw, fh) == (w, h)
        assert np.array_equal(fr, recs[i]), f"frame {i}"


def vectors_leaving_the_picture(cu, w, h):
    """Number of inter 8x8 units whose motion vector reads a sample outside the picture (interpolation taps included)."""
    cu = cu.reshape(h // 8, w // 8)
    y8, x8 = np.nonzero(cu["pred_mode"] == 0)
    n = 1 << cu["log2_size"][y8, x8].astype(int)
    x0, y0 = (x8 * 8) & ~(n - 1), (y8 * 8) & ~(n - 1)
    mx, my = cu["mvx"][y8, x8].astype(int), cu["mvy"][y8, x8].astype(int)
    fx, fy = np.where(mx & 7, 4, 0), np.where(my & 7, 4, 0)
    bad = (x0 + (mx >> 2) - fx < 0) | (x0 + n + (mx >> 2) + fx > w) | (y0 + (my >> 2) - fy < 0) | (y0 + n + (my >> 2) + fy > h)
    return int(bad.sum()), int(bad.size)


@needs_ff
def test_frame_motion_constraint_keeps_every_vector_inside_the_picture():
    """mv_edges = 15 (Kvazaar's mv-constraint frame): without it the sports sequence has vectors that leave the
    picture, with it none does; the stream still decodes bit-exactly."""
    w, h, n = 416, 240, 5
    frames = frames_of("sports", w, h, n)
    for edges in (0, 15):
        enc = OracleEncoder(w, h, qp=30, intra_period=0, me_coarse=16, search_range=4, intra_in_p=1, mv_edges=edges, hash_sei=1)
        aus, out = [], 0
        for f in frames:
            aus.append(enc.encode(f))
            out += vectors_leaving_the_picture(enc.cu_map(), w, h)[0]
        rec = enc.recon()
        enc.close()
        assert (out == 0) == (edges == 15), (edges, out)
        dec, errs = ffhevc.decode_stream(aus)
        assert errs == 0 and np.array_equal(dec[-1][0], rec)


def vaq_float_model(i420, w, h, strength):
    """Kvazaar's formula in floating point: strength * 0.1 * (ln(max(var_ctu, 4)) - ln(var_picture)),
    var = luma variance + the two chroma variances."""
    y = i420[:w * h].reshape(h, w).astype(np.float64)
    u = i420[w * h:w * h * 5 // 4].reshape(h // 2, w // 2).astype(np.float64)
    v = i420[w * h * 5 // 4:].reshape(h // 2, w // 2).astype(np.float64)
    fv = max(y.var() + u.var() + v.var(), 4.0)
    cols, rows = (w + 63) // 64, (h + 63) // 64
    out = np.zeros((rows, cols))
    for r in range(rows):
        for c in range(cols):
            lv = (y[r * 64:(r + 1) * 64, c * 64:(c + 1) * 64].var() + u[r * 32:(r + 1) * 32, c * 32:(c + 1) * 32].var() +
                  v[r * 32:(r + 1) * 32, c * 32:(c + 1) * 32].var())
            out[r, c] = strength * 0.1 * (np.log(max(lv, 4.0)) - np.log(fv))
    return np.clip(out, -12, 12).ravel()


def vaq_frames(kind, w, h, n):
    """Synthetic pictures whose CTUs differ in variance (the generators' texture is even): the left
    third loses 7/8 of its contrast, the right third gains noise."""
    rng = np.random.default_rng(w * 31 + h)
    out = []
    for f in frames_of(kind, w, h, n):
        y = f[:w * h].reshape(h, w).astype(np.int32)
        y[:, :w // 3] = (y[:, :w // 3] - 128) // 8 + 128
        y[:, 2 * w // 3:] += rng.integers(-40, 41, (h, w - 2 * w // 3))
        g = f.copy()
        g[:w * h] = np.clip(y, 0, 255).astype(np.uint8).ravel()
        out.append(g)
    return out


def oracle_vaq_offsets(i420, w, h, strength):
    import ctypes as C
    from oracle.binding import load
    out = np.zeros(((w + 63) // 64) * ((h + 63) // 64), np.int8)
    pic = np.ascontiguousarray(i420)
    load().orc_vaq_offsets(C.c_void_p(pic.ctypes.data), w, h, strength, C.c_void_p(out.ctypes.data))
    return out


@pytest.mark.parametrize("kind,w,h,strength", [
    ("camera", 416, 240, 10), ("screen", 640, 200, 5), ("noise", 256, 136, 20), ("sports", 1920, 1080, 7),
    ("camera", 72, 64, 1), ("flat", 128, 72, 10),
])
def test_vaq_offsets_follow_the_published_formula(kind, w, h, strength):
    """The integer restatement (exact variance fractions, Q8 log2) stays within rounding of Kvazaar's
    floating-point formula; a flat picture gives offset 0 everywhere."""
    if kind == "flat":
        pic = np.full(w * h * 3 // 2, 97, np.uint8)
        assert not oracle_vaq_offsets(pic, w, h, strength).any()
        return
    pic = vaq_frames(kind, w, h, 2)[1]
    got = oracle_vaq_offsets(pic, w, h, strength)
    want = vaq_float_model(pic, w, h, strength)
    assert np.abs(got - want).max() <= 0.5 + 0.02 * strength, (got, want)
    if strength >= 5:
        assert got.min() < got.max()                # flat and busy CTUs really move apart


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("camera", 416, 240, 5, 30, {"vaq": 10, "hash_sei": 1}),
    ("screen", 640, 200, 4, 35, {"vaq": 20, "sao": 2, "intra_period": 3}),
    ("sports", 416, 240, 4, 27, {"vaq": 5, "me_coarse": 16, "search_range": 4, "intra_in_p": 1, "roi": "window"}),
    ("camera", 416, 240, 3, 2, {"vaq": 20}),                             # QP - offset clips at 0
])
def test_vaq_streams_decode_bit_exactly_in_ffmpeg(kind, w, h, n, qp, kw):
    """Variance adaptive quantisation rides the cu_qp_delta path: the per-CTU QPs are those of
    orc_vaq_offsets (plus the ROI offsets), and FFmpeg reconstructs the same pictures."""
    kw = dict(kw)
    roi = kw.pop("roi", None)
    frames = vaq_frames(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, qp_delta=1, **({"intra_period": 0} | kw))
    cols = (w + 63) // 64
    aus, recs = [], []
    for t, f in enumerate(frames):
        d = roi_pattern(w, h, t, roi) if roi else np.zeros(cols * ((h + 63) // 64), np.int8)
        if roi:
            enc.set_ctu_dqp(d)
        aus.append(enc.encode(f))
        recs.append(enc.recon())
        want = np.clip(qp + d.astype(int) + oracle_vaq_offsets(f, w, h, kw["vaq"]), 0, 51)
        cu = enc.cu_map().reshape(h // 8, w // 8)
        coded = (cu["flags"] & 2) != 0                    # CUs with a residual carry their CTU's target QP
        ctu_of = (np.arange(h // 8)[:, None] // 8) * cols + np.arange(w // 8)[None, :] // 8
        assert np.array_equal(cu["qp"][coded], want[ctu_of][coded]), f"frame {t}"
    enc.close()
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i, (fr, fw, fh) in enumerate(dec):
        assert np.array_equal(fr, recs[i]), f"frame {i}: decoder output differs from encoder reconstruction"
    with pytest.raises(ValueError):
        OracleEncoder(w, h, qp=qp, vaq=5)                 # needs qp_delta


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,tiles,kw", [
    ("camera", 416, 240, 5, 30, 2, {"hash_sei": 1}),                  # 7 CTU columns -> tiles of 3 and 4
    ("camera", 416, 240, 6, 27, 3, {"hash_sei": 1, "intra_period": 4}),
    ("noise", 512, 136, 3, 20, 4, {}),                                  # every tile exactly two CTUs wide
    ("screen", 640, 200, 5, 35, 2, {"deblock": 0}),
    ("camera", 1920, 1080, 2, 32, 4, {"hash_sei": 1, "search_range": 12}),
    ("camera", 416, 240, 4, 30, 2, {"hash_sei": 1, "tile_rows": 2}),  # tile grids: motion confined vertically as well
    ("camera", 640, 256, 5, 27, 3, {"hash_sei": 1, "tile_rows": 2, "intra_period": 3}),
    ("noise", 512, 136, 3, 22, 1, {"tile_rows": 2}),
    ("sports", 640, 480, 5, 32, 2, {"hash_sei": 1, "tile_rows": 3, "me_coarse": 16, "search_range": 4, "sao": 2, "intra_in_p": 1}),
])
def test_tile_columns_as_independent_strips_decode_bit_exactly_in_ffmpeg(kind, w, h, n, qp, tiles, kw):
    """Tiles (PPS tile columns, uniform spacing, no loop filter across tiles) built from strip
    encoders whose motion never reaches across an interior tile edge: an independent decoder that
    knows nothing about the trick must reproduce the composite reconstruction."""
    from oracle.encoder import OracleTiledEncoder
    frames = frames_of(kind, w, h, n)
    enc = OracleTiledEncoder(w, h, tiles, qp=qp, **({"intra_period": 0} | kw))
    aus, recs = [], []
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
    enc.close()
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i, (fr, fw, fh) in enumerate(dec):
        assert (fw, fh) == (w, h)
        bad = np.flatnonzero(fr != recs[i])
        assert bad.size == 0, f"frame {i}: {bad.size} samples differ, first at {bad[:6]}"


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("camera", 192, 136, 4, 32, {}),
    ("camera", 416, 240, 6, 37, {"hash_sei": 1, "intra_period": 4}),
    ("noise", 256, 136, 3, 30, {"hash_sei": 1}),                       # band offsets on noise, large offsets
    ("screen", 416, 240, 5, 40, {}),                                    # sharp edges: edge offsets
    ("camera", 200, 72, 3, 45, {"deblock": 0}),                         # SAO without deblocking, partial CTUs
    ("camera", 1280, 720, 2, 35, {"hash_sei": 1, "search_range": 12}),
    ("camera", 416, 240, 6, 37, {"sao": 2, "hash_sei": 1, "intra_period": 4}),    # sao_merge_left / _up flags in use
    ("screen", 640, 256, 5, 40, {"sao": 2}),
    ("noise", 256, 136, 3, 51, {"sao": 2, "hash_sei": 1}),
])
def test_sao_streams_decode_bit_exactly_in_ffmpeg(kind, w, h, n, qp, kw):
    """Sample adaptive offset: per-CTU edge / band offsets after deblocking.  The
    syntax (slice flags, sao() per CTU), the categories at picture and CTU borders and the clipping
    are normative -- FFmpeg must reproduce the oracle's picture -- and it must actually help."""
    frames = frames_of(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, **({"intra_period": 0, "sao": 1} | kw))
    plain = OracleEncoder(w, h, qp=qp, **({"intra_period": 0} | {k: v for k, v in kw.items() if k not in ("hash_sei", "sao")}))
    aus, recs, gain = [], [], 0.0
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
        plain.encode(f)
        gain += synth.psnr(f[:w * h], recs[-1][:w * h]) - synth.psnr(f[:w * h], plain.recon()[:w * h])
    enc.close(); plain.close()
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i, (fr, fw, fh) in enumerate(dec):
        bad = np.flatnonzero(fr != recs[i])
        assert bad.size == 0, f"frame {i}: {bad.size} samples differ, first at {bad[:6]}"
    assert gain / n > -0.05                                              # never worse than without (it may choose "off")
    if kw.get("sao") == 2:                                               # merging changes the rate only, never the picture
        nomerge = OracleEncoder(w, h, qp=qp, **({"intra_period": 0} | kw | {"sao": 1}))
        sizes = [len(nomerge.encode(f)) for f in frames]
        assert np.array_equal(nomerge.recon(), recs[-1]) and sum(len(a) for a in aus) <= sum(sizes)


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("sports", 416, 240, 6, 32, {}),                                     # fast pan + scene cut: many intra CUs
    ("sports", 200, 136, 5, 22, {"hash_sei": 1}),                        # partial CTUs: intra only where 16x16 fits
    ("noise", 128, 72, 3, 30, {}),
    ("camera", 416, 240, 5, 37, {"sao": 2, "hash_sei": 1}),
    ("sports", 416, 240, 6, 35, {"sao": 1, "qp_delta": 1, "intra_period": 4}),
    ("sports", 640, 256, 4, 30, {"deblock": 0}),
])
def test_intra_cus_in_p_pictures_decode_bit_exactly_in_ffmpeg(kind, w, h, n, qp, kw):
    """cfg.intra_in_p: 16x16 intra CUs inside P pictures where inter prediction is poor.  Normative
    side: pred_mode_flag, the MPM derivation next to inter neighbours (they count as DC), intra
    prediction from reconstructed INTER samples, bS 2 deblocking -- FFmpeg must agree bit for bit."""
    frames = frames_of(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, intra_in_p=1, **({"intra_period": 0} | kw))
    if kw.get("qp_delta"):
        enc.set_ctu_dqp(roi_pattern(w, h, 1, "random"))
    aus, recs, n_intra = [], [], 0
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
        if not enc.last_was_idr():
            n_intra += int((enc.cu_map()["pred_mode"] == 1).sum())
    enc.close()
    if kind == "sports":
        assert n_intra > 0                                                # the option is actually exercised
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i in range(n):
        bad = np.flatnonzero(dec[i][0] != recs[i])
        assert bad.size == 0, f"frame {i}: {bad.size} samples differ, first at {bad[:6]}"


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("sports", 416, 240, 6, 32, {"me_coarse": 16, "search_range": 4}),
    ("sports", 200, 136, 5, 27, {"me_coarse": 16, "search_range": 8, "hash_sei": 1}),     # partial CTUs and quadrants
    ("camera", 416, 240, 5, 37, {"me_coarse": 8, "search_range": 4, "sao": 2, "intra_in_p": 1}),
    ("sports", 640, 256, 4, 30, {"me_coarse": 32, "search_range": 6, "intra_in_p": 1, "hash_sei": 1}),
])
def test_two_level_motion_search_streams_decode_in_ffmpeg(kind, w, h, n, qp, kw):
    """cfg.me_coarse: vectors far beyond search_range (up to 4 * me_coarse + search_range samples, also
    pointing outside the picture); the search is not normative, the motion compensation at those
    vectors is -- FFmpeg must agree -- and on fast motion it must beat the zero-centred window."""
    frames = frames_of(kind, w, h, n)
    aus, recs = encode_all(frames, w, h, qp=qp, **({"intra_period": 0} | kw))
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i in range(n):
        assert np.array_equal(dec[i][0], recs[i]), i
    if kind == "sports":
        plain, _ = encode_all(frames, w, h, qp=qp, intra_period=0, search_range=12)
        assert sum(map(len, aus)) < 0.9 * sum(map(len, plain))


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("sports", 416, 240, 7, 32, {"tmvp": 1}),
    ("sports", 416, 240, 7, 32, {"refs": 2}),
    ("camera", 192, 136, 6, 30, {"refs": 3, "tmvp": 1, "hash_sei": 1}),
    ("noise", 128, 72, 5, 32, {"refs": 4, "tmvp": 1}),
    ("sports", 416, 240, 8, 32, {"refs": 4, "tmvp": 1, "me_coarse": 16, "search_range": 4, "sao": 2, "intra_in_p": 1,
                                 "hash_sei": 1, "intra_period": 5}),
    ("sports", 640, 256, 5, 27, {"refs": 3, "tmvp": 1, "qp_delta": 1}),
])
def test_several_references_and_temporal_mv_prediction_decode_in_ffmpeg(kind, w, h, n, qp, kw):
    """cfg.refs / cfg.tmvp: what a Kvazaar peer's streams use (its `lp-g4d3t1` structure keeps several
    references; TMVP is on by default).  Normative: the reference picture set in the slice header while
    the buffer fills, list construction, ref_idx_l0, merge / AMVP candidates across different
    reference pictures (vector scaling by POC distance), collocated candidates from the compressed
    motion field, boundary strength across different references -- FFmpeg must agree bit for bit."""
    frames = frames_of(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, **({"intra_period": 0} | kw))
    if kw.get("qp_delta"):
        enc.set_ctu_dqp(roi_pattern(w, h, 1, "random"))
    aus, recs, max_ref = [], [], 0
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
        max_ref = max(max_ref, int(enc.cu_map()["ref_idx"].max()))
    enc.close()
    assert max_ref == min(kw.get("refs", 1), n - 1, (kw.get("intra_period") or n) - 1) - 1 or kw.get("refs", 1) == 1
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i in range(n):
        assert np.array_equal(dec[i][0], recs[i]), i


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("camera", 192, 136, 4, 30, {"tr_depth": 1}),
    ("sports", 416, 240, 5, 30, {"tr_depth": 2}),
    ("screen", 416, 240, 4, 30, {"tr_depth": 2, "hash_sei": 1}),
    ("noise", 128, 72, 3, 22, {"tr_depth": 1, "cabac_init": 1}),
    ("camera", 416, 240, 5, 32, {"tr_depth": 2, "cabac_init": 1, "sao": 2, "intra_in_p": 1, "refs": 2, "tmvp": 1, "qp_delta": 1,
                                 "hash_sei": 1, "intra_period": 3}),
])
def test_transform_tree_streams_decode_in_ffmpeg(kind, w, h, n, qp, kw):
    """cfg.tr_depth: split transform units (split_transform_flag, the cbf_cb / cbf_cr hierarchy, cbf_luma
    contexts by depth), intra prediction transform unit by transform unit, deblocking of transform
    edges inside CUs, cu_qp_delta in the first coded transform unit; cfg.cabac_init: initialisation
    type 2 for P slices.  All normative: FFmpeg must agree."""
    frames = frames_of(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, **({"intra_period": 0} | kw))
    if kw.get("qp_delta"):
        enc.set_ctu_dqp(roi_pattern(w, h, 1, "random"))
    aus, recs, split = [], [], 0
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
        m = enc.cu_map()
        split += int((m["tu_log2"] < m["log2_size"]).sum())
    enc.close()
    assert split > 0
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i in range(n):
        assert np.array_equal(dec[i][0], recs[i]), i


PEER_SYNTAX_CASES = [
    ("camera", 192, 136, 3, 30, {"tr_depth": 1, "tu4": 1}),                       # 4x4 luma blocks: DST (intra) / DCT (inter)
    ("noise", 128, 72, 3, 22, {"tr_depth": 2, "tu4": 1}),
    ("camera", 192, 136, 3, 30, {"intra_sizes": 1}),                              # 8x8 intra CUs by decision
    ("camera", 192, 136, 3, 30, {"intra_sizes": 2}),                              # 32x32 intra CUs
    ("camera", 192, 136, 3, 30, {"intra_sizes": 5}),                              # NxN partitions
    ("noise", 128, 72, 2, 27, {"intra_sizes": 7, "tr_depth": 2, "tu4": 1}),
    ("camera", 192, 136, 3, 30, {"chroma_modes": 1}),                             # explicit intra_chroma_pred_mode
    ("camera", 192, 136, 3, 27, {"sign_hiding": 1}),
    ("noise", 128, 72, 3, 22, {"sign_hiding": 1, "tr_depth": 2, "tu4": 1, "intra_sizes": 7, "chroma_modes": 1}),
    ("camera", 256, 256, 2, 30, {"intra_sizes": 2, "strong_intra": 1}),
    ("camera", 192, 136, 3, 30, {"cb_qp_offset": 3, "cr_qp_offset": -4}),
    ("camera", 192, 136, 3, 30, {"beta_offset_div2": 2, "tc_offset_div2": -3}),
    ("sports", 416, 240, 5, 30, {"tr_depth": 2, "tu4": 1, "intra_sizes": 7, "chroma_modes": 1, "sign_hiding": 1, "strong_intra": 1,
                                 "cb_qp_offset": 2, "cr_qp_offset": -2, "beta_offset_div2": 1, "tc_offset_div2": 1, "sao": 2,
                                 "intra_in_p": 1, "refs": 2, "tmvp": 1, "qp_delta": 1, "hash_sei": 1, "intra_period": 3, "cabac_init": 1}),
]


@needs_ff
@pytest.mark.parametrize("kind,w,h,n,qp,kw", PEER_SYNTAX_CASES)
def test_peer_syntax_streams_decode_in_ffmpeg(kind, w, h, n, qp, kw):
    """Syntax the GPU encoder does not produce but a Kvazaar-family peer may (streams for the decoder
    tests): 4x4 luma transform blocks with the DST, NxN / 8x8 / 32x32 intra CUs, explicit chroma
    modes, sign data hiding, strong intra smoothing, chroma QP and deblocking offsets.  All
    normative: FFmpeg must reproduce the oracle's reconstruction."""
    frames = frames_of(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, **({"intra_period": 0} | kw))
    if kw.get("qp_delta"):
        enc.set_ctu_dqp(roi_pattern(w, h, 1, "random"))
    aus, recs = [], []
    seen = {"tu4": 0, "nxn": 0, "cu8": 0, "cu32": 0, "chroma": 0}
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
        m = enc.cu_map()
        intra = m["pred_mode"] == 1
        seen["tu4"] += int((m["tu_log2"] == 2).sum())
        seen["nxn"] += int(((m["flags"] & 1) == 1).sum())
        seen["cu8"] += int(((m["log2_size"] == 3) & intra).sum())
        seen["cu32"] += int(((m["log2_size"] == 5) & intra).sum())
        seen["chroma"] += int(((m["chroma_mode"] != m["intra_mode"]) & intra).sum())
    enc.close()
    if kw.get("tu4"):
        assert seen["tu4"] > 0
    if kw.get("intra_sizes", 0) & 4:
        assert seen["nxn"] > 0
    if kw.get("intra_sizes", 0) == 2:
        assert seen["cu32"] > 0
    if kw.get("chroma_modes"):
        assert seen["chroma"] > 0
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i in range(n):
        assert np.array_equal(dec[i][0], recs[i]), i


@needs_ff
def test_sao_with_per_ctu_qp_and_periodic_idr_decodes_in_ffmpeg():
    """The two per-CTU syntax additions together: sao() precedes the coding quadtree, cu_qp_delta sits
    in the first coded transform unit; both follow the WPP context hand-over."""
    w, h, n = 416, 240, 6
    enc = OracleEncoder(w, h, qp=33, sao=1, qp_delta=1, hash_sei=1, intra_period=3)
    aus, recs = [], []
    for t, f in enumerate(frames_of("camera", w, h, n)):
        enc.set_ctu_dqp(roi_pattern(w, h, t, "random" if t % 2 else "window"))
        aus.append(enc.encode(f))
        recs.append(enc.recon())
    enc.close()
    dec, errs = ffhevc.decode_stream(aus)
    assert errs == 0 and len(dec) == n
    for i in range(n):
        assert np.array_equal(dec[i][0], recs[i]), i


def test_per_ctu_qp_changes_rate_where_asked():
    w, h = 416, 240
    frames = frames_of("camera", w, h, 3)
    sizes = {}
    for name, d in (("flat", 0), ("fine", -10), ("coarse", 10)):
        enc = OracleEncoder(w, h, qp=30, qp_delta=1, intra_period=0)
        enc.set_ctu_dqp(np.full(7 * 4, d, np.int8))
        sizes[name] = sum(len(enc.encode(f)) for f in frames)
        enc.close()
    assert sizes["fine"] > sizes["flat"] > sizes["coarse"]
    with pytest.raises(ValueError):
        OracleEncoder(w, h, qp=30).set_ctu_dqp(np.zeros(28, np.int8))


@needs_ff
def test_corrupted_hash_is_detected():
    """Guards the guard: the decoder really does check the MD5 SEI."""
    w, h = 64, 64
    aus, _ = encode_all(frames_of("camera", w, h, 1), w, h, qp=30, intra_period=0, hash_sei=1)
    bad = bytearray(aus[0])
    bad[-5] ^= 0xFF
    _, errs = ffhevc.decode_stream([bytes(bad)], quiet=True)
    assert errs > 0


def test_rate_and_quality_move_with_qp():
    w, h = 128, 72
    frames = frames_of("camera", w, h, 2)
    sizes, psnrs = [], []
    for qp in (22, 32, 42):
        aus, recs = encode_all(frames, w, h, qp=qp, intra_period=0)
        sizes.append(sum(map(len, aus)))
        psnrs.append(synth.psnr(frames[1][:w * h], recs[1][:w * h]))
    assert sizes[0] > sizes[1] > sizes[2] and psnrs[0] > psnrs[1] > psnrs[2]


def test_stream_structure():
    w, h = 128, 72
    aus, _ = encode_all(frames_of("camera", w, h, 2), w, h, qp=32, intra_period=0)

    def nal_types(au):
        t, i = [], 0
        while True:
            i = au.find(b"\0\0\0\1", i)
            if i < 0:
                return t
            t.append((au[i + 4] >> 1) & 63)
            i += 4
    assert nal_types(aus[0]) == [32, 33, 34, 19]        # VPS SPS PPS IDR_W_RADL (filter.h:52-53)
    assert nal_types(aus[1]) == [1]                     # TRAIL_R
    assert all(au[4] >> 7 == 0 for au in aus)


# ---- primitives -----------------------------------------------------------------------------

def i16(a):
    return np.ascontiguousarray(a, dtype=np.int16)


def test_dct_matrix_matches_the_standard_rows(oracle_lib):
    assert [oracle_lib.orc_dct_coef(4, 1, n) for n in range(4)] == [83, 36, -36, -83]
    assert [oracle_lib.orc_dct_coef(8, 1, n) for n in range(8)] == [89, 75, 50, 18, -18, -50, -75, -89]
    assert [oracle_lib.orc_dct_coef(16, 1, n) for n in range(8)] == [90, 87, 80, 70, 57, 43, 25, 9]
    assert [oracle_lib.orc_dct_coef(32, 1, n) for n in range(16)] == [90, 90, 88, 85, 82, 78, 73, 67, 61, 54, 46, 38, 31, 22, 13, 4]
    assert [oracle_lib.orc_dct_coef(32, 31, n) for n in range(4)] == [4, -13, 22, -31]
    for N in (4, 8, 16, 32):
        assert all(oracle_lib.orc_dct_coef(N, 0, n) == 64 for n in range(N))
        M = np.array([[oracle_lib.orc_dct_coef(N, k, n) for n in range(N)] for k in range(N)], np.int64)
        G = M @ M.T                                        # near-orthogonal: diagonal ~ 64*64*N
        assert np.all(np.abs(np.diag(G) - 4096 * N) <= 4096 * N * 0.01)


@pytest.mark.parametrize("log2n", [2, 3, 4, 5])
def test_dct_known_answers_and_roundtrip(oracle_lib, log2n):
    n = 1 << log2n
    coef = np.zeros(n * n, np.int16)
    res = np.zeros(n * n, np.int16)
    dc = i16(np.full(n * n, 100))
    oracle_lib.orc_fdct(ptr(dc), ptr(coef), log2n)
    # DC gain of the two forward passes: 64*N >> (log2N-1), then *64*N >> (log2N+6)
    assert coef[0] == (((100 * 64 * n) >> (log2n - 1)) * 64 * n) >> (log2n + 6)
    assert not coef[1:].any()
    oracle_lib.orc_idct(ptr(coef), ptr(res), log2n)
    assert np.abs(res.astype(int) - 100).max() <= 1
    rng = np.random.default_rng(log2n)
    x = i16(rng.integers(-255, 256, n * n))
    oracle_lib.orc_fdct(ptr(x), ptr(coef), log2n)
    oracle_lib.orc_idct(ptr(coef), ptr(res), log2n)
    assert np.abs(res.astype(int) - x).max() <= 6            # integer DCT pair is near-lossless (not exactly orthogonal)
    imp = np.zeros(n * n, np.int16)
    imp[0] = 64                                              # inverse of a DC-only block is flat
    oracle_lib.orc_idct(ptr(imp), ptr(res), log2n)
    assert len(set(res.tolist())) == 1 and res[0] == (((64 * 64 + 64) >> 7) * 64 + 2048) >> 12


def test_dst_roundtrip(oracle_lib):
    rng = np.random.default_rng(7)
    x = i16(rng.integers(-255, 256, 16))
    c = np.zeros(16, np.int16)
    r = np.zeros(16, np.int16)
    oracle_lib.orc_fdst4(ptr(x), ptr(c))
    oracle_lib.orc_idst4(ptr(c), ptr(r))
    assert np.abs(r.astype(int) - x).max() <= 2


def test_quant_dequant_known_answers(oracle_lib):
    c = i16([1000, -1000, 10, 0] * 4)
    lv = np.zeros(16, np.int16)
    # qp 22: per 3, rem 4 -> scale 16384; log2n 2: qbits = 14+3+5 = 22; inter offset 85<<13
    nz = oracle_lib.orc_quant(ptr(c), ptr(lv), 2, 22, 0)
    exp = (1000 * 16384 + (85 << 13)) >> 22
    assert lv[0] == exp and lv[1] == -exp and lv[2] == 0 and nz == 8
    d = np.zeros(16, np.int16)
    oracle_lib.orc_dequant(ptr(lv), ptr(d), 2, 22)
    assert d[0] == ((exp * 16 * 64 << 3) + 16) >> 5 and d[1] == -d[0]
    assert [oracle_lib.orc_chroma_qp(q) for q in (29, 30, 35, 43, 44, 51)] == [29, 29, 33, 37, 38, 45]


def test_sad_satd_known_answers(oracle_lib):
    a = np.zeros(64, np.uint8)
    b = np.full(64, 3, np.uint8)
    assert oracle_lib.orc_sad(ptr(a), 8, ptr(b), 8, 8, 8) == 192
    # constant difference d on 8x8: only the DC Hadamard term = 64*d -> (64*3+2)>>2
    assert oracle_lib.orc_satd(ptr(a), 8, ptr(b), 8, 8, 8) == (64 * 3 + 2) >> 2
    assert oracle_lib.orc_satd(ptr(a), 8, ptr(b), 8, 4, 4) == (16 * 3 + 1) >> 1
    assert oracle_lib.orc_satd(ptr(a), 8, ptr(a), 8, 8, 8) == 0


def test_intra_known_answers(oracle_lib):
    n, log2n = 8, 3
    refs = np.full(4 * n + 1, 77, np.uint8)
    out = np.zeros(n * n, np.uint8)
    for mode in range(35):                                   # flat neighbours predict a flat block
        oracle_lib.orc_intra_predict(ptr(refs), log2n, mode, 0, ptr(out), n)
        assert (out == 77).all(), mode
    refs = np.arange(4 * n + 1, dtype=np.uint8) * 3
    oracle_lib.orc_intra_predict(ptr(refs), log2n, 26, 1, ptr(out), n)      # pure vertical, chroma: copy top row
    top = refs[2 * n + 1:2 * n + 1 + n]
    assert np.array_equal(out.reshape(n, n), np.tile(top, (n, 1)))
    oracle_lib.orc_intra_predict(ptr(refs), log2n, 10, 1, ptr(out), n)      # pure horizontal: copy left column
    left = refs[2 * n - 1::-1][:n]
    assert np.array_equal(out.reshape(n, n), np.tile(left[:, None], (1, n)))


def test_mc_known_answers(oracle_lib):
    w = h = 32
    ref = synth.noise(9, w * h)
    out = np.zeros(64, np.uint8)
    oracle_lib.orc_mc_luma(ptr(ref), w, w, h, 8, 8, 8, 8, 4 * 2, -4 * 3, ptr(out), 8)    # integer mv = copy
    assert np.array_equal(out.reshape(8, 8), ref.reshape(h, w)[5:13, 10:18])
    flat = np.full(w * h, 200, np.uint8)
    for mv in ((1, 0), (0, 2), (3, 3), (-5, 7)):                                           # filters have unit DC gain
        oracle_lib.orc_mc_luma(ptr(flat), w, w, h, 8, 8, 8, 8, mv[0], mv[1], ptr(out), 8)
        assert (out == 200).all()
        oracle_lib.orc_mc_chroma(ptr(flat), w, w, h, 8, 8, 8, 8, mv[0], mv[1], ptr(out), 8)
        assert (out == 200).all()
    # half-sample horizontal: (-1,4,-11,40,40,-11,4,-1) on a step edge
    row = np.array([0] * 16 + [255] * 16, np.uint8)
    img = np.tile(row, (h, 1)).ravel().copy()
    oracle_lib.orc_mc_luma(ptr(img), w, w, h, 12, 8, 8, 8, 2, 0, ptr(out), 8)
    taps = [-1, 4, -11, 40, 40, -11, 4, -1]
    exp = [min(255, max(0, (sum(t * int(row[12 + x + k - 3]) for k, t in enumerate(taps)) + 32) >> 6)) for x in range(8)]
    assert out[:8].tolist() == exp


def test_deblock_known_answers(oracle_lib):
    # flat step of 10 across a vertical edge at QP 37, bS 2 -> strong filter smooths, bounded by 2*tc
    img = np.zeros((4, 16), np.uint8)
    img[:, :8] = 100
    img[:, 8:] = 110
    buf = img.ravel().copy()
    oracle_lib.orc_deblock_luma_segment(C.c_void_p(buf.ctypes.data + 8), 1, 16, 2, 37)
    r = buf.reshape(4, 16)
    assert (r[:, 7] > 100).all() and (r[:, 8] < 110).all() and (r[:, :5] == 100).all() and (r[:, 11:] == 110).all()
    # a large step is a real edge (|delta| >= 10*tc): untouched
    img[:, :8] = 0
    img[:, 8:] = 255
    buf = img.ravel().copy()
    oracle_lib.orc_deblock_luma_segment(C.c_void_p(buf.ctypes.data + 8), 1, 16, 2, 37)
    assert np.array_equal(buf.reshape(4, 16), img)


def test_oracle_streams_match_the_committed_golden_hashes():
    """tests/golden/hevc_golden.json (generator: make_hevc_golden.py) pins the oracle's streams and
    reconstructions; every stream behind a hash was decoded bit-exactly by FFmpeg when generated."""
    import json
    from pathlib import Path
    from tests.golden.make_hevc_golden import CASES, digest, run_case
    gold = json.loads((Path(__file__).parent / "golden" / "hevc_golden.json").read_text())
    assert set(gold) == {c["name"] for c in CASES}
    for c in CASES:
        got = digest(*run_case(c))
        want = gold[c["name"]]
        assert want["ffmpeg_verified_when_generated"] is True
        assert got["au_bytes"] == want["au_bytes"], c["name"]
        assert got["au_sha256"] == want["au_sha256"] and got["recon_sha256"] == want["recon_sha256"], c["name"]
