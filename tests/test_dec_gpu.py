"""GPU parity: the CUDA HEVC decoder (through the libOpenHevc* C ABI, driven like OpenHEVCFilter)
must reproduce, bit for bit, the reconstruction of the encoder that produced the stream -- which
the oracle tests pin to FFmpeg's independent decoder.  Streams come from the CPU oracle encoder and
from the GPU encoder."""
import numpy as np
import pytest

from kvazzup_b200.encoder import GpuEncoder
from kvazzup_b200.openhevc import OpenHEVCFilter, split_nals
from oracle.encoder import OracleEncoder
from tests import ffhevc
from tests.test_oracle_hevc import frames_of

pytestmark = pytest.mark.gpu


def decode_all(aus):
    f = OpenHEVCFilter()
    assert f.init()
    out = []
    for i, au in enumerate(aus):
        for nal in split_nals(au):
            got = f.process(nal, pts=i)
            if got is not None:
                out.append(got)
    f.close()
    return out


@pytest.mark.parametrize("kind,w,h,n,qp,kw", [
    ("camera", 64, 64, 2, 32, {}),
    ("camera", 192, 136, 4, 32, {}),
    ("camera", 72, 200, 3, 27, {"deblock": 0}),
    ("noise", 128, 72, 3, 0, {}),
    ("noise", 128, 72, 3, 10, {}),
    ("noise", 128, 72, 3, 45, {}),
    ("camera", 128, 72, 3, 51, {}),
    ("camera", 416, 240, 5, 27, {}),
    ("screen", 416, 240, 5, 32, {}),
    ("camera", 200, 200, 7, 30, {"intra_period": 3}),
    ("camera", 256, 128, 4, 22, {"search_range": 16}),
    ("camera", 64, 8, 3, 37, {}),
    ("camera", 640, 480, 3, 32, {}),
    ("camera", 192, 136, 4, 32, {"sao": 1}),
    ("noise", 128, 72, 3, 10, {"sao": 1}),
    ("screen", 416, 240, 5, 32, {"sao": 1}),
    ("camera", 200, 200, 7, 30, {"sao": 1, "intra_period": 3}),
    ("camera", 72, 200, 3, 27, {"sao": 1, "deblock": 0}),
    ("camera", 640, 480, 3, 22, {"sao": 1, "qp_delta": 1}),
    ("camera", 416, 240, 6, 37, {"sao": 2, "intra_period": 4}),          # sao_merge_left / _up
    ("screen", 640, 256, 5, 40, {"sao": 2}),
    ("sports", 416, 240, 6, 32, {"intra_in_p": 1}),                       # intra CUs in P pictures
    ("sports", 200, 136, 5, 22, {"intra_in_p": 1}),
    ("noise", 128, 72, 3, 30, {"intra_in_p": 1}),
    ("sports", 640, 480, 5, 35, {"intra_in_p": 1, "sao": 2, "intra_period": 4, "search_range": 12, "qp_delta": 1}),
    ("sports", 416, 240, 6, 32, {"me_coarse": 16, "search_range": 4}),   # vectors of +-70 samples, also beyond the picture edge
    ("sports", 640, 256, 4, 30, {"me_coarse": 32, "search_range": 6, "intra_in_p": 1}),
    # several reference pictures and temporal MV prediction (what a Kvazaar peer sends)
    ("sports", 416, 240, 7, 32, {"tmvp": 1}),
    ("sports", 416, 240, 7, 32, {"refs": 2}),
    ("camera", 192, 136, 6, 30, {"refs": 3, "tmvp": 1}),
    ("noise", 128, 72, 5, 32, {"refs": 4, "tmvp": 1}),
    ("sports", 416, 240, 8, 32, {"refs": 4, "tmvp": 1, "me_coarse": 16, "search_range": 4, "sao": 2, "intra_in_p": 1, "intra_period": 5}),
    ("sports", 640, 256, 5, 27, {"refs": 3, "tmvp": 1, "qp_delta": 1}),
    ("camera", 1280, 720, 4, 32, {"refs": 3, "tmvp": 1, "sao": 2, "intra_in_p": 1, "me_coarse": 16, "search_range": 6}),
    # transform tree (split transform units, TU-level intra prediction, TU-edge deblocking), cabac_init_flag,
    # one substream per picture (neither WPP nor tiles)
    ("camera", 192, 136, 4, 30, {"tr_depth": 1}),
    ("sports", 416, 240, 5, 30, {"tr_depth": 2}),
    ("screen", 416, 240, 4, 30, {"tr_depth": 2}),
    ("noise", 128, 72, 3, 22, {"tr_depth": 1, "cabac_init": 1}),
    ("camera", 416, 240, 5, 32, {"tr_depth": 2, "cabac_init": 1, "sao": 2, "intra_in_p": 1, "refs": 2, "tmvp": 1, "qp_delta": 1, "intra_period": 3}),
    ("camera", 416, 240, 4, 32, {"no_wpp": 1, "tr_depth": 1, "sao": 1}),
    ("camera", 1920, 1080, 3, 27, {"tr_depth": 2, "refs": 2, "tmvp": 1, "sao": 2, "me_coarse": 16, "search_range": 6}),
    # syntax of a Kvazaar-family peer (oracle/hevc_enc.h: tu4, intra_sizes, chroma_modes, sign_hiding, ...)
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
    ("camera", 640, 480, 3, 24, {"cb_qp_offset": -5, "cr_qp_offset": 6, "qp_delta": 1, "sao": 2, "beta_offset_div2": -2, "tc_offset_div2": 3}),
    ("sports", 416, 240, 5, 30, {"tr_depth": 2, "tu4": 1, "intra_sizes": 7, "chroma_modes": 1, "sign_hiding": 1, "strong_intra": 1,
                                 "cb_qp_offset": 2, "cr_qp_offset": -2, "beta_offset_div2": 1, "tc_offset_div2": 1, "sao": 2,
                                 "intra_in_p": 1, "refs": 2, "tmvp": 1, "qp_delta": 1, "intra_period": 3, "cabac_init": 1}),
    ("camera", 1920, 1080, 3, 27, {"tr_depth": 3, "tu4": 1, "intra_sizes": 7, "chroma_modes": 1, "sign_hiding": 1, "refs": 2, "tmvp": 1,
                                   "sao": 2, "me_coarse": 16, "search_range": 6, "intra_in_p": 1}),
    # scaling lists: the default ones (Kvazaar --scaling-list default), lists carried in the SPS / in the PPS
    # (coded, copied, inferred default; DC coefficients of 16x16 / 32x32 blocks)
    ("camera", 416, 240, 4, 22, {"scaling_list": 1, "intra_period": 3}),
    ("noise", 256, 136, 3, 12, {"scaling_list": 1, "tr_depth": 2, "tu4": 1, "intra_sizes": 7}),
    ("sports", 640, 480, 4, 30, {"scaling_list": 1, "sao": 2, "intra_in_p": 1, "sign_hiding": 1, "tr_depth": 1, "me_coarse": 16, "search_range": 4}),
    ("camera", 416, 240, 4, 27, {"scaling_list": 2, "intra_period": 3, "intra_sizes": 3}),
    ("noise", 256, 136, 3, 17, {"scaling_list": 2, "tr_depth": 2, "tu4": 1, "intra_sizes": 7, "qp_delta": 1}),
    ("screen", 640, 200, 4, 32, {"scaling_list": 3, "tr_depth": 1, "intra_period": 2}),
    ("noise", 512, 136, 3, 22, {"scaling_list": 3, "tr_depth": 2, "tu4": 1, "intra_sizes": 7, "cb_qp_offset": 3, "cr_qp_offset": -4}),
    ("camera", 1920, 1080, 2, 27, {"scaling_list": 2, "tr_depth": 2, "intra_sizes": 3, "sao": 2, "me_coarse": 16, "search_range": 6}),
])
def test_decoder_reproduces_oracle_reconstruction(kind, w, h, n, qp, kw):
    frames = frames_of(kind, w, h, n)
    enc = OracleEncoder(w, h, qp=qp, **({"intra_period": 0} | kw))
    if kw.get("qp_delta"):
        from tests.test_oracle_hevc import roi_pattern
        enc.set_ctu_dqp(roi_pattern(w, h, 1, "random"))
    aus, recs = [], []
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
    dec = decode_all(aus)
    assert len(dec) == n
    for i, (pic, pw, ph) in enumerate(dec):
        assert (pw, ph) == (w, h)
        bad = np.flatnonzero(pic != recs[i])
        assert bad.size == 0, f"picture {i}: {bad.size} samples differ, first at {bad[:6]}"


def test_decoder_on_gpu_encoder_stream_1080p_and_ffmpeg_agreement():
    w, h, n = 1920, 1080, 3
    frames = frames_of("camera", w, h, n)
    g = GpuEncoder(w, h, qp=27, intra_period=0, search_range=12)
    aus, recs = [], []
    for f in frames:
        aus.append(g.encode(f))
        recs.append(g.recon())
    dec = decode_all(aus)
    assert len(dec) == n
    for i in range(n):
        assert np.array_equal(dec[i][0], recs[i]), i
    if ffhevc.required():
        ff, errs = ffhevc.decode_stream(aus)
        assert errs == 0
        for i in range(n):
            assert np.array_equal(dec[i][0], ff[i][0]), i


@pytest.mark.parametrize("threads,mode", [(1, "Frame"), (4, "Frame"), (9, "Both"), (6, "Slice")])
def test_frame_threading_delays_output_and_changes_nothing_else(threads, mode):
    """OpenHEVC frame threads keep `threads` pictures in flight and delay output by threads - 1;
    ours are concurrent parses on the GPU.  Same pictures, same order, IDR in the middle."""
    w, h, n = 416, 240, 12
    enc = OracleEncoder(w, h, qp=30, intra_period=5)
    aus, recs = [], []
    for f in frames_of("camera", w, h, n):
        aus.append(enc.encode(f))
        recs.append(enc.recon())
    delay = threads - 1 if mode != "Slice" else 0
    f = OpenHEVCFilter(threads, mode)
    assert f.init()
    out, counts = [], []
    for i, au in enumerate(aus):
        k = 0
        for nal in split_nals(au):
            got = f.process(nal, pts=1000 + i)
            if got is not None:
                out.append(got)
                k += 1
        counts.append(k)
    assert counts == [0] * min(delay, n) + [1] * max(n - delay, 0)
    out += f.drain()
    assert f.drain() == []
    f.close()
    assert len(out) == n
    for i in range(n):
        assert np.array_equal(out[i][0], recs[i]), i


def test_filter_gates_on_parameter_sets_and_reports_picture_info():
    w, h = 192, 136
    enc = OracleEncoder(w, h, qp=30, intra_period=0)
    aus = [enc.encode(f) for f in frames_of("camera", w, h, 2)]
    f = OpenHEVCFilter()
    assert f.init() and "b200" in f.version()
    nals = split_nals(aus[0])
    assert [n[4] >> 1 for n in nals] == [32, 33, 34, 19]
    assert f.process(nals[3]) is None and f.discarded == 1      # VCL before VPS/SPS/PPS is discarded (:116-182)
    for nal in nals[:3]:
        assert f.process(nal) is None
    pic = f.process(nals[3])
    assert pic is not None and pic[1:] == (w, h)
    # the reference copies the frame rate into vInfo (openhevcfilter.cpp:232-233) and DisplayFilter
    # divides by it (displayfilter.cpp:153): never 0/0, the VUI timing when the stream has one
    assert f.frame_rate() == (30, 1)
    f.close()
    enc = OracleEncoder(w, h, qp=30, intra_period=0, fps_num=30000, fps_den=1001)
    f = OpenHEVCFilter()
    assert f.init()
    for nal in split_nals(enc.encode(frames_of("camera", w, h, 1)[0])):
        pic = f.process(nal)
    assert pic is not None and f.frame_rate() == (30000, 1001)
    f.close()


def test_decoder_rejects_what_it_cannot_decode():
    from kvazzup_b200.capi import B200Error
    w, h = 64, 64
    enc = OracleEncoder(w, h, qp=30, intra_period=0)
    au = enc.encode(frames_of("camera", w, h, 1)[0])
    nals = split_nals(au)
    f = OpenHEVCFilter()
    assert f.init()
    sps = bytearray(nals[1])
    sps[-3] ^= 0x10                      # flip a tool flag near the end of the SPS (SAO / PCM / ... region)
    f.process(nals[0])
    with pytest.raises(B200Error):
        f.process(bytes(sps))
        f.process(nals[2])
        f.process(nals[3])
    f.close()
    # corrupt slice data: must fail or at least never crash; a wrong picture must not be reported as an error-free decode
    g = OpenHEVCFilter()
    assert g.init()
    for nal in nals[:3]:
        g.process(nal)
    bad = bytearray(nals[3])
    for k in range(40, min(len(bad), 200), 7):
        bad[k] ^= 0x5A
    try:
        g.process(bytes(bad))
    except B200Error:
        pass
    g.close()


def test_rtp_loopback_encoder_packets_decoder():
    """Encoder -> RFC 7798 packets (uvgRTP's role) -> one NAL per buffer -> decoder, as in a call;
    a lost slice fragment costs that picture and drifts until the next IDR, which is exact again."""
    from kvazzup_b200 import rtp
    w, h, n = 416, 240, 8
    frames = frames_of("camera", w, h, n)
    g = GpuEncoder(w, h, qp=30, intra_period=4, search_range=8)
    snd, rcv = rtp.RtpSender(5, 96, 300), rtp.RtpReceiver(5)
    f = OpenHEVCFilter()
    assert f.init()
    shown = {}
    for i, fr in enumerate(frames):
        au = g.encode(fr)
        rec = g.recon()
        pkts = snd.push_frame(au, 3000 * i)
        assert len(pkts) > 4
        if i == 1:
            pkts = pkts[:-1]                     # lose the last fragment of picture 1's slice
        for p in pkts:
            for nal, ts, marker in rcv.receive(p):
                assert ts == 3000 * i
                assert rtp.is_hevc_intra(nal) == (nal[4] >> 1 == 19)
                got = f.process(nal, pts=i)
                if got is not None:
                    # pictures 2 and 3 predict from the lost picture 1: like OpenHEVC the decoder
                    # conceals with the last picture it has, so they drift until the next IDR
                    assert np.array_equal(got[0], rec) == (i not in (2, 3)), i
                    shown[i] = True
    assert f.missing_refs() == 1               # picture 2 named POC 1, which never arrived: the application can ask for an IDR
    f.close()
    assert rcv.lost == 1
    assert sorted(shown) == [0, 2, 3, 4, 5, 6, 7]


def test_device_resident_decode_to_rgb32_equals_the_host_chain():
    """decoder output stays on the GPU and feeds the display conversion directly."""
    import torch
    from kvazzup_b200 import convert
    w, h, n = 416, 240, 3
    g = GpuEncoder(w, h, qp=30, intra_period=0)
    aus = [g.encode(f) for f in frames_of("camera", w, h, n)]
    host = [convert.yuv420_to_rgb32(p[0], w, h) for p in decode_all(aus)]
    f = OpenHEVCFilter()
    assert f.init()
    f.set_host_output(False)
    d_rgb = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
    k = 0
    for au in aus:
        for nal in split_nals(au):
            d_pic = f.process_dev(nal)
            if d_pic:
                convert.i420_to_rgb32_dev(d_pic, d_rgb.data_ptr(), w, h, 1, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                assert np.array_equal(d_rgb.cpu().numpy(), host[k]), k
                k += 1
    assert k == n
    f.close()


@pytest.mark.parametrize("kind,w,h,n,qp,tiles,wpp,threads,kw", [
    ("camera", 416, 240, 5, 30, 2, 0, 1, {}),
    ("camera", 416, 240, 6, 27, 3, 1, 1, {"intra_period": 4}),
    ("noise", 512, 136, 3, 20, 4, 0, 1, {}),
    ("screen", 640, 200, 5, 35, 2, 1, 1, {"deblock": 0}),
    ("camera", 640, 256, 9, 30, 3, 1, 4, {"intra_period": 4}),          # tiles + decoder frame threading
    ("camera", 1920, 1080, 3, 32, 4, 1, 1, {"search_range": 12}),
    ("camera", 1920, 1080, 3, 32, 4, 0, 1, {"search_range": 12}),
    ("camera", 416, 240, 5, 30, 2, 0, 1, {"tile_rows": 2}),                       # tile grids
    ("camera", 640, 256, 6, 27, 3, 1, 1, {"tile_rows": 2, "intra_period": 4}),
    ("sports", 640, 480, 5, 32, 2, 0, 4, {"tile_rows": 3, "me_coarse": 16, "search_range": 4, "sao": 2, "intra_in_p": 1}),
    ("camera", 1920, 1080, 3, 32, 2, 1, 1, {"tile_rows": 2, "search_range": 12}),
])
def test_decoder_reads_tile_columns_as_strips(kind, w, h, n, qp, tiles, wpp, threads, kw):
    """Tiled streams (tile columns, no loop filter across tiles, motion inside the tile; with or
    without WPP inside the tiles) decode strip by strip to the encoder's reconstruction; for the
    WPP-less mode that reconstruction is also what FFmpeg produces (tests/test_enc_gpu.py)."""
    from kvazzup_b200.encoder import GpuTiledEncoder
    frames = frames_of(kind, w, h, n)
    g = GpuTiledEncoder(w, h, tiles, qp=qp, wpp=wpp, **({"intra_period": 0} | kw))
    aus, recs = [], []
    for f in frames:
        aus.append(g.encode(f))
        recs.append(g.recon())
    g.close()
    f = OpenHEVCFilter(threads, "Frame" if threads > 1 else "Slice")
    assert f.init()
    out = []
    for i, au in enumerate(aus):
        for nal in split_nals(au):
            got = f.process(nal, pts=i)
            if got is not None:
                out.append(got)
    out += f.drain()
    f.close()
    assert len(out) == n
    for i, (pic, pw, ph) in enumerate(out):
        assert (pw, ph) == (w, h)
        bad = np.flatnonzero(pic != recs[i])
        assert bad.size == 0, f"picture {i}: {bad.size} samples differ, first at {bad[:6]}"


@pytest.mark.parametrize("kind,w,h,n,qp,tiles,kw", [
    ("camera", 410, 234, 3, 30, 1, {}),
    ("screen", 638, 200, 3, 32, 1, {"sao": 2, "tr_depth": 1}),
    ("camera", 1366, 768, 2, 32, 1, {"search_range": 6, "refs": 2, "tmvp": 1}),
    ("camera", 634, 250, 3, 30, 2, {"tile_rows": 2}),                      # window over an assembled tile grid
])
def test_decoder_crops_to_the_conformance_window(kind, w, h, n, qp, tiles, kw):
    """Streams whose source size is not a multiple of 8 (a Kvazaar peer pads and signals a conformance window):
    libOpenHevcGetPictureInfo reports the window, the output is the cropped reconstruction (host and device)."""
    import torch
    from kvazzup_b200 import convert
    from tests.test_oracle_hevc import odd_size_frames, pad_i420, crop_i420
    from oracle.encoder import OracleTiledEncoder
    W, H = (w + 7) & ~7, (h + 7) & ~7
    frames = [pad_i420(f, w, h, W, H) for f in odd_size_frames(kind, w, h, n)]
    args = {"intra_period": 0, "conf_right": W - w, "conf_bottom": H - h} | kw
    if tiles > 1:
        rows = args.pop("tile_rows", 1)
        enc = OracleTiledEncoder(W, H, tiles, qp=qp, tile_rows=rows, **args)
    else:
        enc = OracleEncoder(W, H, qp=qp, **args)
    aus, recs = [], []
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(crop_i420(enc.recon(), W, H, w, h))
    dec = decode_all(aus)
    assert len(dec) == n
    for i, (pic, pw, ph) in enumerate(dec):
        assert (pw, ph) == (w, h)
        assert np.array_equal(pic, recs[i]), f"picture {i}"
    if ffhevc.required():
        ff, errs = ffhevc.decode_stream(aus)
        assert errs == 0 and np.array_equal(ff[-1][0], recs[-1])
    # device-resident hand-over: the window as a packed picture
    f = OpenHEVCFilter()
    assert f.init()
    f.set_host_output(False)
    k = 0
    for au in aus:
        for nal in split_nals(au):
            d_pic = f.process_dev(nal)
            if d_pic:
                d_rgb = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
                convert.i420_to_rgb32_dev(d_pic, d_rgb.data_ptr(), w, h, 1, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                assert np.array_equal(d_rgb.cpu().numpy(), convert.yuv420_to_rgb32(recs[k], w, h)), k
                k += 1
    assert k == n
    f.close()


def test_decoder_switches_between_tiled_and_untiled_streams():
    from kvazzup_b200.encoder import GpuTiledEncoder
    w, h = 416, 240
    frames = frames_of("camera", w, h, 3)
    f = OpenHEVCFilter()
    assert f.init()
    for tiles in (1, 3, 1, 2):
        if tiles == 1:
            e = GpuEncoder(w, h, qp=30, intra_period=0)
        else:
            e = GpuTiledEncoder(w, h, tiles, qp=30, intra_period=0, wpp=1)
        for fr in frames:
            au = e.encode(fr)
            rec = e.recon()
            pics = [p for nal in split_nals(au) if (p := f.process(nal)) is not None]
            assert len(pics) == 1 and np.array_equal(pics[0][0], rec), tiles
        e.close()
    f.close()


def test_decoder_survives_corrupted_peer_streams():
    """Network input: access units of a stream that uses every syntax element the parser knows (transform
    trees down to 4x4, NxN, chroma modes, sign hiding, SAO, several references, temporal candidates,
    per-CTU QP) with bytes flipped at random places.  Every variant must either decode (to whatever) or be
    refused with an error -- and the decoder must then decode a clean stream exactly, i.e. nothing a
    corrupt picture does may damage its state.  (tools/run_sanitizer.sh runs this under memcheck.)"""
    from kvazzup_b200.capi import B200Error
    w, h, n = 192, 136, 4
    kw = {"tr_depth": 2, "tu4": 1, "intra_sizes": 7, "chroma_modes": 1, "sign_hiding": 1, "sao": 2, "intra_in_p": 1,
          "refs": 2, "tmvp": 1, "qp_delta": 1, "intra_period": 0, "cabac_init": 1}
    frames = frames_of("sports", w, h, n)
    enc = OracleEncoder(w, h, qp=30, **kw)
    aus, recs = [], []
    for f in frames:
        aus.append(enc.encode(f))
        recs.append(enc.recon())
    enc.close()
    rng = np.random.default_rng(7)
    refused = decoded = 0
    for trial in range(40):
        dec = OpenHEVCFilter()
        assert dec.init()
        victim = int(rng.integers(0, n))
        for i, au in enumerate(aus):
            for nal in split_nals(au):
                data = bytearray(nal)
                nal_type = (data[4] >> 1) & 63
                if i == victim and nal_type < 32 and len(data) > 24:            # a slice NAL: corrupt its payload
                    for _ in range(int(rng.integers(1, 6))):
                        k = int(rng.integers(8, len(data)))
                        data[k] ^= int(rng.integers(1, 256))
                try:
                    if dec.process(bytes(data), pts=i) is not None:
                        decoded += 1
                except B200Error:
                    refused += 1
        dec.close()
    assert decoded > 0 and refused >= 0
    # and a fresh, clean pass still decodes bit-exactly
    out = decode_all(aus)
    assert len(out) == n and all(np.array_equal(out[i][0], recs[i]) for i in range(n))
