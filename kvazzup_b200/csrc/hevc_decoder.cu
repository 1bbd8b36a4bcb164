// Host side of the B200 HEVC decoder and the libOpenHevc* boundary (include/b200_openhevc.h):
// NAL splitting, parameter-set and slice-header parsing, substream extraction, kernel sequencing
// (CABAC parse -> reconstruction -> deblocking), picture output.  Replaces what
// reference src/media/processing/openhevcfilter.cpp:38-56,145,195-199 calls in OpenHEVC.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <vector>

#include "../../include/b200_openhevc.h"
#include "hevc_kernels.h"
#include "runtime.h"

namespace b200 {

namespace {

struct BitReader {
  const uint8_t *p;
  size_t n, pos = 0;       // pos in bits
  bool bad = false;
  BitReader(const uint8_t *d, size_t len) : p(d), n(len) {}
  uint32_t u(int bits)
  {
    uint32_t v = 0;
    for (int i = 0; i < bits; i++) {
      if (pos >= n * 8) { bad = true; return 0; }
      v = (v << 1) | ((p[pos >> 3] >> (7 - (pos & 7))) & 1);
      pos++;
    }
    return v;
  }
  uint32_t ue()
  {
    int z = 0;
    while (!bad && u(1) == 0 && z < 32) z++;
    if (z >= 32) { bad = true; return 0; }
    return z ? ((1u << z) - 1 + u(z)) : 0;
  }
  int32_t se()
  {
    uint32_t k = ue();
    return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
  }
  void align() { pos = (pos + 7) & ~(size_t)7; }
};

std::vector<uint8_t> unescape(const uint8_t *p, size_t n)
{
  std::vector<uint8_t> out;
  out.reserve(n);
  int zeros = 0;
  for (size_t i = 0; i < n; i++) {
    if (zeros >= 2 && p[i] == 3) { zeros = 0; continue; }
    out.push_back(p[i]);
    zeros = p[i] == 0 ? zeros + 1 : 0;
  }
  return out;
}

struct Sps {
  bool valid = false;
  int width = 0, height = 0, log2_max_poc = 8, num_rps = 0;
};
struct Pps {
  bool valid = false;
  int init_qp = 26, deblock_disabled = 0, loop_across_slices = 0, deblock_ctrl = 0, qp_delta = 0;
};

const uint8_t kChromaQpD[58] = {
  0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
  29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};

}  // namespace

// One picture in flight: everything the CABAC parse of a picture reads and writes.  Parsing needs
// nothing from other pictures (no TMVP), so the parses of up to `frame_delay + 1` pictures run
// concurrently on their own streams; only reconstruction is chained from picture to picture.
struct DecSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_parsed = nullptr;
  uint8_t *d_data = nullptr, *d_small = nullptr, *d_ctu_qp = nullptr;
  CuInfo *d_cu = nullptr;
  int16_t *d_levels = nullptr;
  uint8_t *h_data = nullptr, *h_out = nullptr;
  uint32_t *h_bases = nullptr;
  int *h_status = nullptr;
  FrameParams fp{};
  int64_t pts = 0;
};

struct Decoder {
  Sps sps;
  Pps pps;
  bool vps_seen = false, started = false;
  FrameParams fp{};
  size_t frame_bytes = 0;
  cudaStream_t stream = nullptr;          // reconstruction chain
  uint8_t *d_rec[2] = {nullptr, nullptr};
  int *d_order = nullptr;
  std::vector<DecSlot> slots;
  std::deque<int> pending;                // slots whose parse was launched, oldest first
  size_t data_cap = 0, small_bytes = 0, off_flag = 0, off_prog = 0, off_ticket = 0, off_status = 0, off_bases = 0, off_ctx = 0;
  int frame_delay = 0;                    // pictures held back (OpenHEVC frame threads - 1)
  int next_slot = 0, out_slot = -1;
  bool host_output = true;                // false: pictures stay on the GPU (b200_dec_output_dev)
  const uint8_t *d_out = nullptr;         // device copy of the last output picture (the new reference)
  int cur = 0, have_ref = 0, pictures = 0;
  int64_t out_pts = 0;
  int fr_num = 0, fr_den = 0;

  ~Decoder() { release(); }
  void release()
  {
    if (stream) cudaStreamSynchronize(stream);
    for (DecSlot &s : slots) {
      if (s.stream) cudaStreamSynchronize(s.stream);
      if (s.d_data) cudaFree(s.d_data);
      if (s.d_small) cudaFree(s.d_small);
      if (s.d_ctu_qp) cudaFree(s.d_ctu_qp);
      if (s.d_cu) cudaFree(s.d_cu);
      if (s.d_levels) cudaFree(s.d_levels);
      if (s.h_data) cudaFreeHost(s.h_data);
      if (s.h_out) cudaFreeHost(s.h_out);
      if (s.h_bases) cudaFreeHost(s.h_bases);
      if (s.h_status) cudaFreeHost(s.h_status);
      if (s.ev_parsed) cudaEventDestroy(s.ev_parsed);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
    slots.clear();
    pending.clear();
    for (int i = 0; i < 2; i++) if (d_rec[i]) cudaFree(d_rec[i]);
    if (d_order) cudaFree(d_order);
    if (stream) cudaStreamDestroy(stream);
    d_rec[0] = d_rec[1] = nullptr; d_order = nullptr; stream = nullptr;
    next_slot = 0; out_slot = -1;
  }

  bool alloc(int w, int h)
  {
    release();
    fp.w = w; fp.h = h; fp.w8 = w / 8; fp.h8 = h / 8;
    fp.ctb_cols = (w + kCtb - 1) / kCtb; fp.ctb_rows = (h + kCtb - 1) / kCtb;
    fp.deblock = 1; fp.search_range = 8; fp.lambda_q4 = 0; fp.ctu_qp = nullptr; fp.ctu_delta = nullptr; fp.ctu_first = nullptr;
    frame_bytes = (size_t)w * h * 3 / 2;
    data_cap = frame_bytes * 2 + 65536;
    const int rows = fp.ctb_rows;
    off_flag = 0; off_prog = sizeof(int) * rows; off_ticket = 2 * sizeof(int) * rows;
    off_status = off_ticket + 2 * sizeof(int); off_bases = off_status + 2 * sizeof(int);
    off_ctx = off_bases + sizeof(uint32_t) * (rows + 1);
    small_bytes = off_ctx + (size_t)rows * CTX_COUNT;
    if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    if (!cuda_ok(cudaMalloc((void **)&d_rec[0], frame_bytes), "cudaMalloc")) return false;
    if (!cuda_ok(cudaMalloc((void **)&d_rec[1], frame_bytes), "cudaMalloc")) return false;
    {
      std::vector<int> order((size_t)fp.ctb_cols * fp.ctb_rows);
      intra_wavefront_order(fp.ctb_cols, fp.ctb_rows, order.data());
      if (!cuda_ok(cudaMalloc((void **)&d_order, order.size() * sizeof(int)), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMemcpy(d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice), "H2D order")) return false;
    }
    slots.resize(frame_delay + 1);
    for (DecSlot &s : slots) {
      if (!cuda_ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
      if (!cuda_ok(cudaEventCreateWithFlags(&s.ev_parsed, cudaEventDisableTiming), "cudaEventCreate")) return false;
      if (!cuda_ok(cudaMalloc((void **)&s.d_data, data_cap), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMalloc((void **)&s.d_small, small_bytes), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMalloc((void **)&s.d_ctu_qp, (size_t)fp.ctb_cols * fp.ctb_rows), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMalloc((void **)&s.d_cu, sizeof(CuInfo) * fp.w8 * fp.h8), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMalloc((void **)&s.d_levels, frame_bytes * sizeof(int16_t)), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMallocHost((void **)&s.h_out, frame_bytes), "cudaMallocHost")) return false;
      if (!cuda_ok(cudaMallocHost((void **)&s.h_data, data_cap), "cudaMallocHost")) return false;
      if (!cuda_ok(cudaMallocHost((void **)&s.h_bases, sizeof(uint32_t) * (rows + 1)), "cudaMallocHost")) return false;
      if (!cuda_ok(cudaMallocHost((void **)&s.h_status, sizeof(int) * 2), "cudaMallocHost")) return false;
    }
    cur = 0; have_ref = 0;
    return true;
  }

  bool parse_sps(const std::vector<uint8_t> &rbsp)
  {
    BitReader b(rbsp.data(), rbsp.size());
    b.u(4);
    int max_sub = (int)b.u(3);
    b.u(1);
    if (max_sub != 0) { set_error("decoder: SPS with sub-layers is not supported"); return false; }
    b.u(2); b.u(1);
    int profile = (int)b.u(5);
    b.u(32); b.u(4); b.u(32); b.u(11); b.u(1); b.u(8);
    (void)profile;
    b.ue();
    if (b.ue() != 1) { set_error("decoder: only 4:2:0 is supported"); return false; }
    int w = (int)b.ue(), h = (int)b.ue();
    if (b.u(1)) {
      uint32_t l = b.ue(), r = b.ue(), t = b.ue(), bo = b.ue();
      if (l | r | t | bo) { set_error("decoder: conformance window cropping is not supported"); return false; }
    }
    if (b.ue() != 0 || b.ue() != 0) { set_error("decoder: only 8-bit video is supported"); return false; }
    int log2_max_poc = (int)b.ue() + 4;
    b.u(1);
    b.ue(); b.ue(); b.ue();
    uint32_t min_cb = b.ue(), diff_cb = b.ue(), min_tb = b.ue(), diff_tb = b.ue(), depth_inter = b.ue(), depth_intra = b.ue();
    if (min_cb != 0 || diff_cb != 3) { set_error("decoder: coding block sizes other than 8..64 are not supported"); return false; }
    if (min_tb != 0 || diff_tb != 3) { set_error("decoder: transform block sizes other than 4..32 are not supported"); return false; }
    if (depth_inter != 0 || depth_intra != 0) { set_error("decoder: transform hierarchy depth > 0 is not supported"); return false; }
    if (b.u(1)) { set_error("decoder: scaling lists are not supported"); return false; }
    if (b.u(1)) { set_error("decoder: AMP is not supported"); return false; }
    if (b.u(1)) { set_error("decoder: SAO is not supported"); return false; }
    if (b.u(1)) { set_error("decoder: PCM is not supported"); return false; }
    int num_rps = (int)b.ue();
    if (num_rps > 1) { set_error("decoder: more than one short-term RPS in the SPS is not supported"); return false; }
    if (num_rps == 1) {
      uint32_t neg = b.ue(), pos = b.ue();
      if (neg != 1 || pos != 0 || b.ue() != 0 || b.u(1) != 1) { set_error("decoder: only the previous picture as reference is supported"); return false; }
    }
    if (b.u(1)) { set_error("decoder: long-term reference pictures are not supported"); return false; }
    if (b.u(1)) { set_error("decoder: temporal MVP is not supported"); return false; }
    b.u(1);      // strong_intra_smoothing: irrelevant, no 32x32 intra blocks are accepted
    if (b.bad || w <= 0 || h <= 0 || (w & 7) || (h & 7)) { set_error("decoder: malformed SPS"); return false; }
    if (!sps.valid || sps.width != w || sps.height != h) {
      if (!alloc(w, h)) return false;
    }
    sps.valid = true; sps.width = w; sps.height = h; sps.log2_max_poc = log2_max_poc; sps.num_rps = num_rps;
    return true;
  }

  bool parse_pps(const std::vector<uint8_t> &rbsp)
  {
    BitReader b(rbsp.data(), rbsp.size());
    b.ue(); b.ue();
    if (b.u(1)) { set_error("decoder: dependent slice segments are not supported"); return false; }
    if (b.u(1)) { set_error("decoder: output_flag_present is not supported"); return false; }
    if (b.u(3)) { set_error("decoder: extra slice header bits are not supported"); return false; }
    if (b.u(1)) { set_error("decoder: sign data hiding is not supported"); return false; }
    if (b.u(1)) { set_error("decoder: cabac_init_present is not supported"); return false; }
    if (b.ue() != 0) { set_error("decoder: more than one reference picture is not supported"); return false; }
    b.ue();
    int init_qp = 26 + b.se();
    if (b.u(1)) { set_error("decoder: constrained intra prediction is not supported"); return false; }
    if (b.u(1)) { set_error("decoder: transform skip is not supported"); return false; }
    const int qp_delta = (int)b.u(1);
    if (qp_delta && b.ue() != 0) { set_error("decoder: quantisation groups smaller than the CTU are not supported"); return false; }
    if (b.se() != 0 || b.se() != 0) { set_error("decoder: chroma QP offsets are not supported"); return false; }
    if (b.u(1)) { set_error("decoder: slice chroma QP offsets are not supported"); return false; }
    if (b.u(1) || b.u(1)) { set_error("decoder: weighted prediction is not supported"); return false; }
    if (b.u(1)) { set_error("decoder: transquant bypass is not supported"); return false; }
    if (b.u(1)) { set_error("decoder: tiles are not supported"); return false; }
    if (!b.u(1)) { set_error("decoder: streams without WPP entry points are not supported"); return false; }
    pps.loop_across_slices = (int)b.u(1);
    pps.deblock_ctrl = (int)b.u(1);
    pps.deblock_disabled = 0;
    if (pps.deblock_ctrl) {
      if (b.u(1)) { set_error("decoder: deblocking override is not supported"); return false; }
      pps.deblock_disabled = (int)b.u(1);
      if (!pps.deblock_disabled && (b.se() != 0 || b.se() != 0)) { set_error("decoder: deblocking offsets are not supported"); return false; }
    }
    if (b.u(1)) { set_error("decoder: scaling lists are not supported"); return false; }
    if (b.u(1)) { set_error("decoder: reference list modification is not supported"); return false; }
    if (b.ue() != 0) { set_error("decoder: parallel merge level > 2 is not supported"); return false; }
    if (b.u(1)) { set_error("decoder: slice header extensions are not supported"); return false; }
    if (b.bad) { set_error("decoder: malformed PPS"); return false; }
    pps.valid = true; pps.init_qp = init_qp; pps.qp_delta = qp_delta;
    return true;
  }

  // returns 1 picture decoded, 0 nothing, -1 error
  int decode_slice(int nal_type, const uint8_t *payload, size_t n, int64_t pts)
  {
    if (!sps.valid || !pps.valid) { set_error("decoder: slice before parameter sets"); return -1; }
    // The slice header sits in the first bytes; unescape the whole NAL payload once and keep a map
    // from escaped to unescaped offsets for the entry points (which count escaped bytes).
    std::vector<uint8_t> rbsp;
    std::vector<uint32_t> removed_before;     // number of 0x03 bytes removed before escaped offset i (sampled at removals)
    rbsp.reserve(n);
    std::vector<uint32_t> epb_pos;            // escaped offsets of removed bytes
    int zeros = 0;
    for (size_t i = 0; i < n; i++) {
      if (zeros >= 2 && payload[i] == 3) { epb_pos.push_back((uint32_t)i); zeros = 0; continue; }
      rbsp.push_back(payload[i]);
      zeros = payload[i] == 0 ? zeros + 1 : 0;
    }
    BitReader b(rbsp.data(), rbsp.size());
    const bool irap = nal_type >= 16 && nal_type <= 23;
    const bool idr = nal_type == 19 || nal_type == 20;
    if (!b.u(1)) { set_error("decoder: multiple slice segments per picture are not supported"); return -1; }
    if (irap) b.u(1);
    b.ue();
    int slice_type = (int)b.ue();
    if (slice_type == 0) { set_error("decoder: B slices are not supported"); return -1; }
    if (!idr) {
      b.u(sps.log2_max_poc);
      if (!b.u(1)) {
        uint32_t neg = b.ue(), pos = b.ue();
        if (neg != 1 || pos != 0 || b.ue() != 0 || b.u(1) != 1) { set_error("decoder: only the previous picture as reference is supported"); return -1; }
      } else if (sps.num_rps < 1) {
        set_error("decoder: slice refers to a missing RPS"); return -1;
      }
    }
    if (slice_type == 1) {
      if (b.u(1)) {
        if (b.ue() != 0) { set_error("decoder: more than one reference picture is not supported"); return -1; }
      }
      if (b.ue() != 0) { set_error("decoder: MaxNumMergeCand other than 5 is not supported"); return -1; }
    }
    int qp = pps.init_qp + b.se();
    const int deblock = !pps.deblock_disabled;
    if (pps.loop_across_slices && deblock) b.u(1);
    const int rows = fp.ctb_rows;
    int n_entry = (int)b.ue();
    std::vector<uint32_t> entry(n_entry);
    if (n_entry > 0) {
      int len = (int)b.ue() + 1;
      for (int i = 0; i < n_entry; i++) entry[i] = b.u(len) + 1;
    }
    if (!b.u(1)) { set_error("decoder: malformed slice header (alignment bit)"); return -1; }
    b.align();
    if (b.bad || qp < 0 || qp > 51) { set_error("decoder: malformed slice header"); return -1; }
    if (n_entry != rows - 1) { set_error("decoder: %d entry points for %d CTU rows (WPP expected)", n_entry, rows); return -1; }
    // escaped offset of the first slice-data byte
    const size_t hdr_unesc = b.pos >> 3;
    size_t hdr_esc = hdr_unesc;
    for (uint32_t e : epb_pos) { if (e < hdr_esc + 1) hdr_esc++; else break; }
    // substream boundaries in escaped bytes -> unescaped offsets
    auto to_unesc = [&](size_t esc) {
      size_t k = std::lower_bound(epb_pos.begin(), epb_pos.end(), (uint32_t)esc) - epb_pos.begin();
      return esc - k;
    };
    DecSlot &sl = slots[next_slot];
    size_t esc = hdr_esc;
    for (int r = 0; r < rows; r++) {
      sl.h_bases[r] = (uint32_t)(to_unesc(esc) - hdr_unesc);
      if (r < rows - 1) esc += entry[r];
    }
    const size_t data_len = rbsp.size() - hdr_unesc;
    sl.h_bases[rows] = (uint32_t)data_len;
    for (int r = 0; r < rows; r++)
      if (sl.h_bases[r] > sl.h_bases[r + 1]) { set_error("decoder: entry points run past the slice data"); return -1; }
    if (data_len > data_cap) { set_error("decoder: slice larger than the staging buffer"); return -1; }
    memcpy(sl.h_data, rbsp.data() + hdr_unesc, data_len);

    sl.fp = fp;
    sl.fp.qp = qp; sl.fp.qp_c = kChromaQpD[qp]; sl.fp.is_idr = slice_type == 2 ? 1 : 0; sl.fp.deblock = deblock;
    sl.fp.ctu_qp = pps.qp_delta ? sl.d_ctu_qp : nullptr; sl.fp.ctu_delta = nullptr; sl.fp.ctu_first = nullptr;
    sl.pts = pts;
    int *sync_flag = (int *)(sl.d_small + off_flag), *progress = (int *)(sl.d_small + off_prog);
    int *status = (int *)(sl.d_small + off_status);
    uint32_t *d_bases = (uint32_t *)(sl.d_small + off_bases);
#define DEC_CHECK(expr, what) do { if (!cuda_ok((expr), (what))) return -1; } while (0)
    DEC_CHECK(cudaMemcpyAsync(sl.d_data, sl.h_data, data_len, cudaMemcpyHostToDevice, sl.stream), "H2D slice");
    DEC_CHECK(cudaMemcpyAsync(d_bases, sl.h_bases, sizeof(uint32_t) * (rows + 1), cudaMemcpyHostToDevice, sl.stream), "H2D bases");
    DEC_CHECK(cudaMemsetAsync(sl.d_levels, 0, frame_bytes * sizeof(int16_t), sl.stream), "memset levels");
    DEC_CHECK(launch_parse(sl.fp, sl.d_data, d_bases, sl.d_cu, sl.d_levels, sl.d_small + off_ctx, sync_flag, progress, status, sl.stream), "parse launch");
    count_launch(1);
    DEC_CHECK(cudaMemcpyAsync(sl.h_status, status, sizeof(int) * 2, cudaMemcpyDeviceToHost, sl.stream), "D2H status");
    DEC_CHECK(cudaEventRecord(sl.ev_parsed, sl.stream), "record parse");
    pending.push_back(next_slot);
    next_slot = (next_slot + 1) % (int)slots.size();
    if ((int)pending.size() <= frame_delay) return 0;
    return finish_oldest();
  }

  // Reconstructs the oldest parsed picture (prediction + residual, deblocking) and copies it to the
  // host.  Returns 1 with the picture in slots[out_slot].h_out, -1 on error.
  int finish_oldest()
  {
    const int idx = pending.front();
    pending.pop_front();
    DecSlot &sl = slots[idx];
    FrameParams &f = sl.fp;
    int *progress = (int *)(sl.d_small + off_prog), *ticket = (int *)(sl.d_small + off_ticket);
    if (!cuda_ok(cudaEventSynchronize(sl.ev_parsed), "sync parse")) { have_ref = 0; return -1; }
    if (sl.h_status[0] != 0) {
      static const char *const why[] = {"", "escape code too long", "intra CU in a P slice", "partition other than 2Nx2N", "mvd too long",
        "NxN intra partition", "intra chroma mode other than derived", "64x64 CU with residual", "end_of_slice_segment_flag mismatch",
        "end_of_subset_one_bit missing", "intra CU size other than 16x16 (8x8 at the picture edge)", "cu_qp_delta out of range"};
      int c = sl.h_status[0];
      set_error("decoder: unsupported or corrupt slice data (%s)", c > 0 && c <= 11 ? why[c] : "unknown");
      have_ref = 0;
      return -1;
    }
    if (!f.is_idr && !have_ref) { set_error("decoder: P slice without a reference picture"); return -1; }
    uint8_t *rec = d_rec[cur], *ref = d_rec[cur ^ 1];
    have_ref = 0;                          // until this picture is complete
    if (f.is_idr) {
      DEC_CHECK(launch_intra_decode(f, rec, sl.d_levels, sl.d_cu, progress, ticket, d_order, stream), "intra decode launch");
      count_launch(1);
    } else {
      f.search_range = std::max(1, (sl.h_status[1] + 3) / 4 + 1);
      cudaError_t e = launch_inter_decode(f, ref, rec, sl.d_levels, sl.d_cu, stream);
      if (e == cudaErrorInvalidValue) { set_error("decoder: motion vectors of +-%d samples exceed the supported window", f.search_range); return -1; }
      DEC_CHECK(e, "inter decode launch");
      count_launch(1);
    }
    if (f.deblock) {
      DEC_CHECK(launch_deblock(f, rec, sl.d_cu, stream), "deblock launch");
      count_launch(2);
    }
    if (host_output) DEC_CHECK(cudaMemcpyAsync(sl.h_out, rec, frame_bytes, cudaMemcpyDeviceToHost, stream), "D2H picture");
    DEC_CHECK(cudaStreamSynchronize(stream), "sync picture");
    d_out = rec;
#undef DEC_CHECK
    cur ^= 1;
    have_ref = 1; out_slot = idx; out_pts = sl.pts; pictures++;
    return 1;
  }

  // libOpenHevcDecode(h, NULL, 0, pts): hand over the next held-back picture, 0 when none is left.
  int drain_one() { return pending.empty() ? 0 : finish_oldest(); }

  int decode_nal(const uint8_t *nal, size_t n, int64_t pts)
  {
    if (n < 2) return 0;
    if (nal[0] & 0x80) { set_error("decoder: forbidden_zero_bit set"); return -1; }
    const int type = (nal[0] >> 1) & 63;
    if (type == 32) { vps_seen = true; return 0; }
    if (type == 33) return parse_sps(unescape(nal + 2, n - 2)) ? 0 : -1;
    if (type == 34) return parse_pps(unescape(nal + 2, n - 2)) ? 0 : -1;
    if (type <= 9 || (type >= 16 && type <= 21)) return decode_slice(type, nal + 2, n - 2, pts);
    return 0;      // SEI, AUD, ... ignored
  }
};

}  // namespace b200

using b200::Decoder;

extern "C" {

OpenHevc_Handle libOpenHevcInit(int nb_pthreads, int thread_type)
{
  // Frame threading (thread_type 1 or 3) in OpenHEVC keeps nb_pthreads pictures in flight and
  // delays output by nb_pthreads - 1 pictures; here the pictures in flight are concurrent CABAC
  // parses on the GPU.  Slice threading (2) has no delay: WPP substreams are always parallel.
  Decoder *d = new Decoder();
  if (thread_type & 1) d->frame_delay = std::min(std::max(nb_pthreads, 1), 64) - 1;
  return d;
}

int libOpenHevcStartDecoder(OpenHevc_Handle h)
{
  Decoder *d = (Decoder *)h;
  if (!d) return -1;
  if (b200_device_count() <= 0) { b200::set_error("no CUDA device: the B200 decoder has no CPU fallback"); return -1; }
  d->started = true;
  return 0;
}

int libOpenHevcDecode(OpenHevc_Handle h, const unsigned char *buff, int nal_len, int64_t pts)
{
  Decoder *d = (Decoder *)h;
  if (!d || !d->started) { b200::set_error("libOpenHevcDecode: decoder not started"); return -1; }
  if (!buff || nal_len <= 0) return d->drain_one();
  // split on start codes (00 00 01 / 00 00 00 01)
  int got = 0;
  size_t n = (size_t)nal_len;
  auto find_sc = [&](size_t from, size_t &sc_len) -> size_t {
    for (size_t k = from; k + 3 <= n; k++)
      if (buff[k] == 0 && buff[k + 1] == 0 && buff[k + 2] == 1) { sc_len = 3; return k; }
    return n;
  };
  size_t sc_len = 0;
  size_t pos = find_sc(0, sc_len);
  if (pos == n) { b200::set_error("libOpenHevcDecode: no start code in the buffer"); return -1; }
  while (pos < n) {
    size_t start = pos + sc_len;
    size_t next_len = 0;
    size_t next = find_sc(start, next_len);
    size_t end = next;
    while (end > start && next < n && buff[end - 1] == 0) end--;     // zero_byte of a 4-byte start code / trailing zeros
    int rc = d->decode_nal(buff + start, end - start, pts);
    if (rc < 0) return -1;
    got |= rc;
    pos = next; sc_len = next_len;
  }
  return got;
}

int libOpenHevcGetOutput(OpenHevc_Handle h, int got_picture, OpenHevc_Frame *frame)
{
  Decoder *d = (Decoder *)h;
  if (!d || !frame || !got_picture || d->out_slot < 0) return 0;
  const size_t ysz = (size_t)d->fp.w * d->fp.h;
  uint8_t *out = d->slots[d->out_slot].h_out;
  frame->pvY = out;
  frame->pvU = out + ysz;
  frame->pvV = out + ysz + ysz / 4;
  libOpenHevcGetPictureInfo(h, &frame->frameInfo);
  return 1;
}

void libOpenHevcGetPictureInfo(OpenHevc_Handle h, OpenHevc_FrameInfo *info)
{
  Decoder *d = (Decoder *)h;
  if (!d || !info) return;
  memset(info, 0, sizeof(*info));
  info->nWidth = d->fp.w; info->nHeight = d->fp.h;
  info->nYPitch = d->fp.w; info->nUPitch = d->fp.w / 2; info->nVPitch = d->fp.w / 2;
  info->nBitDepth = 8; info->chromat_format = 1;
  info->sample_aspect_ratio.num = 1; info->sample_aspect_ratio.den = 1;
  info->frameRate.num = d->fr_num; info->frameRate.den = d->fr_den;     // 0/0: the stream carries no VUI timing
  info->display_picture_number = d->pictures - 1;
  info->nTimeStamp = d->out_pts;
}

void libOpenHevcSetTemporalLayer_id(OpenHevc_Handle, int) {}
void libOpenHevcSetActiveDecoders(OpenHevc_Handle, int) {}
void libOpenHevcSetViewLayers(OpenHevc_Handle, int) {}
void libOpenHevcSetDebugMode(OpenHevc_Handle, int) {}
void libOpenHevcSetCheckMD5(OpenHevc_Handle, int) {}
const char *libOpenHevcVersion(OpenHevc_Handle) { return "b200-hevc-dec 0.1 (sm_100a)"; }
void libOpenHevcFlush(OpenHevc_Handle h)
{
  Decoder *d = (Decoder *)h;
  if (!d) return;
  for (b200::DecSlot &s : d->slots) cudaStreamSynchronize(s.stream);
  d->pending.clear();
  d->have_ref = 0; d->out_slot = -1; d->d_out = nullptr;
}
void libOpenHevcClose(OpenHevc_Handle h) { delete (Decoder *)h; }

const uint8_t *b200_dec_output_dev(OpenHevc_Handle h)
{
  Decoder *d = (Decoder *)h;
  return d && d->out_slot >= 0 ? d->d_out : NULL;
}

void b200_dec_set_host_output(OpenHevc_Handle h, int on)
{
  Decoder *d = (Decoder *)h;
  if (d) d->host_output = on != 0;
}

int b200_dec_last_picture(OpenHevc_Handle h, uint8_t *dst, int cap)
{
  Decoder *d = (Decoder *)h;
  if (!d || !dst || d->out_slot < 0 || (size_t)cap < d->frame_bytes) return -1;
  memcpy(dst, d->slots[d->out_slot].h_out, d->frame_bytes);
  return (int)d->frame_bytes;
}

}  // extern "C"
