// Host side of the B200 HEVC encoder: owns the HBM-resident state of one stream, sequences the
// per-picture kernels on one CUDA stream, writes the parameter sets / slice header and assembles
// the Annex-B access unit from the per-row substreams the CABAC kernel produced.
//
// This is the engine under the kvz_api C ABI (kvz_api.cu); the b200_enc_* entry points below are
// the same engine exposed with plain arguments for the benchmark and the parity tests.
#include "hevc_encoder.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../include/b200_hevc.h"
#include "hevc_headers.h"
#include "hevc_kernels.h"
#include "runtime.h"

namespace b200 {

// ---- bit writer / NAL helpers -------------------------------------------------------------------

namespace {

struct BitWriter {
  std::vector<uint8_t> bytes;
  uint32_t cur = 0;
  int nbits = 0;
  void put(uint32_t v, int n)
  {
    for (int i = n - 1; i >= 0; i--) {
      cur = (cur << 1) | ((v >> i) & 1);
      if (++nbits == 8) { bytes.push_back((uint8_t)cur); cur = 0; nbits = 0; }
    }
  }
  void ue(uint32_t v)
  {
    uint32_t x = v + 1;
    int len = 0;
    while ((x >> len) > 1) len++;
    put(0, len);
    put(x, len + 1);
  }
  void se(int32_t v) { ue(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
  void trailing()
  {
    put(1, 1);
    while (nbits) put(0, 1);
  }
};

void append_escaped(std::vector<uint8_t> &out, const uint8_t *p, size_t n)
{
  int zeros = 0;
  for (size_t i = 0; i < n; i++) {
    if (zeros >= 2 && p[i] <= 3) { out.push_back(3); zeros = 0; }
    out.push_back(p[i]);
    zeros = p[i] == 0 ? zeros + 1 : 0;
  }
}

void start_nal(std::vector<uint8_t> &out, int type)
{
  static const uint8_t sc[4] = {0, 0, 0, 1};
  out.insert(out.end(), sc, sc + 4);
  out.push_back((uint8_t)(type << 1));
  out.push_back(1);
}

void put_profile_tier_level(BitWriter &b, int level_idc)
{
  b.put(0, 2); b.put(0, 1); b.put(1, 5);       // profile space, tier, Main
  b.put(0x60000000u, 32);                      // compatible with Main and Main 10
  b.put(1, 1); b.put(0, 1); b.put(0, 1); b.put(1, 1);   // progressive, !interlaced, !non-packed, frame-only
  b.put(0, 32); b.put(0, 11);                  // 43 reserved zero bits
  b.put(0, 1);
  b.put((uint32_t)level_idc, 8);
}

int level_for(int w, int h)
{
  long px = (long)w * h;
  return px <= 552960 ? 93 : px <= 983040 ? 120 : px <= 2228224 ? 123 : px <= 8912896 ? 153 : 183;
}

const uint16_t kLambdaQ4[52] = {   // round(16 * sqrt(0.57 * 2^((qp-12)/3)))
  3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15, 17, 19, 22, 24, 27, 30, 34, 38, 43, 48, 54, 61, 68,
  77, 86, 97, 108, 122, 137, 153, 172, 193, 217, 244, 273, 307, 344, 387, 434, 487, 547, 614, 689, 773,
  868, 974, 1093};

const uint8_t kChromaQp[58] = {
  0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
  29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};

#define ENC_CHECK(expr, what)                                  \
  do {                                                         \
    if (!cuda_ok((expr), (what))) return false;                \
  } while (0)

}  // namespace

// ---- Encoder ------------------------------------------------------------------------------------

Encoder::~Encoder() { release(); }

void Encoder::release()
{
  if (stream) cudaStreamSynchronize(stream);
  for (FrameSlot &s : slots) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.d_cu) cudaFree(s.d_cu);
    if (s.d_levels) cudaFree(s.d_levels);
    if (s.d_rows) cudaFree(s.d_rows);
    if (s.d_recs) cudaFree(s.d_recs);
    if (s.d_small) cudaFree(s.d_small);
    if (s.d_qpinfo) cudaFree(s.d_qpinfo);
    if (s.d_ctu_done) cudaFree(s.d_ctu_done);
    if (s.d_dbk) cudaFree(s.d_dbk);
    if (s.d_sao) cudaFree(s.d_sao);
    if (s.h_ctu_qp) cudaFreeHost(s.h_ctu_qp);
    if (s.d_vaq) cudaFree(s.d_vaq);
    if (s.d_src) cudaFree(s.d_src);
    if (s.h_src) cudaFreeHost(s.h_src);
    if (s.h_pack) cudaFreeHost(s.h_pack);
    if (s.h_hdr) cudaFreeHost(s.h_hdr);
    for (cudaEvent_t &e : s.pev) if (e) cudaEventDestroy(e);
    if (s.ev_pred) cudaEventDestroy(s.ev_pred);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  slots.clear();
  inflight.clear();
  if (intra_stream) cudaStreamSynchronize(intra_stream);
  for (int i = 0; i < kRecRing; i++) { if (d_rec[i]) cudaFree(d_rec[i]); if (ev_ring[i]) cudaEventDestroy(ev_ring[i]); d_rec[i] = nullptr; ev_ring[i] = nullptr; }
  if (d_order) cudaFree(d_order);
  d_order = nullptr;
  if (ev_intra) cudaEventDestroy(ev_intra);
  if (intra_stream) cudaStreamDestroy(intra_stream);
  if (upload_stream) { cudaStreamSynchronize(upload_stream); cudaStreamDestroy(upload_stream); }
  if (ev_upload) cudaEventDestroy(ev_upload);
  upload_stream = nullptr; ev_upload = nullptr;
  ev_intra = nullptr; intra_stream = nullptr;
  if (d_rec_pre) cudaFree(d_rec_pre);
  if (d_me_stats) cudaFree(d_me_stats);
  if (d_scaling) cudaFree(d_scaling);
  d_me_stats = nullptr; d_scaling = nullptr;
  if (d_src_q) cudaFree(d_src_q);
  if (d_ref_q) cudaFree(d_ref_q);
  d_src_q = d_ref_q = nullptr;
  if (ev_base) cudaEventDestroy(ev_base);
  if (stream) cudaStreamDestroy(stream);
  d_rec_pre = nullptr; stream = nullptr; ev_base = nullptr;
}

bool Encoder::open(const EncoderConfig &c)
{
  if (c.width <= 0 || c.height <= 0 || (c.width & 7) || (c.height & 7)) { set_error("encoder: width/height must be positive multiples of 8 (got %dx%d)", c.width, c.height); return false; }
  if (c.qp < 0 || c.qp > 51) { set_error("encoder: qp %d out of range 0..51", c.qp); return false; }
  if (c.search_range < 1 || c.search_range > 32) { set_error("encoder: search range %d out of range 1..32", c.search_range); return false; }
  if (c.me_coarse < 0 || c.me_coarse > 32 || (c.me_coarse & 3) || (c.me_coarse > 0 && c.search_range > 16)) { set_error("encoder: me_coarse %d must be a multiple of 4 in 0..32 (and search_range <= 16 with it)", c.me_coarse); return false; }
  if (c.src_width || c.src_height) {
    const int sw = c.src_width ? c.src_width : c.width, sh = c.src_height ? c.src_height : c.height;
    if (sw <= 0 || sh <= 0 || (sw & 1) || (sh & 1) || sw > c.width || sh > c.height || c.width - sw > 6 || c.height - sh > 6) {
      set_error("encoder: source size %dx%d must be even and at most 6 samples short of the coded size %dx%d", sw, sh, c.width, c.height);
      return false;
    }
  }
  if (c.vaq < 0 || c.vaq > 20 || (c.vaq && !c.qp_delta)) { set_error("encoder: vaq %d must be 0..20 and needs qp_delta", c.vaq); return false; }
  if (c.depth < 1 || c.depth > 128) { set_error("encoder: depth %d out of range 1..128", c.depth); return false; }
  if (b200_device_count() <= 0) { set_error("no CUDA device: the B200 encoder has no CPU fallback"); return false; }
  cfg = c;
  if (const char *ev = getenv("B200_OVERLAP_IDR")) cfg.overlap_idr = atoi(ev);
  fp.w = c.width; fp.h = c.height; fp.w8 = c.width / 8; fp.h8 = c.height / 8;
  fp.ctb_cols = (c.width + kCtb - 1) / kCtb; fp.ctb_rows = (c.height + kCtb - 1) / kCtb;
  fp.qp = c.qp; fp.qp_c = kChromaQp[std::min(std::max(c.qp, 0), 57)];
  fp.lambda_q4 = kLambdaQ4[c.qp];
  fp.search_range = c.search_range; fp.is_idr = 1; fp.deblock = c.deblock;
  frame_bytes = (size_t)fp.w * fp.h * 3 / 2;
  src_w = c.src_width ? c.src_width : c.width; src_h = c.src_height ? c.src_height : c.height;
  src_bytes = (size_t)src_w * src_h * 3 / 2;
  row_cap = (uint32_t)fp.w * kCtb * 4 + 4096;
  pack_cap = (uint32_t)std::min<size_t>((size_t)row_cap * fp.ctb_rows, frame_bytes * 3 + 65536);
  // Stream priorities: the entropy-coding kernels are tiny (one warp per CTU row) but long-running
  // and need all their rows resident to make progress, while the motion-search / reconstruction
  // CTAs fill every register of an SM.  Entropy streams therefore get the HIGH priority, so that
  // whenever a prediction CTA retires, pending entropy CTAs are placed first; the prediction chain
  // runs at the low priority.
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (const char *ev = getenv("B200_FLAT_PRIO")) { if (atoi(ev)) prio_hi = prio_lo; }
  ENC_CHECK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_lo), "cudaStreamCreate");
  ENC_CHECK(cudaStreamCreateWithPriority(&intra_stream, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate intra");
  ENC_CHECK(cudaEventCreateWithFlags(&ev_intra, cudaEventDisableTiming), "cudaEventCreate");
  ENC_CHECK(cudaStreamCreateWithPriority(&upload_stream, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate upload");
  ENC_CHECK(cudaEventCreateWithFlags(&ev_upload, cudaEventDisableTiming), "cudaEventCreate");
  {
    std::vector<int> order((size_t)fp.ctb_cols * fp.ctb_rows);
    intra_wavefront_order(fp.ctb_cols, fp.ctb_rows, order.data());
    ENC_CHECK(cudaMalloc((void **)&d_order, order.size() * sizeof(int)), "cudaMalloc order");
    ENC_CHECK(cudaMemcpy(d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice), "H2D order");
  }
  for (int i = 0; i < kRecRing; i++) {
    ENC_CHECK(cudaMalloc((void **)&d_rec[i], frame_bytes), "cudaMalloc rec");
    ENC_CHECK(cudaEventCreateWithFlags(&ev_ring[i], cudaEventDisableTiming), "cudaEventCreate");
  }
  if (c.debug) ENC_CHECK(cudaMalloc((void **)&d_rec_pre, frame_bytes), "cudaMalloc rec_pre");
  if (c.me_coarse > 0) {
    ENC_CHECK(cudaMalloc((void **)&d_src_q, (size_t)(fp.w / 4) * (fp.h / 4)), "cudaMalloc src_q");
    ENC_CHECK(cudaMalloc((void **)&d_ref_q, (size_t)(fp.w / 4) * (fp.h / 4)), "cudaMalloc ref_q");
  }
  // small state: row_len[rows] | sync_flag[rows] | progress[rows] | ticket | bins(8) | sync_ctx[rows*CTX_COUNT]
  off_flag = sizeof(int) * fp.ctb_rows; off_prog = 2 * off_flag; off_ticket = 3 * off_flag;
  off_bins = (off_ticket + sizeof(int) + 7) & ~(size_t)7; off_ctx = off_bins + 8;
  small_bytes = off_ctx + (size_t)fp.ctb_rows * CTX_COUNT;
  slots.resize(c.depth);
  for (FrameSlot &s : slots) {
    ENC_CHECK(cudaMalloc((void **)&s.d_cu, sizeof(CuInfo) * fp.w8 * fp.h8), "cudaMalloc cu");
    ENC_CHECK(cudaMemset(s.d_cu, 0, sizeof(CuInfo) * fp.w8 * fp.h8), "memset cu");
    ENC_CHECK(cudaMalloc((void **)&s.d_levels, frame_bytes * sizeof(int16_t)), "cudaMalloc levels");
    ENC_CHECK(cudaMalloc((void **)&s.d_rows, (size_t)row_cap * fp.ctb_rows), "cudaMalloc rows");
    ENC_CHECK(cudaMalloc((void **)&s.d_recs, (size_t)fp.ctb_cols * fp.ctb_rows * 64 * kRecUnitCap * sizeof(uint32_t)), "cudaMalloc recs");
    ENC_CHECK(cudaMalloc((void **)&s.d_small, small_bytes), "cudaMalloc small");
    if (c.qp_delta) {
      ENC_CHECK(cudaMalloc((void **)&s.d_qpinfo, 3 * (size_t)fp.ctb_cols * fp.ctb_rows), "cudaMalloc qp info");
      ENC_CHECK(cudaMallocHost((void **)&s.h_ctu_qp, (size_t)fp.ctb_cols * fp.ctb_rows), "cudaMallocHost ctu qp");
      if (c.vaq) ENC_CHECK(cudaMalloc((void **)&s.d_vaq, 6 * sizeof(uint32_t) * fp.ctb_cols * fp.ctb_rows), "cudaMalloc vaq");
    }
    ENC_CHECK(cudaMemset(s.d_small, 0, small_bytes), "memset small");
    ENC_CHECK(cudaMalloc((void **)&s.d_ctu_done, sizeof(int) * (fp.ctb_cols * fp.ctb_rows + 1)), "cudaMalloc ctu done");
    if (c.sao) {
      ENC_CHECK(cudaMalloc((void **)&s.d_dbk, frame_bytes), "cudaMalloc dbk");
      ENC_CHECK(cudaMalloc((void **)&s.d_sao, sizeof(SaoCtu) * fp.ctb_cols * fp.ctb_rows), "cudaMalloc sao");
    }
    ENC_CHECK(cudaMalloc((void **)&s.d_src, frame_bytes), "cudaMalloc src");
    ENC_CHECK(cudaMallocHost((void **)&s.h_src, frame_bytes), "cudaMallocHost src");
    ENC_CHECK(cudaHostAlloc((void **)&s.h_pack, pack_cap, cudaHostAllocMapped), "cudaHostAlloc pack");
    ENC_CHECK(cudaHostAlloc((void **)&s.h_hdr, sizeof(uint32_t) * (fp.ctb_rows + 2), cudaHostAllocMapped), "cudaHostAlloc hdr");
    ENC_CHECK(cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate slot");
    ENC_CHECK(cudaEventCreateWithFlags(&s.ev_pred, cudaEventDisableTiming), "cudaEventCreate");
    ENC_CHECK(cudaEventCreateWithFlags(&s.ev_done, wait_event_flags(c.depth > 1)), "cudaEventCreate");
    for (cudaEvent_t &e : s.pev) ENC_CHECK(cudaEventCreate(&e), "cudaEventCreate");
  }
  if (c.scaling_list) {
    ScalingTable t;
    t.set_default();
    ENC_CHECK(cudaMalloc((void **)&d_scaling, sizeof(t)), "cudaMalloc scaling lists");
    ENC_CHECK(cudaMemcpy(d_scaling, &t, sizeof(t), cudaMemcpyHostToDevice), "H2D scaling lists");
  }
  fp.scaling = d_scaling;
  ENC_CHECK(cudaMalloc((void **)&d_me_stats, 4 * sizeof(unsigned long long)), "cudaMalloc me stats");
  ENC_CHECK(cudaMemset(d_me_stats, 0, 4 * sizeof(unsigned long long)), "memset me stats");
  ENC_CHECK(cudaEventCreate(&ev_base), "cudaEventCreate");
  ENC_CHECK(cudaEventRecord(ev_base, stream), "event record");
  frame_idx = 0; poc = 0; cur = 0; cur_qp = c.qp; idr_count = 0;
  return true;
}

void write_parameter_sets(const StreamLayout &l, std::vector<uint8_t> &out)
{
  const int level = level_for(l.w, l.h);
  {  // VPS (7.3.2.1)
    BitWriter b;
    b.put(0, 4); b.put(1, 1); b.put(1, 1); b.put(0, 6); b.put(0, 3); b.put(1, 1); b.put(0xffff, 16);
    put_profile_tier_level(b, level);
    b.put(0, 1);
    b.ue(1); b.ue(0); b.ue(0);               // max_dec_pic_buffering_minus1, num_reorder, latency
    b.put(0, 6); b.ue(0); b.put(0, 1); b.put(0, 1);
    b.trailing();
    start_nal(out, 32);
    append_escaped(out, b.bytes.data(), b.bytes.size());
  }
  {  // SPS (7.3.2.2)
    BitWriter b;
    b.put(0, 4); b.put(0, 3); b.put(1, 1);
    put_profile_tier_level(b, level);
    b.ue(0); b.ue(1);                        // sps id, chroma_format_idc 4:2:0
    b.ue((uint32_t)l.w); b.ue((uint32_t)l.h);
    if (l.conf_right || l.conf_bottom) {     // conformance window, offsets in chroma samples (4:2:0)
      b.put(1, 1);
      b.ue(0); b.ue((uint32_t)l.conf_right / 2); b.ue(0); b.ue((uint32_t)l.conf_bottom / 2);
    } else {
      b.put(0, 1);
    }
    b.ue(0); b.ue(0);                        // bit depths
    b.ue(4);                                 // log2_max_pic_order_cnt_lsb_minus4
    b.put(0, 1); b.ue(1); b.ue(0); b.ue(0);  // sub-layer ordering info
    b.ue(0); b.ue(3);                        // CB 8..64
    b.ue(0); b.ue(3);                        // TB 4..32
    b.ue(0); b.ue(0);                        // max_transform_hierarchy_depth inter / intra
    b.put(l.scaling_list ? 1 : 0, 1);        // scaling_list_enabled_flag
    if (l.scaling_list) b.put(0, 1);         // sps_scaling_list_data_present_flag = 0: the default lists
    b.put(0, 1); b.put(l.sao ? 1 : 0, 1); b.put(0, 1);   // AMP off; SAO; PCM off
    b.ue(1);                                 // one short-term RPS: the previous picture
    b.ue(1); b.ue(0); b.ue(0); b.put(1, 1);
    b.put(0, 1); b.put(0, 1); b.put(0, 1);   // long-term refs, TMVP, strong intra smoothing off
    if (l.fps_num > 0 && l.fps_den > 0) {    // VUI (E.2.1) with nothing but the timing info
      b.put(1, 1);
      b.put(0, 1); b.put(0, 1); b.put(0, 1); b.put(0, 1);   // aspect ratio, overscan, video signal type, chroma loc
      b.put(0, 3); b.put(0, 1);              // neutral chroma / field seq / frame-field info, default display window
      b.put(1, 1);                           // vui_timing_info_present_flag
      b.put((uint32_t)l.fps_den, 32);        // vui_num_units_in_tick
      b.put((uint32_t)l.fps_num, 32);        // vui_time_scale
      b.put(0, 1); b.put(0, 1);              // poc proportional to timing, HRD
      b.put(0, 1);                           // bitstream_restriction_flag
    } else {
      b.put(0, 1);
    }
    b.put(0, 1);                             // sps_extension_present_flag
    b.trailing();
    start_nal(out, 33);
    append_escaped(out, b.bytes.data(), b.bytes.size());
  }
  {  // PPS (7.3.2.3)
    BitWriter b;
    b.ue(0); b.ue(0);
    b.put(0, 1); b.put(0, 1); b.put(0, 3); b.put(0, 1); b.put(0, 1);
    b.ue(0); b.ue(0);
    b.se(0);                                 // init_qp_minus26
    b.put(0, 1); b.put(0, 1);                // constrained intra, transform skip
    b.put(l.qp_delta ? 1 : 0, 1);          // cu_qp_delta_enabled_flag
    if (l.qp_delta) b.ue(0);               // diff_cu_qp_delta_depth: one quantisation group per CTB
    b.se(0); b.se(0);
    b.put(0, 1); b.put(0, 1); b.put(0, 1); b.put(0, 1);
    b.put(l.tile_cols > 1 || l.tile_rows > 1 ? 1 : 0, 1);       // tiles_enabled_flag
    b.put(l.wpp ? 1 : 0, 1);                 // entropy_coding_sync_enabled_flag
    if (l.tile_cols > 1 || l.tile_rows > 1) {
      b.ue((uint32_t)(l.tile_cols - 1));     // num_tile_columns_minus1
      b.ue((uint32_t)(l.tile_rows - 1));     // num_tile_rows_minus1
      b.put(1, 1);                           // uniform_spacing_flag
      b.put(0, 1);                           // loop_filter_across_tiles_enabled_flag
    }
    b.put(1, 1);                             // pps_loop_filter_across_slices_enabled_flag
    if (l.deblock) {
      b.put(0, 1);
    } else {
      b.put(1, 1); b.put(0, 1); b.put(1, 1);
    }
    b.put(0, 1); b.put(0, 1);
    b.ue(0);
    b.put(0, 1); b.put(0, 1);
    b.trailing();
    start_nal(out, 34);
    append_escaped(out, b.bytes.data(), b.bytes.size());
  }
}

void write_slice_nal(const StreamLayout &l, bool idr, int poc, int qp, const uint32_t *sub_len, int n_sub,
                     const uint8_t *data, size_t data_len, std::vector<uint8_t> &out)
{
  BitWriter b;
  b.put(1, 1);
  if (idr) b.put(0, 1);
  b.ue(0);
  b.ue(idr ? 2 : 1);
  if (!idr) {
    b.put((uint32_t)(poc & 255), 8);
    b.put(1, 1);
  }
  if (l.sao) b.put(3, 2);                    // slice_sao_luma_flag, slice_sao_chroma_flag
  if (!idr) {
    b.put(0, 1);
    b.ue(5 - kMaxMerge);
  }
  b.se(qp - 26);
  if (l.deblock || l.sao) b.put(1, 1);       // slice_loop_filter_across_slices_enabled_flag
  b.ue((uint32_t)(n_sub - 1));
  if (n_sub > 1) {
    uint32_t mx = 1;
    for (int r = 0; r < n_sub - 1; r++) mx = std::max(mx, sub_len[r]);
    int len = 1;
    while (((mx - 1) >> len) > 0) len++;
    b.ue((uint32_t)(len - 1));
    for (int r = 0; r < n_sub - 1; r++) b.put(sub_len[r] - 1, len);
  }
  b.trailing();
  start_nal(out, idr ? 19 : 1);
  append_escaped(out, b.bytes.data(), b.bytes.size());
  out.insert(out.end(), data, data + data_len);               // substreams are already escaped
}

StreamLayout Encoder::layout() const
{
  StreamLayout l;
  l.w = fp.w; l.h = fp.h; l.deblock = cfg.deblock; l.qp_delta = cfg.qp_delta; l.tile_cols = 1; l.wpp = cfg.no_wpp ? 0 : 1;
  l.fps_num = cfg.fps_num; l.fps_den = cfg.fps_den; l.sao = cfg.sao; l.scaling_list = cfg.scaling_list;
  l.conf_right = fp.w - src_w; l.conf_bottom = fp.h - src_h;
  return l;
}

// Enqueue everything for one picture; returns without waiting for the GPU.
bool Encoder::submit(FrameSlot &s, const uint8_t *d_i420)
{
  const bool idr = frame_idx == 0 || (cfg.intra_period > 0 && frame_idx % cfg.intra_period == 0);
  if (idr) poc = 0;
  s.idr = idr; s.poc = poc; s.qp = cur_qp; s.seq = frame_idx;
  s.headers = idr && (idr_count == 0 || (cfg.vps_period > 0 && idr_count % cfg.vps_period == 0));
  if (idr) idr_count++;
  FrameParams p = fp;
  p.is_idr = idr ? 1 : 0;
  p.qp = cur_qp; p.qp_c = kChromaQp[cur_qp]; p.lambda_q4 = kLambdaQ4[cur_qp];
  uint8_t *const out_rec = d_rec[frame_idx % kRecRing], *ref = d_rec[(frame_idx + kRecRing - 1) % kRecRing];
  // with SAO the prediction chain reconstructs and deblocks into the slot's own picture; SAO reads it
  // (a CTU needs its neighbours' DEBLOCKED samples) and writes the reconstruction ring
  uint8_t *rec = cfg.sao ? s.d_dbk : out_rec;
  p.ctu_done = s.d_ctu_done; p.any_intra = s.d_ctu_done + fp.ctb_cols * fp.ctb_rows; p.intra_in_p = cfg.intra_in_p; p.intra_satd = cfg.intra_satd; p.subme_satd = cfg.subme_satd;
  p.init_type = idr ? 0 : 1; p.tr_depth_inter = 0; p.tr_depth_intra = 0;
  p.n_refs = 1; p.max_merge = kMaxMerge; p.ref_dist[0] = 1; p.col_mvf = nullptr;
  p.me_stats = profile ? d_me_stats : nullptr;
  p.me_coarse = cfg.me_coarse; p.src_q = d_src_q; p.ref_q = d_ref_q;
  p.sao = cfg.sao ? s.d_sao : nullptr; p.sao_flags = cfg.sao ? (cfg.sao == 2 ? 7 : 3) : 0;
  p.ctu_qp = nullptr; p.ctu_delta = nullptr; p.ctu_first = nullptr;
  p.mv_edges = cfg.mv_edges; p.more_tiles = cfg.more_tiles; p.no_wpp = cfg.no_wpp;
  if (cfg.qp_delta) {
    const int ctus = fp.ctb_cols * fp.ctb_rows;
    cudaStream_t qs = (idr && cfg.overlap_idr) ? intra_stream : stream;       // the stream that consumes the input picture
    for (int i = 0; i < ctus; i++) {
      const int q = cur_qp + (ctu_dqp.empty() ? 0 : ctu_dqp[i]);
      s.h_ctu_qp[i] = cfg.vaq ? (uint8_t)(std::min(std::max(q, -kVaqBias), 127) + kVaqBias) : (uint8_t)std::min(std::max(q, 0), 51);
    }
    p.ctu_qp = s.d_qpinfo; p.ctu_delta = (int8_t *)(s.d_qpinfo + ctus); p.ctu_first = s.d_qpinfo + 2 * ctus;
    ENC_CHECK(cudaMemcpyAsync(s.d_qpinfo, s.h_ctu_qp, ctus, cudaMemcpyHostToDevice, qs), "H2D ctu qp");
    if (cfg.vaq) {                       // every CTU's QP moves with its sample variance against the picture's
      ENC_CHECK(launch_vaq(p, d_i420, cfg.vaq, s.d_vaq, s.d_qpinfo, qs), "vaq launch");
      count_launch(2);
    }
  }
  uint32_t *row_len = (uint32_t *)s.d_small;
  int *sync_flag = (int *)(s.d_small + off_flag), *ticket = (int *)(s.d_small + off_ticket);
  unsigned long long *bins = (unsigned long long *)(s.d_small + off_bins);
  s.prof_mask = 0;
#define PROF_BEGIN(id, st) do { if (profile) { ENC_CHECK(cudaEventRecord(s.pev[2 * (id)], st), "prof event"); s.prof_mask |= 1u << (id); } } while (0)
#define PROF_END(id, st) do { if (profile) ENC_CHECK(cudaEventRecord(s.pev[2 * (id) + 1], st), "prof event"); } while (0)
  if (idr) {
    // An IDR depends on no other picture: it runs on its own stream, concurrently with the P
    // pictures still queued on the main stream.  Its buffer ring[n % R] was last read (as a
    // reference) by picture n - R + 1, so that is the only thing to wait for (the input copy of
    // an IDR is enqueued on the intra stream as well, see input_stream()).
    cudaStream_t is = cfg.overlap_idr ? intra_stream : stream;
    if (cfg.overlap_idr && frame_idx >= kRecRing - 1) ENC_CHECK(cudaStreamWaitEvent(is, ev_ring[(frame_idx + 1) % kRecRing], 0), "stream wait");
    // (no memset of the level planes: both reconstruction kernels write every level of the picture)
    PROF_BEGIN(K_INTRA, is);
    ENC_CHECK(launch_intra_frame(p, d_i420, rec, s.d_levels, s.d_cu, ticket, d_order, is), "intra launch");
    PROF_END(K_INTRA, is);
    if (cfg.overlap_idr) {
      ENC_CHECK(cudaEventRecord(ev_intra, is), "event record");
      ENC_CHECK(cudaStreamWaitEvent(stream, ev_intra, 0), "stream wait");     // the prediction chain continues after it
    }
    count_launch(2);
  } else {
    // prediction chain, main stream, picture order
    if (cfg.intra_in_p) ENC_CHECK(cudaMemsetAsync(p.any_intra, 0, sizeof(int), stream), "memset any_intra");
    PROF_BEGIN(K_ME, stream);
    if (cfg.me_coarse > 0) { ENC_CHECK(launch_down4(d_i420, fp.w, fp.h, d_src_q, stream), "down4 launch"); count_launch(1); }
    ENC_CHECK(launch_inter_me(p, d_i420, ref, s.d_cu, stream), "me launch");
    PROF_END(K_ME, stream);
    PROF_BEGIN(K_RECON, stream);
    ENC_CHECK(launch_inter_recon(p, d_i420, ref, rec, s.d_levels, s.d_cu, stream), "recon launch");
    if (cfg.intra_in_p) {                // the intra CUs the search chose predict from the reconstructed inter CUs
      ENC_CHECK(launch_intra_in_p(p, d_i420, rec, s.d_levels, s.d_cu, ticket, d_order, stream), "intra-in-P launch");
      count_launch(1);
    }
    PROF_END(K_RECON, stream);
    PROF_BEGIN(K_MODES, stream);
    ENC_CHECK(launch_inter_modes(p, s.d_cu, stream), "modes launch");
    PROF_END(K_MODES, stream);
    count_launch(3);
  }
  if (cfg.qp_delta) {                    // per-CU QPs and the delta each CTU codes (deblocking and binarisation read them)
    ENC_CHECK(launch_cu_qps(p, s.d_cu, stream), "cu qp launch");
    count_launch(1);
  }
  ENC_CHECK(cudaEventRecord(ev_ring[frame_idx % kRecRing], stream), "event record");   // done reading the reference
  if (cfg.debug) ENC_CHECK(cudaMemcpyAsync(d_rec_pre, rec, frame_bytes, cudaMemcpyDeviceToDevice, stream), "copy pre-deblock");
  if (cfg.deblock) {
    PROF_BEGIN(K_DEBLOCK, stream);
    ENC_CHECK(launch_deblock(p, rec, s.d_cu, stream), "deblock launch");
    PROF_END(K_DEBLOCK, stream);
    count_launch(2);
  }
  if (cfg.sao) {
    PROF_BEGIN(K_SAO, stream);
    ENC_CHECK(launch_sao_encode(p, d_i420, rec, out_rec, s.d_sao, stream), "sao launch");
    PROF_END(K_SAO, stream);
    count_launch(1);
  }
  // the finished picture at quarter resolution: the coarse search level of the next picture reads it
  if (cfg.me_coarse > 0) { ENC_CHECK(launch_down4(out_rec, fp.w, fp.h, d_ref_q, stream), "down4 launch"); count_launch(1); }
  // Entropy coding (slot stream) needs only the cu map and the levels, but it is released after
  // the deblocking: started before it, the binariser's 16k CTAs share the SMs with the two short
  // deblocking kernels and stretch them from 15 us to 75 us on the critical prediction chain.
  ENC_CHECK(cudaEventRecord(s.ev_pred, stream), "event record");
  ENC_CHECK(cudaStreamWaitEvent(s.stream, s.ev_pred, 0), "stream wait");
  PROF_BEGIN(K_BINARISE, s.stream);
  ENC_CHECK(launch_binarise(p, s.d_cu, s.d_levels, s.d_recs, s.stream), "binarise launch");
  PROF_END(K_BINARISE, s.stream);
  PROF_BEGIN(K_ARITH, s.stream);
  ENC_CHECK(launch_arith(p, s.d_cu, s.d_recs, s.d_rows, row_cap, row_len, s.d_small + off_ctx, sync_flag, bins, cfg.depth <= 2, s.stream), "arith launch");
  PROF_END(K_ARITH, s.stream);
  PROF_BEGIN(K_PACK, s.stream);
  ENC_CHECK(launch_pack_rows(p.ctb_rows, s.d_rows, row_cap, row_len, s.h_pack, pack_cap, s.h_hdr, s.stream), "pack launch");
  PROF_END(K_PACK, s.stream);
  count_launch(2);                                 // entropy coding: k_binarise, k_entropy_rows ...
#undef PROF_BEGIN
#undef PROF_END
  count_launch(cfg.depth <= 2 ? 1 : 2);            // ... (k_arith_rows when the phases are two launches,) k_pack_rows
  ENC_CHECK(cudaEventRecord(s.ev_done, s.stream), "event record");
  frame_idx++;
  poc++;
  return true;
}

// Wait for a submitted picture and assemble its access unit.
bool Encoder::collect(FrameSlot &s, std::vector<uint8_t> &out)
{
  ENC_CHECK(cudaEventSynchronize(s.ev_done), "wait for picture");
  if (s.prof_mask) {
    // the deblocking of this picture (main stream) may still be running: wait for ITS end event
    // only -- synchronising the whole main stream would drain the pictures submitted after it
    if (s.prof_mask & (1u << K_DEBLOCK)) cudaEventSynchronize(s.pev[2 * K_DEBLOCK + 1]);
    if (s.prof_mask & (1u << K_SAO)) cudaEventSynchronize(s.pev[2 * K_SAO + 1]);
    for (int k = 0; k < K_COUNT; k++) {
      if (!(s.prof_mask & (1u << k))) continue;
      float ms = 0;
      if (cudaEventElapsedTime(&ms, s.pev[2 * k], s.pev[2 * k + 1]) == cudaSuccess) { prof_ms[k] += ms; prof_cnt[k]++; }
      cudaEventElapsedTime(&timeline[2 * k], ev_base, s.pev[2 * k]);
      cudaEventElapsedTime(&timeline[2 * k + 1], ev_base, s.pev[2 * k + 1]);
    }
    s.prof_mask = 0;
  }
  if (s.h_hdr[0] == 0xffffffffu || s.h_hdr[0] == 0) {
    set_error("encoder: a substream overflowed its %u-byte buffer (picture %lld)", row_cap, s.seq);
    return false;
  }
  out.clear();
  const int n_sub = cfg.no_wpp ? 1 : fp.ctb_rows;
  last_idr = s.idr; last_qp = s.qp; last_poc = s.poc; last_headers = s.headers;
  if (cfg.raw) {                         // tile-column mode: the compositor writes the headers
    last_sub_len.assign(s.h_hdr + 1, s.h_hdr + 1 + n_sub);
    last_data = s.h_pack; last_data_len = s.h_hdr[0];
    out.push_back(0);                    // "a picture is ready"
    return true;
  }
  if (s.headers) write_parameter_sets(layout(), out);
  write_slice_nal(layout(), s.idr, s.poc, s.qp, s.h_hdr + 1, n_sub, s.h_pack, s.h_hdr[0], out);
  return true;
}

void Encoder::set_qp(int qp) { cur_qp = std::min(std::max(qp, 0), 51); }

bool Encoder::set_ctu_dqp(const int8_t *dqp, int n)
{
  if (!cfg.qp_delta) { set_error("encoder: per-CTU QP offsets need an encoder opened with qp_delta"); return false; }
  if (!dqp) { ctu_dqp.clear(); return true; }
  if (n != fp.ctb_cols * fp.ctb_rows) { set_error("encoder: %d QP offsets for %d CTUs", n, fp.ctb_cols * fp.ctb_rows); return false; }
  ctu_dqp.assign(dqp, dqp + n);
  return true;
}

// stream that consumes the next picture's input: the intra stream for an IDR, else the main stream
cudaStream_t Encoder::input_stream() const
{
  const bool idr = frame_idx == 0 || (cfg.intra_period > 0 && frame_idx % cfg.intra_period == 0);
  return (idr && cfg.overlap_idr) ? intra_stream : stream;
}

bool Encoder::encode_device(const uint8_t *d_i420, std::vector<uint8_t> &out)
{
  out.clear();
  const int slot = (int)(frame_idx % cfg.depth);
  FrameSlot &s = slots[slot];
  // the slot is free: its previous picture was collected when the pipeline was full.
  // Take a private copy so that the caller may reuse its buffer as soon as this call returns
  // (the caller orders its producer before this call; the copy is ~1 us of HBM time).
  if (d_i420 != s.d_src) {
    if (padded()) { if (!stage_source(s, d_i420, cudaMemcpyDeviceToDevice, input_stream())) return false; }
    else ENC_CHECK(cudaMemcpyAsync(s.d_src, d_i420, frame_bytes, cudaMemcpyDeviceToDevice, input_stream()), "D2D frame");
  }
  if (padded()) { ENC_CHECK(launch_pad_edges(s.d_src, fp.w, fp.h, src_w, src_h, input_stream()), "pad launch"); count_launch(1); }
  if (!submit(s, s.d_src)) return false;
  inflight.push_back(slot);
  last_slot = slot;
  if ((int)inflight.size() >= cfg.depth) return flush(out);
  return true;
}

// The three planes of a src_w x src_h picture into the top-left corner of the coded picture
bool Encoder::stage_source(FrameSlot &s, const uint8_t *pic, cudaMemcpyKind kind, cudaStream_t st)
{
  const size_t sy = (size_t)src_w * src_h, ysz = (size_t)fp.w * fp.h;
  ENC_CHECK(cudaMemcpy2DAsync(s.d_src, fp.w, pic, src_w, src_w, src_h, kind, st), "copy source Y");
  ENC_CHECK(cudaMemcpy2DAsync(s.d_src + ysz, fp.w / 2, pic + sy, src_w / 2, src_w / 2, src_h / 2, kind, st), "copy source U");
  ENC_CHECK(cudaMemcpy2DAsync(s.d_src + ysz + ysz / 4, fp.w / 2, pic + sy + sy / 4, src_w / 2, src_w / 2, src_h / 2, kind, st), "copy source V");
  return true;
}

bool Encoder::encode_host(const uint8_t *i420, std::vector<uint8_t> &out, bool pinned)
{
  FrameSlot &s = slots[frame_idx % cfg.depth];
  const uint8_t *from = i420;
  if (!pinned) {
    memcpy(s.h_src, i420, src_bytes);
    from = s.h_src;
  }
  // the upload runs on its own stream so that the copy engine works while the previous picture's
  // kernels execute; the consuming stream waits for it
  if (padded()) { if (!stage_source(s, from, cudaMemcpyHostToDevice, upload_stream)) return false; }
  else ENC_CHECK(cudaMemcpyAsync(s.d_src, from, frame_bytes, cudaMemcpyHostToDevice, upload_stream), "H2D frame");
  ENC_CHECK(cudaEventRecord(ev_upload, upload_stream), "event record");
  ENC_CHECK(cudaStreamWaitEvent(input_stream(), ev_upload, 0), "stream wait");
  return encode_device(s.d_src, out);
}

bool Encoder::encode_host_strip(const uint8_t *pic, int pic_w, int pic_h, int x0, int y0, std::vector<uint8_t> &out)
{
  FrameSlot &s = slots[frame_idx % cfg.depth];
  const int w = fp.w, h = fp.h;
  const size_t pic_y = (size_t)pic_w * pic_h, ysz = (size_t)w * h;
  // the three planes of the tile, straight from the caller's picture into device memory
  const uint8_t *py = pic + (size_t)y0 * pic_w + x0, *pu = pic + pic_y + (size_t)(y0 / 2) * (pic_w / 2) + x0 / 2, *pv = pu + pic_y / 4;
  ENC_CHECK(cudaMemcpy2DAsync(s.d_src, w, py, pic_w, w, h, cudaMemcpyHostToDevice, upload_stream), "H2D strip Y");
  ENC_CHECK(cudaMemcpy2DAsync(s.d_src + ysz, w / 2, pu, pic_w / 2, w / 2, h / 2, cudaMemcpyHostToDevice, upload_stream), "H2D strip U");
  ENC_CHECK(cudaMemcpy2DAsync(s.d_src + ysz + ysz / 4, w / 2, pv, pic_w / 2, w / 2, h / 2, cudaMemcpyHostToDevice, upload_stream), "H2D strip V");
  ENC_CHECK(cudaEventRecord(ev_upload, upload_stream), "event record");
  ENC_CHECK(cudaStreamWaitEvent(input_stream(), ev_upload, 0), "stream wait");
  return encode_device(s.d_src, out);
}

bool Encoder::flush(std::vector<uint8_t> &out)
{
  out.clear();
  if (inflight.empty()) return true;
  int slot = inflight.front();
  inflight.pop_front();
  return collect(slots[slot], out);
}

}  // namespace b200

// ---- plain C entry points (benchmark / parity tests) ------------------------------------------

using b200::Encoder;
using b200::EncoderConfig;

extern "C" {

void *b200_enc_open(int width, int height, int qp, int intra_period, int search_range, int deblock, int debug, int depth)
{
  Encoder *e = new Encoder();
  EncoderConfig c;
  c.width = width; c.height = height; c.qp = qp; c.intra_period = intra_period; c.search_range = search_range;
  c.deblock = deblock; c.debug = debug; c.depth = depth;
  if (!e->open(c)) { delete e; return nullptr; }
  return e;
}

void *b200_enc_open_roi(int width, int height, int qp, int intra_period, int search_range, int deblock, int debug, int depth)
{
  Encoder *e = new Encoder();
  EncoderConfig c;
  c.width = width; c.height = height; c.qp = qp; c.intra_period = intra_period; c.search_range = search_range;
  c.deblock = deblock; c.debug = debug; c.depth = depth; c.qp_delta = 1;
  if (!e->open(c)) { delete e; return nullptr; }
  return e;
}

void b200_enc_params_default(b200_enc_params *p)
{
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->struct_size = (int)sizeof(*p);
  p->qp = 32; p->intra_period = 64; p->search_range = 8; p->deblock = 1; p->depth = 1; p->vps_period = 1;
}

int b200_enc_parameter_sets(const b200_enc_params *up, int tile_cols, int tile_rows, int wpp, uint8_t *out, int cap)
{
  if (!up || up->struct_size < (int)(9 * sizeof(int)) || !out || cap < 0 || tile_cols < 1 || tile_rows < 1) {
    b200::set_error("b200_enc_parameter_sets: bad arguments");
    return B200_ERR_ARG;
  }
  b200_enc_params p;
  b200_enc_params_default(&p);
  memcpy(&p, up, std::min<size_t>((size_t)up->struct_size, sizeof(p)));
  b200::StreamLayout l;
  l.w = p.width; l.h = p.height; l.deblock = p.deblock; l.qp_delta = p.qp_delta; l.fps_num = p.fps_num; l.fps_den = p.fps_den;
  l.sao = p.sao; l.tile_cols = tile_cols; l.tile_rows = tile_rows; l.wpp = wpp ? 1 : 0; l.scaling_list = p.scaling_list ? 1 : 0;
  l.conf_right = p.src_width ? p.width - p.src_width : 0; l.conf_bottom = p.src_height ? p.height - p.src_height : 0;
  if (l.w <= 0 || l.h <= 0 || (l.w & 7) || (l.h & 7) || l.conf_right < 0 || l.conf_right > 6 || l.conf_bottom < 0 || l.conf_bottom > 6 ||
      ((l.conf_right | l.conf_bottom) & 1)) {
    b200::set_error("b200_enc_parameter_sets: bad picture size");
    return B200_ERR_ARG;
  }
  std::vector<uint8_t> ps;
  b200::write_parameter_sets(l, ps);
  if ((size_t)cap < ps.size()) return -(int)ps.size();
  memcpy(out, ps.data(), ps.size());
  return (int)ps.size();
}

void *b200_enc_open_params(const b200_enc_params *up)
{
  if (!up || up->struct_size < (int)(9 * sizeof(int))) { b200::set_error("b200_enc_open_params: bad arguments"); return nullptr; }
  b200_enc_params p;
  b200_enc_params_default(&p);
  memcpy(&p, up, std::min<size_t>((size_t)up->struct_size, sizeof(p)));
  Encoder *e = new Encoder();
  EncoderConfig c;
  c.width = p.width; c.height = p.height; c.qp = p.qp; c.intra_period = p.intra_period; c.search_range = p.search_range;
  c.deblock = p.deblock; c.debug = p.debug; c.depth = p.depth; c.qp_delta = p.qp_delta;
  c.fps_num = p.fps_num; c.fps_den = p.fps_den; c.sao = p.sao; c.intra_in_p = p.intra_in_p; c.me_coarse = p.me_coarse;
  c.intra_satd = p.intra_satd; c.subme_satd = p.subme_satd; c.vaq = p.vaq; c.scaling_list = p.scaling_list ? 1 : 0;
  c.src_width = p.src_width; c.src_height = p.src_height; c.mv_edges = p.mv_edges & 15; c.vps_period = p.vps_period;
  if (!e->open(c)) { delete e; return nullptr; }
  return e;
}

int b200_enc_set_ctu_dqp(void *h, const int8_t *dqp, int n)
{
  Encoder *e = (Encoder *)h;
  if (!e) { b200::set_error("b200_enc_set_ctu_dqp: bad arguments"); return B200_ERR_ARG; }
  return e->set_ctu_dqp(dqp, n) ? B200_OK : B200_ERR_ARG;
}

void b200_enc_close(void *h) { delete (Encoder *)h; }

static int finish_au(Encoder *e, uint8_t *out, int cap)
{
  if (e->au.empty()) return 0;
  if ((size_t)cap < e->au.size()) { b200::set_error("b200_enc_encode: output buffer too small (%zu needed)", e->au.size()); return -(int)e->au.size(); }
  memcpy(out, e->au.data(), e->au.size());
  return (int)e->au.size();
}

int b200_enc_encode(void *h, const uint8_t *i420, uint8_t *out, int cap)
{
  Encoder *e = (Encoder *)h;
  if (!e || !i420 || !out) { b200::set_error("b200_enc_encode: bad arguments"); return B200_ERR_ARG; }
  if (!e->encode_host(i420, e->au)) return B200_ERR_CUDA;
  return finish_au(e, out, cap);
}

int b200_enc_encode_dev(void *h, const uint8_t *d_i420, uint8_t *out, int cap)
{
  Encoder *e = (Encoder *)h;
  if (!e || !d_i420 || !out) { b200::set_error("b200_enc_encode_dev: bad arguments"); return B200_ERR_ARG; }
  // order the caller's stream before ours: the caller must have synchronised its producer
  if (!e->encode_device(d_i420, e->au)) return B200_ERR_CUDA;
  return finish_au(e, out, cap);
}

int b200_enc_flush(void *h, uint8_t *out, int cap)
{
  Encoder *e = (Encoder *)h;
  if (!e || !out) { b200::set_error("b200_enc_flush: bad arguments"); return B200_ERR_ARG; }
  if (!e->flush(e->au)) return B200_ERR_CUDA;
  return finish_au(e, out, cap);
}

int b200_enc_pending(void *h) { return h ? ((Encoder *)h)->pending() : 0; }

// Per-kernel device time from CUDA events on the launching stream.  Kernel ids: 0 intra, 1 me,
// 2 inter recon, 3 modes, 4 deblock (both passes), 5 binarise, 6 arithmetic coder, 7 pack.
int b200_enc_set_profile(void *h, int on)
{
  Encoder *e = (Encoder *)h;
  if (!e) return B200_ERR_ARG;
  e->profile = on;
  for (int k = 0; k < Encoder::K_COUNT; k++) { e->prof_ms[k] = 0; e->prof_cnt[k] = 0; }
  if (on) { cudaStreamSynchronize(e->stream); cudaMemset(e->d_me_stats, 0, 4 * sizeof(unsigned long long)); }
  return B200_OK;
}

// Work counters of the motion search since b200_enc_set_profile(enc, 1): out[0] CTUs, [1] 32x32
// quadrants whose second centre set was searched, [2] 16x16 intra mode searches, [3] intra CUs chosen.
int b200_enc_get_me_stats(void *h, unsigned long long *out, int n)
{
  Encoder *e = (Encoder *)h;
  if (!e || !out || n < 4) return B200_ERR_ARG;
  B200_CHECK(cudaStreamSynchronize(e->stream), "me stats sync");
  B200_CHECK(cudaMemcpy(out, e->d_me_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost), "me stats read");
  return 4;
}

// begin/end times (ms since the encoder was opened) of each kernel of the last collected picture
int b200_enc_get_timeline(void *h, float *out, int n)
{
  Encoder *e = (Encoder *)h;
  if (!e || !out) return B200_ERR_ARG;
  int k = 0;
  for (; k < n && k < 2 * Encoder::K_COUNT; k++) out[k] = e->timeline[k];
  return k;
}

int b200_enc_get_profile(void *h, double *ms, unsigned long long *count, int n)
{
  Encoder *e = (Encoder *)h;
  if (!e || !ms || !count) return B200_ERR_ARG;
  int k = 0;
  for (; k < n && k < Encoder::K_COUNT; k++) { ms[k] = e->prof_ms[k]; count[k] = e->prof_cnt[k]; }
  return k;
}

// what: 0 reconstruction (after deblocking), 1 reconstruction before deblocking (debug=1 only),
//       2 cu map, 3 levels -- of the most recently submitted picture (waits for it)
int b200_enc_debug_read(void *h, int what, void *dst, size_t bytes)
{
  Encoder *e = (Encoder *)h;
  if (!e || !dst) return B200_ERR_ARG;
  const b200::FrameSlot &s = e->slots[e->last_slot];
  B200_CHECK(cudaStreamSynchronize(e->stream), "debug sync");
  B200_CHECK(cudaStreamSynchronize(s.stream), "debug sync");
  const void *src = nullptr;
  size_t n = 0;
  switch (what) {
  case 0: src = e->last_rec(); n = e->frame_bytes; break;
  case 1: src = e->d_rec_pre; n = e->frame_bytes; break;
  case 2: src = s.d_cu; n = sizeof(b200::CuInfo) * e->fp.w8 * e->fp.h8; break;
  case 3: src = s.d_levels; n = e->frame_bytes * sizeof(int16_t); break;
  default: return B200_ERR_ARG;
  }
  if (bytes < n || !src) return B200_ERR_ARG;
  B200_CHECK(cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost), "debug read");
  return (int)B200_OK;
}

// Replace the reference picture the next P frame will predict from (test hook: lets a single
// kernel stage be compared against the oracle without depending on earlier stages).
int b200_enc_debug_set_reference(void *h, const uint8_t *i420)
{
  Encoder *e = (Encoder *)h;
  if (!e || !i420) return B200_ERR_ARG;
  B200_CHECK(cudaStreamSynchronize(e->stream), "debug sync");
  B200_CHECK(cudaMemcpy(e->last_rec(), i420, e->frame_bytes, cudaMemcpyHostToDevice), "set reference");
  return B200_OK;
}

unsigned long long b200_enc_last_bins(void *h)
{
  Encoder *e = (Encoder *)h;
  unsigned long long v = 0;
  if (!e) return 0;
  const b200::FrameSlot &s = e->slots[e->last_slot];
  cudaStreamSynchronize(s.stream);
  cudaMemcpy(&v, s.d_small + e->off_bins, sizeof(v), cudaMemcpyDeviceToHost);
  return v;
}

int b200_enc_last_was_idr(void *h) { return h ? ((Encoder *)h)->last_idr : 0; }

}  // extern "C"
