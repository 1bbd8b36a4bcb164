// Host side of the B200 HEVC encoder: owns the HBM-resident state of one stream, sequences the
// per-picture kernels on one CUDA stream, writes the parameter sets / slice header and assembles
// the Annex-B access unit from the per-row substreams the CABAC kernel produced.
//
// This is the engine under the kvz_api C ABI (kvz_api.cu); the b200_enc_* entry points below are
// the same engine exposed with plain arguments for the benchmark and the parity tests.
#include "hevc_encoder.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>

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
  if (d_src) cudaFree(d_src);
  for (int i = 0; i < 2; i++) if (d_rec[i]) cudaFree(d_rec[i]);
  if (d_rec_pre) cudaFree(d_rec_pre);
  if (d_cu) cudaFree(d_cu);
  if (d_levels) cudaFree(d_levels);
  if (d_rows) cudaFree(d_rows);
  if (d_small) cudaFree(d_small);
  if (h_src) cudaFreeHost(h_src);
  if (h_rows) cudaFreeHost(h_rows);
  if (h_small) cudaFreeHost(h_small);
  if (stream) cudaStreamDestroy(stream);
  d_src = d_rec[0] = d_rec[1] = d_rec_pre = d_rows = nullptr;
  d_cu = nullptr; d_levels = nullptr; d_small = nullptr; h_src = h_rows = nullptr; h_small = nullptr; stream = nullptr;
}

bool Encoder::open(const EncoderConfig &c)
{
  if (c.width <= 0 || c.height <= 0 || (c.width & 7) || (c.height & 7)) { set_error("encoder: width/height must be positive multiples of 8 (got %dx%d)", c.width, c.height); return false; }
  if (c.qp < 0 || c.qp > 51) { set_error("encoder: qp %d out of range 0..51", c.qp); return false; }
  if (c.search_range < 1 || c.search_range > 32) { set_error("encoder: search range %d out of range 1..32", c.search_range); return false; }
  if (b200_device_count() <= 0) { set_error("no CUDA device: the B200 encoder has no CPU fallback"); return false; }
  cfg = c;
  fp.w = c.width; fp.h = c.height; fp.w8 = c.width / 8; fp.h8 = c.height / 8;
  fp.ctb_cols = (c.width + kCtb - 1) / kCtb; fp.ctb_rows = (c.height + kCtb - 1) / kCtb;
  fp.qp = c.qp; fp.qp_c = kChromaQp[std::min(std::max(c.qp, 0), 57)];
  fp.lambda_q4 = kLambdaQ4[c.qp];
  fp.search_range = c.search_range; fp.is_idr = 1; fp.deblock = c.deblock;
  frame_bytes = (size_t)fp.w * fp.h * 3 / 2;
  row_cap = (uint32_t)fp.w * kCtb * 4 + 4096;
  ENC_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
  ENC_CHECK(cudaMalloc((void **)&d_src, frame_bytes), "cudaMalloc src");
  ENC_CHECK(cudaMalloc((void **)&d_rec[0], frame_bytes), "cudaMalloc rec0");
  ENC_CHECK(cudaMalloc((void **)&d_rec[1], frame_bytes), "cudaMalloc rec1");
  ENC_CHECK(cudaMalloc((void **)&d_rec_pre, frame_bytes), "cudaMalloc rec_pre");
  ENC_CHECK(cudaMalloc((void **)&d_cu, sizeof(CuInfo) * fp.w8 * fp.h8), "cudaMalloc cu");
  ENC_CHECK(cudaMalloc((void **)&d_levels, frame_bytes * sizeof(int16_t)), "cudaMalloc levels");
  ENC_CHECK(cudaMalloc((void **)&d_rows, (size_t)row_cap * fp.ctb_rows), "cudaMalloc rows");
  // small state: row_len[rows] | sync_flag[rows] | progress[rows] | ticket | bins(8) | sync_ctx[rows*CTX_COUNT]
  off_flag = sizeof(int) * fp.ctb_rows; off_prog = 2 * off_flag; off_ticket = 3 * off_flag;
  off_bins = (off_ticket + sizeof(int) + 7) & ~(size_t)7; off_ctx = off_bins + 8;
  small_bytes = off_ctx + (size_t)fp.ctb_rows * CTX_COUNT;
  ENC_CHECK(cudaMalloc((void **)&d_small, small_bytes), "cudaMalloc small");
  ENC_CHECK(cudaMemset(d_small, 0, small_bytes), "memset small");
  ENC_CHECK(cudaMemset(d_cu, 0, sizeof(CuInfo) * fp.w8 * fp.h8), "memset cu");
  ENC_CHECK(cudaMallocHost((void **)&h_src, frame_bytes), "cudaMallocHost src");
  ENC_CHECK(cudaMallocHost((void **)&h_rows, (size_t)row_cap * fp.ctb_rows), "cudaMallocHost rows");
  ENC_CHECK(cudaMallocHost((void **)&h_small, sizeof(uint32_t) * (fp.ctb_rows + 4)), "cudaMallocHost small");
  frame_idx = 0; poc = 0; cur = 0;
  return true;
}

void Encoder::write_parameter_sets(std::vector<uint8_t> &out) const
{
  const int level = level_for(fp.w, fp.h);
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
    b.ue((uint32_t)fp.w); b.ue((uint32_t)fp.h);
    b.put(0, 1);                             // conformance_window_flag
    b.ue(0); b.ue(0);                        // bit depths
    b.ue(4);                                 // log2_max_pic_order_cnt_lsb_minus4
    b.put(0, 1); b.ue(1); b.ue(0); b.ue(0);  // sub-layer ordering info
    b.ue(0); b.ue(3);                        // CB 8..64
    b.ue(0); b.ue(3);                        // TB 4..32
    b.ue(0); b.ue(0);                        // max_transform_hierarchy_depth inter / intra
    b.put(0, 1); b.put(0, 1); b.put(0, 1); b.put(0, 1);   // scaling lists, AMP, SAO, PCM off
    b.ue(1);                                 // one short-term RPS: the previous picture
    b.ue(1); b.ue(0); b.ue(0); b.put(1, 1);
    b.put(0, 1); b.put(0, 1); b.put(0, 1);   // long-term refs, TMVP, strong intra smoothing off
    b.put(0, 1); b.put(0, 1);                // VUI, extension
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
    b.put(0, 1); b.put(0, 1); b.put(0, 1);   // constrained intra, transform skip, cu_qp_delta
    b.se(0); b.se(0);
    b.put(0, 1); b.put(0, 1); b.put(0, 1); b.put(0, 1);
    b.put(0, 1);                             // tiles_enabled_flag
    b.put(1, 1);                             // entropy_coding_sync_enabled_flag
    b.put(1, 1);                             // pps_loop_filter_across_slices_enabled_flag
    if (cfg.deblock) {
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

void Encoder::write_slice(std::vector<uint8_t> &out, bool idr, const uint32_t *row_len) const
{
  const int rows = fp.ctb_rows;
  BitWriter b;
  b.put(1, 1);
  if (idr) b.put(0, 1);
  b.ue(0);
  b.ue(idr ? 2 : 1);
  if (!idr) {
    b.put((uint32_t)(poc & 255), 8);
    b.put(1, 1);
    b.put(0, 1);
    b.ue(5 - kMaxMerge);
  }
  b.se(cfg.qp - 26);
  if (cfg.deblock) b.put(1, 1);
  b.ue((uint32_t)(rows - 1));
  if (rows > 1) {
    uint32_t mx = 1;
    for (int r = 0; r < rows - 1; r++) mx = std::max(mx, row_len[r]);
    int len = 1;
    while (((mx - 1) >> len) > 0) len++;
    b.ue((uint32_t)(len - 1));
    for (int r = 0; r < rows - 1; r++) b.put(row_len[r] - 1, len);
  }
  b.trailing();
  start_nal(out, idr ? 19 : 1);
  append_escaped(out, b.bytes.data(), b.bytes.size());
  for (int r = 0; r < rows; r++) out.insert(out.end(), h_rows + (size_t)r * row_cap, h_rows + (size_t)r * row_cap + row_len[r]);
}

bool Encoder::encode_device(const uint8_t *d_i420, std::vector<uint8_t> &au)
{
  const bool idr = frame_idx == 0 || (cfg.intra_period > 0 && frame_idx % cfg.intra_period == 0);
  if (idr) poc = 0;
  fp.is_idr = idr ? 1 : 0;
  uint8_t *rec = d_rec[cur], *ref = d_rec[cur ^ 1];
  ENC_CHECK(cudaMemsetAsync(d_levels, 0, frame_bytes * sizeof(int16_t), stream), "memset levels");
  if (idr) {
    ENC_CHECK(launch_intra_frame(fp, d_i420, rec, d_levels, d_cu, d_progress(), d_ticket(), stream), "intra launch");
    count_launch(1);
  } else {
    ENC_CHECK(launch_inter_me(fp, d_i420, ref, d_cu, stream), "me launch");
    ENC_CHECK(launch_inter_recon(fp, d_i420, ref, rec, d_levels, d_cu, stream), "recon launch");
    ENC_CHECK(launch_inter_modes(fp, d_cu, stream), "modes launch");
    count_launch(3);
  }
  if (cfg.debug) ENC_CHECK(cudaMemcpyAsync(d_rec_pre, rec, frame_bytes, cudaMemcpyDeviceToDevice, stream), "copy pre-deblock");
  if (cfg.deblock) {
    ENC_CHECK(launch_deblock(fp, rec, d_cu, stream), "deblock launch");
    count_launch(2);
  }
  ENC_CHECK(launch_cabac(fp, d_cu, d_levels, d_rows, row_cap, d_row_len(), d_sync_ctx(), d_sync_flag(), d_bins(), stream),
            "cabac launch");
  count_launch(1);
  ENC_CHECK(cudaMemcpyAsync(h_small, d_row_len(), sizeof(uint32_t) * fp.ctb_rows, cudaMemcpyDeviceToHost, stream), "D2H row_len");
  ENC_CHECK(cudaStreamSynchronize(stream), "sync after cabac");
  for (int r = 0; r < fp.ctb_rows; r++) {
    if (h_small[r] == 0xffffffffu || h_small[r] == 0) { set_error("encoder: substream %d overflowed its %u-byte buffer", r, row_cap); return false; }
    ENC_CHECK(cudaMemcpyAsync(h_rows + (size_t)r * row_cap, d_rows + (size_t)r * row_cap, h_small[r], cudaMemcpyDeviceToHost, stream), "D2H row");
  }
  ENC_CHECK(cudaStreamSynchronize(stream), "sync rows");
  au.clear();
  if (idr) write_parameter_sets(au);
  write_slice(au, idr, h_small);
  last_idr = idr;
  cur ^= 1;
  frame_idx++;
  poc++;
  return true;
}

bool Encoder::encode_host(const uint8_t *i420, std::vector<uint8_t> &au)
{
  memcpy(h_src, i420, frame_bytes);
  ENC_CHECK(cudaMemcpyAsync(d_src, h_src, frame_bytes, cudaMemcpyHostToDevice, stream), "H2D frame");
  return encode_device(d_src, au);
}

}  // namespace b200

// ---- plain C entry points (benchmark / parity tests) ------------------------------------------

using b200::Encoder;
using b200::EncoderConfig;

extern "C" {

void *b200_enc_open(int width, int height, int qp, int intra_period, int search_range, int deblock, int debug)
{
  Encoder *e = new Encoder();
  EncoderConfig c;
  c.width = width; c.height = height; c.qp = qp; c.intra_period = intra_period; c.search_range = search_range;
  c.deblock = deblock; c.debug = debug;
  if (!e->open(c)) { delete e; return nullptr; }
  return e;
}

void b200_enc_close(void *h) { delete (Encoder *)h; }

static int finish_au(Encoder *e, uint8_t *out, int cap)
{
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

// what: 0 reconstruction (after deblocking), 1 reconstruction before deblocking (debug=1 only),
//       2 cu map, 3 levels, 4 reference picture the NEXT frame will use (== 0)
int b200_enc_debug_read(void *h, int what, void *dst, size_t bytes)
{
  Encoder *e = (Encoder *)h;
  if (!e || !dst) return B200_ERR_ARG;
  const void *src = nullptr;
  size_t n = 0;
  switch (what) {
  case 0: case 4: src = e->d_rec[e->cur ^ 1]; n = e->frame_bytes; break;
  case 1: src = e->d_rec_pre; n = e->frame_bytes; break;
  case 2: src = e->d_cu; n = sizeof(b200::CuInfo) * e->fp.w8 * e->fp.h8; break;
  case 3: src = e->d_levels; n = e->frame_bytes * sizeof(int16_t); break;
  default: return B200_ERR_ARG;
  }
  if (bytes < n) return B200_ERR_ARG;
  B200_CHECK(cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost), "debug read");
  return (int)B200_OK;
}

// Replace the reference picture the next P frame will predict from (test hook: lets a single
// kernel stage be compared against the oracle without depending on earlier stages).
int b200_enc_debug_set_reference(void *h, const uint8_t *i420)
{
  Encoder *e = (Encoder *)h;
  if (!e || !i420) return B200_ERR_ARG;
  B200_CHECK(cudaMemcpy(e->d_rec[e->cur ^ 1], i420, e->frame_bytes, cudaMemcpyHostToDevice), "set reference");
  return B200_OK;
}

unsigned long long b200_enc_last_bins(void *h)
{
  Encoder *e = (Encoder *)h;
  unsigned long long v = 0;
  if (e) cudaMemcpy(&v, e->d_bins(), sizeof(v), cudaMemcpyDeviceToHost);
  return v;
}

int b200_enc_last_was_idr(void *h) { return h ? ((Encoder *)h)->last_idr : 0; }

}  // extern "C"
