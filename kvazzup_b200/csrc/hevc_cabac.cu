// CABAC entropy coding of the B200 HEVC encoder (sm_100a); SURVEY.md 8a-K row K8.
//
// Only the range-coder update (low, range) is inherently serial per substream; binarisation and
// even the evolution of the context states are not.  So entropy coding is three kernels that meet in
// a record buffer in HBM (hevc_common.h kRecUnitCap):
//
//   k_binarise     one warp per CTU quadrant, every CU of the picture at once.  Writes each CU's
//                  complete syntax as a list of bin RECORDS (context index + value, a group of
//                  <= 16 bypass bins).  All context indices of HEVC's coding_unit / residual_coding
//                  syntax are functions of the cu map and the levels only -- never of the coder's
//                  state -- so the whole picture binarises in parallel; inside a transform block
//                  every lane binarises whole 4x4 sub-blocks.
//   k_entropy_rows two warps per WPP substream (CTU row), one launch:
//     ctx_rows     context-state resolution, 32 records at a time (see the comment above it); owns the
//                  WPP context hand-over; publishes the number of CTUs it has resolved.
//     arith_rows   the range coder over resolved records, one CTU behind ctx_rows, no dependency
//                  between rows; every lane runs it redundantly, lane 0 stores the bytes, escaped
//                  on the fly (exact: each substream starts after a non-zero byte).
#include "hevc_device.cuh"
#include "hevc_kernels.h"

namespace b200 {

namespace {

struct Coder {
  uint32_t low, range;
  int bits_left, num_buffered, buffered_byte;
  uint8_t *out;
  uint32_t pos, cap;
  int zeros;
  unsigned bins;              // of this substream (added to the picture's 64-bit total at the end)
  bool writer;                // lane 0
};

__device__ __forceinline__ void emit(Coder &c, uint32_t byte)
{
  byte &= 0xff;
  if (c.zeros >= 2 && byte <= 3) {
    if (c.writer && c.pos < c.cap) c.out[c.pos] = 3;
    c.pos++;
    c.zeros = 0;
  }
  if (c.writer && c.pos < c.cap) c.out[c.pos] = (uint8_t)byte;
  c.pos++;
  c.zeros = byte == 0 ? c.zeros + 1 : 0;
}

__device__ __forceinline__ void coder_start(Coder &c)
{
  c.low = 0; c.range = 510; c.bits_left = 23; c.num_buffered = 0; c.buffered_byte = 0xff;
}

__device__ __forceinline__ void write_out(Coder &c)
{
  uint32_t lead = c.low >> (24 - c.bits_left);
  c.bits_left += 8;
  c.low &= 0xffffffffu >> c.bits_left;
  if (lead == 0xff) {
    c.num_buffered++;
  } else if (c.num_buffered > 0) {
    uint32_t carry = lead >> 8;
    uint32_t byte = (uint32_t)c.buffered_byte + carry;
    c.buffered_byte = (int)(lead & 0xff);
    emit(c, byte);
    byte = (0xff + carry) & 0xff;
    while (c.num_buffered > 1) { emit(c, byte); c.num_buffered--; }
  } else {
    c.num_buffered = 1;
    c.buffered_byte = (int)lead;
  }
}

// n <= 16 bypass bins at once (HM encodeBinsEP): low = low * 2^n + range * value, eight at a time
__device__ __forceinline__ void enc_bypass_group(Coder &c, uint32_t v, int n)
{
  c.bins += n;
  if (n > 8) {
    n -= 8;
    uint32_t pat = v >> n;
    c.low = (c.low << 8) + c.range * pat;
    v -= pat << n;
    c.bits_left -= 8;
    if (c.bits_left < 12) write_out(c);
  }
  c.low = (c.low << n) + c.range * v;
  c.bits_left -= n;
  if (c.bits_left < 12) write_out(c);
}

__device__ __forceinline__ void enc_terminate(Coder &c, int bin)
{
  c.bins++;
  c.range -= 2;
  if (bin) {
    c.low += c.range;
    c.low <<= 7;
    c.range = 2 << 7;
    c.bits_left -= 7;
  } else if (c.range >= 256) {
    return;
  } else {
    c.low <<= 1;
    c.range <<= 1;
    c.bits_left--;
  }
  if (c.bits_left < 12) write_out(c);
}

// flush (9.3.4.5) + rbsp_stop_one_bit / alignment bit + zero bits up to the byte boundary
__device__ __forceinline__ void coder_finish(Coder &c)
{
  if (c.low >> (32 - c.bits_left)) {
    emit(c, (uint32_t)(c.buffered_byte + 1));
    while (c.num_buffered > 1) { emit(c, 0x00); c.num_buffered--; }
    c.low -= 1u << (32 - c.bits_left);
  } else {
    if (c.num_buffered > 0) emit(c, (uint32_t)c.buffered_byte);
    while (c.num_buffered > 1) { emit(c, 0xff); c.num_buffered--; }
  }
  int nb = 24 - c.bits_left;                      // 1..12 bits of (low >> 8), then the stop bit
  uint32_t v = (((c.low >> 8) & ((1u << nb) - 1)) << 1) | 1u;
  nb += 1;
  int pad = (8 - (nb & 7)) & 7;
  v <<= pad; nb += pad;
  for (int i = nb - 8; i >= 0; i -= 8) emit(c, v >> i);
}

// Every lane keeps a private copy of the context table (entry i of lane l at s_ctx[i*32 + l]):
// the lanes run the same coder redundantly, and private tables make that independent of warp
// reconvergence timing.
// ---- bin records ------------------------------------------------------------------------------------
//   bit31 = 0, bit30 = 0 : context-coded bin,  bits 15..1 context index, bit 0 value
//   bit31 = 1            : bypass group,       bits 28..24 count (1..16), bits 15..0 value

struct Syn { uint32_t *r; int k; };      // syntax-element record list (shared memory; lanes write identical values)

__device__ __forceinline__ void put_ctx(Syn &s, int ctx, int bin) { s.r[s.k++] = ((uint32_t)ctx << 1) | (uint32_t)(bin & 1); }
__device__ __forceinline__ void put_byp(uint32_t *rec, int &k, unsigned long long bits, int n)
{
  while (n > 0) {
    int m = n > 16 ? 16 : n;
    uint32_t v = (uint32_t)((bits >> (n - m)) & ((1u << m) - 1));
    rec[k++] = 0x80000000u | ((uint32_t)m << 24) | v;
    n -= m;
  }
}
__device__ __forceinline__ void put_byp(Syn &s, unsigned long long bits, int n) { put_byp(s.r, s.k, bits, n); }

// ---- residual_coding (7.3.8.11) ---------------------------------------------------------------

__device__ __forceinline__ void scan_pos(int scan_idx, int blk_log2, int i, int &x, int &y)
{
  int n = 1 << blk_log2;
  if (scan_idx == 1) { x = i & (n - 1); y = i >> blk_log2; return; }
  if (scan_idx == 2) { x = i >> blk_log2; y = i & (n - 1); return; }
  int v = blk_log2 == 2 ? c_diag4[i] : (blk_log2 == 1 ? c_diag2[i] : (blk_log2 == 3 ? c_diag8[i] : 0));
  x = v & 15; y = v >> 4;
}

__device__ __forceinline__ void put_last_prefix(Syn &s, int pos, int log2n, int cidx, int base)
{
  int offset, shift;
  if (cidx == 0) { offset = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; }
  else { offset = 15; shift = log2n - 2; }
  int prefix = pos < 4 ? pos : ((pos < 8) ? 4 + ((pos - 4) >> 1) : (pos < 16 ? 6 + ((pos - 8) >> 2) : 8 + ((pos - 16) >> 3)));
  int cmax = (log2n << 1) - 1;
  for (int i = 0; i < prefix; i++) put_ctx(s, base + offset + (i >> shift), 1);
  if (prefix < cmax) put_ctx(s, base + offset + (prefix >> shift), 0);
}
__device__ __forceinline__ void put_last_suffix(Syn &s, int pos)
{
  if (pos < 4) return;
  int g = (pos < 8) ? 4 + ((pos - 4) >> 1) : (pos < 16 ? 6 + ((pos - 8) >> 2) : 8 + ((pos - 16) >> 3));
  int nb = (g >> 1) - 1;
  int min_in_group = (2 + (g & 1)) << nb;
  put_byp(s, (unsigned)(pos - min_in_group), nb);
}

// Phase A (parallel): every lane binarises whole 4x4 sub-blocks.  All context indices of
// residual_coding are pure functions of the levels (the greater1 context set depends on the
// previous coded sub-block only through "did it contain a level > 1 among its first eight"),
// so each sub-block's bins are produced independently.  Phase B (serial, lane-redundant): the
// arithmetic coder walks the record lists in coding order.
constexpr int kRecSlot = 80;          // records per sub-block: 1 + 16 + 8 + 1 + 1 + up to 3 per remaining level

struct ResidualShared {
  uint32_t rec[64][kRecSlot];
  uint8_t cnt[64];
  uint8_t any[64], hasg1[64], lastp[64];
  uint8_t csbf_pos[64];               // coded_sub_block_flag by position (ys * 8 + xs)
};

// Returns the index of the last sub-block (-1: block is all zero); the last-position syntax is
// appended to `syn`, the per-sub-block records are left in rs.rec / rs.cnt.
__device__ int binarise_residual(Syn &syn, ResidualShared &rs, const int16_t *lv, int log2n, int cidx, int scan_idx, int lane)
{
  const int n = 1 << log2n, sb_log2 = log2n - 2, nsb = 1 << (2 * sb_log2), sbw = 1 << sb_log2;
  for (int i = lane; i < nsb; i += 32) {
    int xs, ys;
    scan_pos(scan_idx, sb_log2, i, xs, ys);
    int lastp = -1, g1seen = 0, nsig = 0;
    for (int p = 15; p >= 0; p--) {
      int xp, yp;
      scan_pos(scan_idx, 2, p, xp, yp);
      int v = lv[(ys * 4 + yp) * n + xs * 4 + xp];
      if (v) {
        if (lastp < 0) lastp = p;
        if (nsig < 8 && (v > 1 || v < -1)) g1seen = 1;
        nsig++;
      }
    }
    rs.any[i] = nsig != 0; rs.hasg1[i] = (uint8_t)g1seen; rs.lastp[i] = (uint8_t)(lastp < 0 ? 0 : lastp);
  }
  __syncwarp();
  // last sub-block in scan order: ballot over the (at most two) sub-blocks each lane owns
  unsigned lo = __ballot_sync(0xffffffffu, lane < nsb && rs.any[lane]);
  unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < nsb && rs.any[lane + 32]);
  int last_sb = hi ? 63 - __clz(hi) : (lo ? 31 - __clz(lo) : -1);
  if (last_sb < 0) return -1;
  const int last_pos = rs.lastp[last_sb];
  for (int i = lane; i < nsb; i += 32) {
    int xs, ys;
    scan_pos(scan_idx, sb_log2, i, xs, ys);
    rs.csbf_pos[ys * 8 + xs] = i > last_sb ? 0 : ((i == last_sb || i == 0) ? 1 : rs.any[i]);
  }
  __syncwarp();
  for (int i = lane; i <= last_sb; i += 32) {
    uint32_t *rec = rs.rec[i];
    int k = 0;
    int xs, ys;
    scan_pos(scan_idx, sb_log2, i, xs, ys);
    int right = xs + 1 < sbw ? rs.csbf_pos[ys * 8 + xs + 1] : 0;
    int below = ys + 1 < sbw ? rs.csbf_pos[(ys + 1) * 8 + xs] : 0;
    int prev_csbf = right | (below << 1);
    int absv[16];
    unsigned sig = 0, neg = 0;
#pragma unroll 1
    for (int p = 0; p < 16; p++) {
      int xp, yp;
      scan_pos(scan_idx, 2, p, xp, yp);
      int v = lv[(ys * 4 + yp) * n + xs * 4 + xp];
      if (i == last_sb && p > last_pos) v = 0;
      absv[p] = abs(v);
      if (v) sig |= 1u << p;
      if (v < 0) neg |= 1u << p;
    }
    int any = sig != 0, infer_dc = 0;
    if (i < last_sb && i > 0) {
      rec[k++] = (uint32_t)((CTX_CSBF + (prev_csbf ? 1 : 0) + (cidx ? 2 : 0)) << 1) | (uint32_t)any;
      infer_dc = 1;
    } else {
      any = 1;
    }
    if (any) {
      int start = (i == last_sb) ? last_pos - 1 : 15;
      for (int p = start; p >= 0; p--) {
        if (p == 0 && infer_dc) break;
        int xp, yp;
        scan_pos(scan_idx, 2, p, xp, yp);
        int xc = xs * 4 + xp, yc = ys * 4 + yp, sctx;
        if (log2n == 2) sctx = c_sig_ctx_4x4[(yc << 2) + xc];
        else if (xc + yc == 0) sctx = 0;
        else {
          if (prev_csbf == 0) sctx = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
          else if (prev_csbf == 1) sctx = yp == 0 ? 2 : (yp == 1 ? 1 : 0);
          else if (prev_csbf == 2) sctx = xp == 0 ? 2 : (xp == 1 ? 1 : 0);
          else sctx = 2;
          if (cidx == 0) {
            if (xs || ys) sctx += 3;
            sctx += log2n == 3 ? (scan_idx == 0 ? 9 : 15) : 21;
          } else {
            sctx += log2n == 3 ? 9 : 12;
          }
        }
        int sbit = (sig >> p) & 1;
        rec[k++] = (uint32_t)((CTX_SIG + (cidx == 0 ? sctx : 27 + sctx)) << 1) | (uint32_t)sbit;
        if (sbit) infer_dc = 0;
      }
      // greater1 context set: carried from the previous sub-block that had coefficients
      int ctx_set = (i > 0 && cidx == 0) ? 2 : 0;
      for (int j = i + 1; j <= last_sb; j++)
        if (rs.any[j]) { if (rs.hasg1[j]) ctx_set++; break; }
      int c1 = 1, num_g1 = 0, first_g1 = -1;
      unsigned g1 = 0;
      for (int p = 15; p >= 0; p--) {
        if (!((sig >> p) & 1) || num_g1 >= 8) continue;
        int f = absv[p] > 1;
        rec[k++] = (uint32_t)((CTX_GT1 + (cidx ? 16 : 0) + 4 * ctx_set + c1) << 1) | (uint32_t)f;
        if (f) { g1 |= 1u << p; c1 = 0; if (first_g1 < 0) first_g1 = p; }
        else if (c1 < 3 && c1 > 0) c1++;
        num_g1++;
      }
      int g2 = 0;
      if (first_g1 >= 0) {
        g2 = absv[first_g1] > 2;
        rec[k++] = (uint32_t)((CTX_GT2 + (cidx ? 4 : 0) + ctx_set) << 1) | (uint32_t)g2;
      }
      {
        unsigned long long sb = 0;
        int ns = 0;
        for (int p = 15; p >= 0; p--)
          if ((sig >> p) & 1) { sb = (sb << 1) | ((neg >> p) & 1); ns++; }
        put_byp(rec, k, sb, ns);
      }
      int num_sig = 0, rice = 0;
      for (int p = 15; p >= 0; p--) {
        if (!((sig >> p) & 1)) continue;
        int base = 1 + (num_sig < 8 ? (int)((g1 >> p) & 1) : 0) + (p == first_g1 ? g2 : 0);
        int thresh = num_sig < 8 ? (p == first_g1 ? 3 : 2) : 1;
        if (base == thresh) {
          int value = absv[p] - base;
          if (value < (3 << rice)) {
            int len = value >> rice;
            unsigned long long bits = ((((1ull << (len + 1)) - 2) << rice) | (unsigned)(value & ((1 << rice) - 1)));
            put_byp(rec, k, bits, len + 1 + rice);
          } else {
            int len = rice, vv = value - (3 << rice);
            while (vv >= (1 << len)) { vv -= 1 << len; len++; }
            int pre = 3 + len + 1 - rice;
            unsigned long long bits = ((((1ull << pre) - 2) << len) | (unsigned)vv);
            put_byp(rec, k, bits, pre + len);
          }
          if (absv[p] > (3 << rice)) rice = min(rice + 1, 4);
        }
        num_sig++;
      }
    }
    rs.cnt[i] = (uint8_t)k;
  }
  {
    int xs, ys, xp, yp;
    scan_pos(scan_idx, sb_log2, last_sb, xs, ys);
    scan_pos(scan_idx, 2, last_pos, xp, yp);
    int px = xs * 4 + xp, py = ys * 4 + yp;
    if (scan_idx == 2) { int tt = px; px = py; py = tt; }
    put_last_prefix(syn, px, log2n, cidx, CTX_LAST_X);
    put_last_prefix(syn, py, log2n, cidx, CTX_LAST_Y);
    put_last_suffix(syn, px);
    put_last_suffix(syn, py);
  }
  __syncwarp();
  return last_sb;
}

// ---- coding quadtree ------------------------------------------------------------------------------

__device__ __forceinline__ unsigned coding_order_c(const FrameParams &fp, int x, int y)
{
  return (unsigned)((y >> kCtbLog2) * fp.ctb_cols + (x >> kCtbLog2)) * 64u + (unsigned)xy_to_z((x >> 3) & 7, (y >> 3) & 7);
}

// cu-map access for the binariser (straight from HBM / L2: thousands of warps hide the latency)
struct CuView {
  const FrameParams *fp;
  const CuInfo *cu;
};
__device__ __forceinline__ const CuInfo &cu_at(const CuView &v, int x, int y) { return v.cu[(size_t)(y >> 3) * v.fp->w8 + (x >> 3)]; }

struct NbMv { bool ok; int mvx, mvy; };
__device__ __forceinline__ NbMv nb_mv(const CuView &v, unsigned cur, int xn, int yn)
{
  const FrameParams &fp = *v.fp;
  NbMv n{false, 0, 0};
  if (xn < 0 || yn < 0 || xn >= fp.w || yn >= fp.h) return n;
  if (coding_order_c(fp, xn, yn) >= cur) return n;
  const CuInfo *c = &cu_at(v, xn, yn);
  if (c->pred_mode != 0) return n;
  n.ok = true; n.mvx = c->mvx; n.mvy = c->mvy;
  return n;
}

__device__ __forceinline__ int scan_idx_for(int pred_mode, int intra_mode, int log2n, int cidx)
{
  if (pred_mode != 1) return 0;
  if (!(log2n == 2 || (log2n == 3 && cidx == 0))) return 0;
  if (intra_mode >= 6 && intra_mode <= 14) return 2;
  if (intra_mode >= 22 && intra_mode <= 30) return 1;
  return 0;
}

__device__ __forceinline__ void put_mvd(Syn &s, int dx, int dy)
{
  int ax = abs(dx), ay = abs(dy);
  put_ctx(s, CTX_MVD_GT0, ax > 0);
  put_ctx(s, CTX_MVD_GT0, ay > 0);
  if (ax) put_ctx(s, CTX_MVD_GT1, ax > 1);
  if (ay) put_ctx(s, CTX_MVD_GT1, ay > 1);
  for (int k = 0; k < 2; k++) {
    int a = k ? ay : ax, d = k ? dy : dx;
    if (!a) continue;
    if (a > 1) {                                   // abs_mvd_minus2: EG1
      int v = a - 2, kk = 1, ones = 0;
      while (v >= (1 << kk)) { ones++; v -= 1 << kk; kk++; }
      unsigned long long bits = ((((1ull << (ones + 1)) - 2) << kk) | (unsigned)v);
      put_byp(s, bits, ones + 1 + kk);
    }
    put_byp(s, d < 0, 1);
  }
}

// split_cu_flag (7.3.8.4): coded when the block fits the picture and is larger than the minimum CU
__device__ __forceinline__ void put_split(Syn &s, const CuView &v, int x0, int y0, int log2, int depth, int split)
{
  const FrameParams &fp = *v.fp;
  const int n = 1 << log2;
  if (!(x0 + n <= fp.w && y0 + n <= fp.h && log2 > 3)) return;       // inferred
  int ctx = 0;
  if (x0 > 0) ctx += (kCtbLog2 - cu_at(v, x0 - 1, y0).log2_size) > depth;
  if (y0 > 0) ctx += (kCtbLog2 - cu_at(v, x0, y0 - 1).log2_size) > depth;
  put_ctx(s, CTX_SPLIT_CU + ctx, split);
}

// coding_unit header (7.3.8.5-7.3.8.9) up to and including the cbf flags; returns the cbf mask
// of the transform blocks whose residual follows (0 = none).
__device__ __forceinline__ int put_cu_header(Syn &s, const CuView &v, const CuInfo &cu, int x0, int y0, int log2)
{
  const FrameParams &fp = *v.fp;
  const int n = 1 << log2;
  bool tu = false;
  const bool intra = cu.pred_mode == 1;
  if (!fp.is_idr) {
    int ctx = 0;
    if (x0 > 0) ctx += cu_at(v, x0 - 1, y0).skip;
    if (y0 > 0) ctx += cu_at(v, x0, y0 - 1).skip;
    put_ctx(s, CTX_SKIP + ctx, cu.skip);
    if (intra) {
      put_ctx(s, CTX_PRED_MODE, 1);                 // intra CU in a P slice
    } else if (cu.merge_idx != 0xff) {
      int midx = cu.merge_idx;
      if (!cu.skip) {
        put_ctx(s, CTX_PRED_MODE, 0);
        put_ctx(s, CTX_PART_MODE, 1);
        put_ctx(s, CTX_MERGE_FLAG, 1);
      }
      put_ctx(s, CTX_MERGE_IDX, midx > 0);
      for (int i = 1; i < kMaxMerge - 1 && midx >= i; i++) put_byp(s, midx > i, 1);
      tu = !cu.skip;                                // rqt_root_cbf inferred 1
    } else {
      put_ctx(s, CTX_PRED_MODE, 0);
      put_ctx(s, CTX_PART_MODE, 1);
      put_ctx(s, CTX_MERGE_FLAG, 0);
      // AMVP predictor (8.5.3.2.6-7): first available of (A0,A1), first of (B0,B1,B2), zero padding
      unsigned cur = coding_order_c(fp, x0, y0);
      NbMv a = nb_mv(v, cur, x0 - 1, y0 + n);
      if (!a.ok) a = nb_mv(v, cur, x0 - 1, y0 + n - 1);
      NbMv b = nb_mv(v, cur, x0 + n, y0 - 1);
      if (!b.ok) b = nb_mv(v, cur, x0 + n - 1, y0 - 1);
      if (!b.ok) b = nb_mv(v, cur, x0 - 1, y0 - 1);
      int px[2], py[2], k = 0;
      if (a.ok) { px[k] = a.mvx; py[k++] = a.mvy; }
      if (b.ok && !(a.ok && a.mvx == b.mvx && a.mvy == b.mvy)) { px[k] = b.mvx; py[k++] = b.mvy; }
      while (k < 2) { px[k] = 0; py[k++] = 0; }
      int pi = cu.mvp_idx;
      put_mvd(s, cu.mvx - px[pi], cu.mvy - py[pi]);
      put_ctx(s, CTX_MVP_IDX, pi);
      put_ctx(s, CTX_RQT_ROOT_CBF, cu.cbf != 0);
      tu = cu.cbf != 0;
    }
  }
  if (intra) {
    if (log2 == 3) put_ctx(s, CTX_PART_MODE, 1);
    // candidate modes (8.4.2): a neighbour that is not intra-coded, or lies in the CTB row above, counts as DC
    int cand[3];
    int a = 1, b = 1;
    if (x0 > 0) { const CuInfo &l = cu_at(v, x0 - 1, y0); if (l.pred_mode == 1) a = l.intra_mode; }
    if (y0 > 0 && (y0 & (kCtb - 1))) { const CuInfo &u = cu_at(v, x0, y0 - 1); if (u.pred_mode == 1) b = u.intra_mode; }
    if (a == b) {
      if (a < 2) { cand[0] = 0; cand[1] = 1; cand[2] = 26; }
      else { cand[0] = a; cand[1] = 2 + ((a + 29) % 32); cand[2] = 2 + ((a - 2 + 1) % 32); }
    } else {
      cand[0] = a; cand[1] = b;
      cand[2] = (a != 0 && b != 0) ? 0 : ((a != 1 && b != 1) ? 1 : 26);
    }
    int mode = cu.intra_mode, mpm = -1;
    for (int i = 0; i < 3; i++) if (cand[i] == mode) mpm = i;
    put_ctx(s, CTX_PREV_INTRA_LUMA, mpm >= 0);
    if (mpm >= 0) {
      put_byp(s, mpm > 0, 1);
      if (mpm > 0) put_byp(s, mpm > 1, 1);
    } else {
      if (cand[0] > cand[1]) { int tt = cand[0]; cand[0] = cand[1]; cand[1] = tt; }
      if (cand[0] > cand[2]) { int tt = cand[0]; cand[0] = cand[2]; cand[2] = tt; }
      if (cand[1] > cand[2]) { int tt = cand[1]; cand[1] = cand[2]; cand[2] = tt; }
      int rem = mode;
      for (int i = 2; i >= 0; i--) if (rem > cand[i]) rem--;
      put_byp(s, (unsigned)rem, 5);
    }
    put_ctx(s, CTX_INTRA_CHROMA, 0);
    tu = true;
  }
  if (!tu) return 0;
  // transform_tree at depth 0 with no split (7.3.8.8): cbf_cb, cbf_cr, cbf_luma
  int cb = (cu.cbf >> 1) & 1, cr = (cu.cbf >> 2) & 1, lu = cu.cbf & 1;
  put_ctx(s, CTX_CBF_CHROMA, cb);
  put_ctx(s, CTX_CBF_CHROMA, cr);
  if (cu.pred_mode == 1 || cb || cr) put_ctx(s, CTX_CBF_LUMA + 1, lu);
  if (fp.ctu_qp && cu.cbf) {
    // transform_unit (7.3.8.10): cu_qp_delta_abs / sign once per quantisation group (= CTU), in the
    // first TU with a coded block flag (k_cu_qps found it).  Binarisation 9.3.3.10: prefix TR cMax 5
    // (first bin ctx 0, then ctx 1), suffix EG0 bypass, sign bypass.
    const int ctu = (y0 >> kCtbLog2) * fp.ctb_cols + (x0 >> kCtbLog2);
    if (xy_to_z((x0 >> 3) & 7, (y0 >> 3) & 7) == fp.ctu_first[ctu]) {
      const int d = fp.ctu_delta[ctu], a = abs(d);
      const int pre = min(a, 5);
      for (int i = 0; i < pre; i++) put_ctx(s, CTX_CU_QP_DELTA + (i ? 1 : 0), 1);
      if (pre < 5) {
        put_ctx(s, CTX_CU_QP_DELTA + (pre ? 1 : 0), 0);
      } else {
        int v = a - 5, k = 0, ones = 0;
        while (v >= (1 << k)) { ones++; v -= 1 << k; k++; }
        put_byp(s, ((((1ull << (ones + 1)) - 2) << k) | (unsigned)v), ones + 1 + k);
      }
      if (a) put_byp(s, d < 0, 1);
    }
  }
  return cu.cbf;
}

// sao() of one CTU (7.3.8.3), both slice flags on.  fp.sao_flags bit 2: a CTU whose parameters equal
// its left (else its upper) neighbour's codes the merge flag instead of repeating them.
// sao_type_idx: TR cMax 2, first bin context-coded; sao_offset_abs: TR cMax 7, bypass; signs, band
// position (5 bits) and edge class (2 bits) bypass.
__device__ __forceinline__ void put_sao(Syn &s, const FrameParams &fp, int ctu)
{
  const SaoCtu p = fp.sao[ctu];
  const int rx = ctu % fp.ctb_cols, ry = ctu / fp.ctb_cols;
  auto same = [&](int other) {
    const uint32_t *a = (const uint32_t *)&p, *b = (const uint32_t *)(fp.sao + other);
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3] && ((a[4] ^ b[4]) & 0xffffffu) == 0;
  };
  const bool try_merge = (fp.sao_flags & 4) != 0;
  if (rx > 0) {
    const bool m = try_merge && same(ctu - 1);
    put_ctx(s, CTX_SAO_MERGE, m);
    if (m) return;
  }
  if (ry > 0) {
    const bool m = try_merge && same(ctu - fp.ctb_cols);
    put_ctx(s, CTX_SAO_MERGE, m);
    if (m) return;
  }
  for (int comp = 0; comp < 3; comp++) {
    const int g = comp ? 1 : 0, type = p.type[g];
    if (comp < 2) {
      put_ctx(s, CTX_SAO_TYPE, type != 0);
      if (type) put_byp(s, type == 2, 1);
    }
    if (!type) continue;
    for (int k = 0; k < 4; k++) {
      const int a = abs((int)p.offset[comp][k]);
      if (a < 7) put_byp(s, ((1ull << a) - 1) << 1, a + 1);
      else put_byp(s, 0x7f, 7);
    }
    if (type == 1) {
      for (int k = 0; k < 4; k++)
        if (p.offset[comp][k]) put_byp(s, p.offset[comp][k] < 0, 1);
      put_byp(s, p.band_pos[comp], 5);
    } else if (comp < 2) {
      put_byp(s, p.eo_class[g], 2);
    }
  }
}

// Record region of the CU whose first unit is (ctu, z): fixed address, no prefix sums needed.
// Word 0 = record count | log2 CU size << 24; records follow.
__device__ __forceinline__ uint32_t *cu_region(uint32_t *recs, int ctu, int z) { return recs + ((size_t)ctu * 64 + z) * kRecUnitCap; }

constexpr int kBinWarps = 2;           // warps (CUs) per CTA of the binariser

struct BinWarpShared {
  int16_t lv[32 * 32];
  ResidualShared rs;
  uint32_t syn[192];
};

// One warp per 32x32 quadrant of a CTU; the warp walks the (1..16) CUs that start inside it.
// (One warp per 8x8 unit would launch 16k CTAs of which three quarters exit at once.)
__global__ void __launch_bounds__(32 * kBinWarps)
k_binarise(FrameParams fp, const CuInfo *__restrict__ cu, const int16_t *__restrict__ levels, uint32_t *__restrict__ recs)
{
  __shared__ BinWarpShared s_all[kBinWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int quad = blockIdx.x * kBinWarps + wib;            // CTU-major, four quadrants per CTU
  const int ctu = quad >> 2;
  if (ctu >= fp.ctb_cols * fp.ctb_rows) return;
  const int cx = (ctu % fp.ctb_cols) * kCtb, cy = (ctu / fp.ctb_cols) * kCtb;
  CuView v{&fp, cu};
  BinWarpShared &sh = s_all[wib];
  const size_t ysz = (size_t)fp.w * fp.h;
  for (int z = (quad & 3) * 16; z < (quad & 3) * 16 + 16;) {
    const int x0 = cx + 8 * z_to_x(z), y0 = cy + 8 * z_to_y(z);
    if (x0 >= fp.w || y0 >= fp.h) { z++; continue; }
    const CuInfo cur = cu_at(v, x0, y0);
    const int log2 = cur.log2_size;
    const int units = 1 << (2 * (log2 - 3));
    if (z & (units - 1)) { z++; continue; }                  // inside a CU that started in another quadrant (64x64)
    __syncwarp();
    Syn syn{sh.syn, 0};
    uint32_t *out = cu_region(recs, ctu, z);
    int o = 1;                                                // word 0 is the header
    if (z == 0 && fp.sao) put_sao(syn, fp, ctu);              // the CTU's sao() syntax precedes its first CU
    // split flags of every ancestor block that starts at this unit, top down, then this CU's own
    for (int L = 6; L > 3; L--) {
      if (L < log2) break;
      if (z & ((1 << (2 * (L - 3))) - 1)) continue;
      put_split(syn, v, x0, y0, L, 6 - L, L > log2);
    }
    const int cbf = put_cu_header(syn, v, cur, x0, y0, log2);
    __syncwarp();
    for (int i = lane; i < syn.k; i += 32) out[o + i] = syn.r[i];
    o += syn.k;
    for (int k = 0; k < 3; k++) {
      if (!((cbf >> k) & 1)) continue;
      const int sft = k ? 1 : 0, l2 = log2 - sft, n = 1 << l2, pw = fp.w >> sft;
      const int16_t *plane = levels + (k == 0 ? 0 : ysz + (k == 2 ? ysz / 4 : 0)) + (size_t)(y0 >> sft) * pw + (x0 >> sft);
      __syncwarp();
      for (int i = lane; i < n * n; i += 32) sh.lv[i] = plane[(size_t)(i >> l2) * pw + (i & (n - 1))];
      __syncwarp();
      syn.k = 0;
      const int last_sb = binarise_residual(syn, sh.rs, sh.lv, l2, k, scan_idx_for(cur.pred_mode, cur.intra_mode, l2, k), lane);
      if (last_sb < 0) continue;
      for (int i = lane; i < syn.k; i += 32) out[o + i] = syn.r[i];      // last significant position
      o += syn.k;
      // sub-block record lists, concatenated in coding order (last_sb down to 0): exclusive scan of the counts
      for (int base = last_sb; base >= 0; base -= 32) {
        const int i = base - lane;                                       // lane 0 owns the first sub-block in coding order
        const int cnt = i >= 0 ? sh.rs.cnt[i] : 0;
        int pre = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, pre, d);
          if (lane >= d) pre += t;
        }
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        const int start = o + pre - cnt;
        for (int j = 0; j < cnt; j++) out[start + j] = sh.rs.rec[i][j];
        o += total;
      }
    }
    if (lane == 0) out[0] = (uint32_t)(o - 1) | ((uint32_t)log2 << 24);
    z += units;
  }
}

// ---- entropy coding in two phases ---------------------------------------------------------------
// The state of a CABAC context evolves with the bins coded in it (MPS / LPS outcomes) and with
// nothing else -- not with range or low.  So the arithmetic coder is split:
//
//   ctx_rows     (phase A) walks the bin records of a CTU row 32 at a time and replaces every
//                context-coded record (context index, bin) by (pStateIdx, is-LPS).  Records of one
//                batch that share a context are grouped with __match_any_sync and walked in order
//                by the group's first lane; the groups run side by side.  The
//                WPP hand-over (contexts after the second CTU of the row above) lives here, so the
//                two-CTU stagger between rows costs two CTUs of this cheap pass only.
//   arith_rows   (phase B) is the serial range coder over the resolved records: no context table,
//                no dependency between rows at all, and the rangeTabLps row of the next record is
//                fetched while the current one is coded.
//
// Before the split one kernel did both and every row waited for the row above to arithmetic-code
// two CTUs: 3.4 ms per 1080p P picture, 26 ms per IDR (profiles/r01_summary.md).  The two phases run
// as the two halves of one launch (k_entropy_rows): phase B follows phase A of its substream CTU by CTU
// instead of waiting for the whole picture.

__device__ void
ctx_rows(const FrameParams &fp, const int sub, const CuInfo *__restrict__ cu, uint32_t *recs, uint8_t *sync_ctx, int *sync_flag,
         int *row_prog)
{
  __shared__ uint8_t s_ctx[CTX_COUNT + 2];
  __shared__ uint8_t s_trans[64];
  __shared__ uint8_t s_out[32];
  const int lane = threadIdx.x;
  const int r_first = fp.no_wpp ? 0 : sub, r_last = fp.no_wpp ? fp.ctb_rows - 1 : sub;
  for (int i = lane; i < 64; i += 32) s_trans[i] = c_trans_lps[i];
  if (r_first == 0 || fp.ctb_cols < 2) {
    const int qp = clip3(0, 51, fp.qp), init_type = fp.init_type;
    for (int i = lane; i < CTX_COUNT; i += 32) {
      int iv = c_ctx_init[init_type][i];
      int m = (iv >> 4) * 5 - 45, n = ((iv & 15) << 3) - 16;
      int pre = clip3(1, 126, ((m * qp) >> 4) + n);
      int mps = pre <= 63 ? 0 : 1;
      int st = mps ? pre - 64 : 63 - pre;
      s_ctx[i] = (uint8_t)((st << 1) | mps);
    }
  } else {
    volatile int *f = sync_flag;
    while (f[r_first - 1] == 0) __nanosleep(100);
    __threadfence();
    for (int i = lane; i < CTX_COUNT; i += 32) s_ctx[i] = __ldcg(sync_ctx + (size_t)(r_first - 1) * CTX_COUNT + i);
  }
  __syncwarp();
  for (int r = r_first; r <= r_last; r++) {
    for (int col = 0; col < fp.ctb_cols; col++) {
      // the CUs of this CTU from the cu map: unit z starts a CU when z is a multiple of the CU's unit count
      unsigned starts[2];
      for (int h = 0; h < 2; h++) {
        const int z = 32 * h + lane;
        const int x8 = col * 8 + z_to_x(z), y8 = r * 8 + z_to_y(z);
        bool st = false;
        if (x8 < fp.w8 && y8 < fp.h8) {
          const int l2 = cu[(size_t)y8 * fp.w8 + x8].log2_size;
          st = l2 >= 3 && (z & ((1 << (2 * (l2 - 3))) - 1)) == 0;
        }
        starts[h] = __ballot_sync(0xffffffffu, st);
      }
      const int n0 = __popc(starts[0]), n_cu = n0 + __popc(starts[1]);
      uint32_t *const ctu_recs = recs + (size_t)(r * fp.ctb_cols + col) * 64 * kRecUnitCap;
      // record counts of up to 64 CUs, fetched in parallel (lane j: CUs j and j + 32 of the CTU)
      int zs[2] = {64, 64}, cnts[2] = {0, 0};
      for (int h = 0; h < 2; h++) {
        const int j = 32 * h + lane;
        if (j < n_cu) {
          zs[h] = j < n0 ? (int)__fns(starts[0], 0, j + 1) : 32 + (int)__fns(starts[1], 0, j - n0 + 1);
          cnts[h] = (int)(__ldg(ctu_recs + (size_t)zs[h] * kRecUnitCap) & 0xffffffu);
        }
      }
      // Records are fetched four chunks (128 records) at a time and one such group ahead: a chunk
      // resolves in a few hundred cycles, far less than a round trip to L2, and this warp has
      // nothing else to run meanwhile.
      auto load4 = [&](const uint32_t *rg, int count, int base, uint32_t v4[4]) {
#pragma unroll
        for (int j = 0; j < 4; j++) v4[j] = base + 32 * j + lane < count ? __ldcg(rg + 1 + base + 32 * j + lane) : 0x80000000u;
      };
      int z = __shfl_sync(0xffffffffu, zs[0], 0), cnt = __shfl_sync(0xffffffffu, cnts[0], 0);
      uint32_t *reg = ctu_recs + (size_t)z * kRecUnitCap;
      uint32_t v4[4];
      load4(reg, n_cu > 0 ? cnt : 0, 0, v4);
      for (int k = 0; k < n_cu; k++) {
        // cursor of the next CU, whose first group is fetched while this CU's last group is resolved
        const int kn = k + 1;
        const int zn = __shfl_sync(0xffffffffu, zs[kn >> 5 & 1], kn & 31), cntn = __shfl_sync(0xffffffffu, cnts[kn >> 5 & 1], kn & 31);
        uint32_t *regn = ctu_recs + (size_t)zn * kRecUnitCap;
        for (int base = 0; base < cnt; base += 128) {
          uint32_t n4[4];
          if (base + 128 < cnt) load4(reg, cnt, base + 128, n4);
          else load4(regn, kn < n_cu ? cntn : 0, 0, n4);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int off = base + 32 * j;
            if (off >= cnt) break;
            const uint32_t v = v4[j];
            const bool active = off + lane < cnt && !(v & 0x80000000u);
            const unsigned ctx = v >> 1, bin = v & 1;
            // records of this chunk that share a context form a group; the first lane of each group
            // walks its members in order, entirely in registers, and leaves their results in s_out
            const unsigned grp = __match_any_sync(0xffffffffu, active ? ctx : 0x10000u + lane);
            const unsigned bins32 = __ballot_sync(0xffffffffu, active && bin);
            if (active && (grp & ((1u << lane) - 1)) == 0) {
              const unsigned s = s_ctx[ctx];
              unsigned st = s >> 1, mps = s & 1;
              for (unsigned m = grp; m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                const unsigned is_lps = ((bins32 >> j) & 1) != mps;
                s_out[j] = (uint8_t)((st << 1) | is_lps);
                if (is_lps) { mps ^= (st == 0); st = s_trans[st]; }
                else st = min(st + 1, 62u);
              }
              s_ctx[ctx] = (uint8_t)((st << 1) | mps);
            }
            __syncwarp();
            if (active) reg[1 + off + lane] = s_out[lane];
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; j++) v4[j] = n4[j];
        }
        if (cnt == 0) load4(regn, kn < n_cu ? cntn : 0, 0, v4);     // (a CU always has records; keep the cursor sound anyway)
        z = zn; cnt = cntn; reg = regn;
      }
      if (col == 1 && r + 1 < fp.ctb_rows && !fp.no_wpp) {       // WPP: hand the contexts to the row below
        __syncwarp();
        for (int i = lane; i < CTX_COUNT; i += 32) sync_ctx[(size_t)r * CTX_COUNT + i] = s_ctx[i];
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicExch(&sync_flag[r], 1);
      }
      // the CTU's records are resolved: the range coder of this substream (the other role of the launch) may take it
      __threadfence();
      __syncwarp();
      if (row_prog && lane == 0) atomicExch(&row_prog[r_first], (r - r_first) * fp.ctb_cols + col + 1);
    }
  }
}

// phase B: context-coded records are (pStateIdx << 1) | is-LPS here
__device__ __forceinline__ void enc_bin_resolved(Coder &c, const uint2 e, unsigned is_lps)
{
  // Branch-free: a lone warp pays a pipeline refill for every taken branch, and LPS / MPS /
  // renormalise-by-one are only selects on the same three registers.
  const uint32_t lps = (e.x >> (((c.range >> 6) & 3) * 8)) & 0xff;
  const uint32_t rmps = c.range - lps;
  const int nb = is_lps ? __clz(lps) - 23 : (rmps < 256 ? 1 : 0);
  c.low = (c.low + (is_lps ? rmps : 0u)) << nb;
  c.range = (is_lps ? lps : rmps) << nb;
  c.bits_left -= nb;
  c.bins++;
  if (c.bits_left < 12) write_out(c);
}

__device__ void
arith_rows(const FrameParams &fp, const int sub, const uint32_t *__restrict__ recs, uint8_t *rows, uint32_t row_cap, uint32_t *row_len,
           unsigned long long *bins, const int *row_prog)
{
  __shared__ uint2 s_tab[64];
  __shared__ uint32_t s_chunk[2][32];
  // WPP: one substream per CTU row (r_first == r_last == sub).  no_wpp (a tile without
  // entropy_coding_sync): one block codes every row into a single substream.
  const int lane = threadIdx.x;
  const int r_first = fp.no_wpp ? 0 : sub, r_last = fp.no_wpp ? fp.ctb_rows - 1 : sub;
  // Phase A of this substream runs concurrently (the other role of the launch) and publishes how many
  // CTUs it has resolved; a CTU's records are read only after that (all lanes poll: no lane-dependent branch).
  auto wait_ctu = [&](int r, int col) {
    if (!row_prog) return;                                  // phase A ran to completion in a launch of its own
    const int need = (r - r_first) * fp.ctb_cols + col + 1;
    volatile const int *p = row_prog + r_first;
    while (*p < need) __nanosleep(64);
    __threadfence();
  };
  for (int i = lane; i < 64; i += 32) s_tab[i] = make_uint2(c_range_lps[i], 0u);
  Coder c;
  c.out = rows + (size_t)r_first * row_cap; c.pos = 0; c.cap = fp.no_wpp ? row_cap * fp.ctb_rows : row_cap;
  c.zeros = 0; c.bins = 0;
  c.writer = lane == 0;
  __syncwarp();
  coder_start(c);
  for (int r = r_first; r <= r_last; r++) {
    // Flat walk over the CUs of the row.  The address of the next CU's record list is known as
    // soon as the current header is read, so its header and first 32 records are fetched while the
    // current list is being coded (a lone warp has nothing else to hide the ~1 us load latency).
    const int cy = r * kCtb;
    auto first_unit = [&](int col, int z) -> int {          // first z >= given whose unit lies inside the picture, 64 if none
      while (z < 64 && (col * kCtb + 8 * z_to_x(z) >= fp.w || cy + 8 * z_to_y(z) >= fp.h)) z++;
      return z;
    };
    int col = 0, z = first_unit(0, 0);
    int buf = 0;
    const uint32_t *reg = recs + ((size_t)(r * fp.ctb_cols) * 64 + z) * kRecUnitCap;
    wait_ctu(r, 0);
    uint32_t hdr = __ldcg(reg);
    uint32_t first = __ldcg(reg + 1 + lane);
    while (col < fp.ctb_cols) {
      const int cnt = (int)(hdr & 0xffffffu), log2 = (int)(hdr >> 24);
      // cursor of the CU after this one
      int ncol = col, nz = first_unit(col, z + (1 << (2 * (log2 - 3))));
      if (nz >= 64) { ncol = col + 1; nz = ncol < fp.ctb_cols ? first_unit(ncol, 0) : 64; }
      const uint32_t *nreg = recs + ((size_t)(r * fp.ctb_cols + ncol) * 64 + nz) * kRecUnitCap;
      uint32_t nhdr = 0, nfirst = 0;
      if (ncol < fp.ctb_cols) {
        if (ncol != col) wait_ctu(r, ncol);                 // the fetch ahead crosses into the next CTU
        nhdr = __ldcg(nreg); nfirst = __ldcg(nreg + 1 + lane);
      }
      // code this CU: each 32-record chunk is staged in shared memory (double buffered); the
      // rangeTabLps row of record k + 1 is fetched while record k is coded -- neither load depends
      // on the coder state
      // (`buf` keeps alternating across CUs: a chunk buffer is rewritten only after the __syncwarp of the
      // chunk in between, which every lane reaches after its last read -- found by racecheck)
      uint32_t curv = first;
      for (int base = 0; base < cnt; base += 32) {
        const int nb = base + 32;
        uint32_t nxt = nb + lane < cnt ? __ldcg(reg + 1 + nb + lane) : 0;
        s_chunk[buf][lane] = curv;
        __syncwarp();
        const int m = min(32, cnt - base);
        const uint32_t *ch = s_chunk[buf];
        uint32_t vn = ch[0];
        uint2 en = s_tab[(vn >> 1) & 63];
        for (int k = 0; k < m; k++) {
          const uint32_t v = vn;
          const uint2 e = en;
          vn = ch[min(k + 1, 31)];
          en = s_tab[(vn >> 1) & 63];
          if (v & 0x80000000u) enc_bypass_group(c, v & 0xffffu, (int)((v >> 24) & 31));
          else enc_bin_resolved(c, e, v & 1);
        }
        curv = nxt;
        buf ^= 1;
      }
      if (ncol != col) {                                    // the CTU is complete
        const bool end_sub = col == fp.ctb_cols - 1 && (!fp.no_wpp || r == fp.ctb_rows - 1);
        const bool last = r == fp.ctb_rows - 1 && col == fp.ctb_cols - 1 && !fp.more_tiles;
        enc_terminate(c, last);                                           // end_of_slice_segment_flag
        if (end_sub && !last) enc_terminate(c, 1);                        // end_of_subset_one_bit
      }
      col = ncol; z = nz; reg = nreg; hdr = nhdr; first = nfirst;
    }
  }                                                                       // rows of this substream
  coder_finish(c);
  if (lane == 0) {
    row_len[r_first] = c.pos <= c.cap ? c.pos : 0xffffffffu;
    for (int i = r_first + 1; i <= r_last; i++) row_len[i] = 0;
    atomicAdd(bins, c.bins);
  }
}

// Both phases of the entropy coder in one launch: CTAs [0, subs) resolve the context states of their
// substream (phase A), CTAs [subs, 2 * subs) run its range coder (phase B) one CTU behind.  Every wait is
// on a CTA with a smaller index (phase A of row r on phase A of row r - 1, phase B on phase A), which the
// hardware has placed before it, so the waits cannot starve what they wait for.
__global__ void __launch_bounds__(32)
k_entropy_rows(FrameParams fp, const CuInfo *__restrict__ cu, uint32_t *recs, uint8_t *sync_ctx, int *sync_flag, int *row_prog,
               uint8_t *rows, uint32_t row_cap, uint32_t *row_len, unsigned long long *bins)
{
  const int subs = fp.no_wpp ? 1 : fp.ctb_rows;
  if ((int)blockIdx.x < subs) ctx_rows(fp, (int)blockIdx.x, cu, recs, sync_ctx, sync_flag, row_prog);
  else arith_rows(fp, (int)blockIdx.x - subs, recs, rows, row_cap, row_len, bins, row_prog);
}

// ... or as two launches, phase B after phase A: no warp ever waits for another launch's progress, which
// is what a deep pipeline wants (dozens of pictures' entropy coders share the SMs with the prediction
// chain; the polling of the fused form costs the chain a few per cent).
__global__ void __launch_bounds__(32)
k_ctx_rows(FrameParams fp, const CuInfo *__restrict__ cu, uint32_t *recs, uint8_t *sync_ctx, int *sync_flag)
{
  ctx_rows(fp, (int)blockIdx.x, cu, recs, sync_ctx, sync_flag, nullptr);
}
__global__ void __launch_bounds__(32)
k_arith_rows(FrameParams fp, const uint32_t *__restrict__ recs, uint8_t *rows, uint32_t row_cap, uint32_t *row_len,
             unsigned long long *bins)
{
  arith_rows(fp, (int)blockIdx.x, recs, rows, row_cap, row_len, bins, nullptr);
}

// Gather the substreams into one contiguous buffer (normally mapped pinned host memory, so the
// bytes cross PCIe exactly once and the host needs a single event wait per picture).
// hdr[0] = total bytes, hdr[1..rows] = row lengths (0xffffffff marks an overflowed row).
__global__ void __launch_bounds__(128)
k_pack_rows(int rows, const uint8_t *__restrict__ src, uint32_t row_cap, const uint32_t *__restrict__ row_len,
            uint8_t *dst, uint32_t dst_cap, uint32_t *hdr)
{
  __shared__ uint32_t s_off, s_len, s_total;
  const int r = blockIdx.x;
  if (threadIdx.x == 0) {
    uint32_t off = 0, total = 0;
    bool bad = false;
    for (int i = 0; i < rows; i++) {
      uint32_t l = row_len[i];
      if (l == 0xffffffffu) { bad = true; l = 0; }
      if (i < r) off += l;
      total += l;
    }
    if (bad || total > dst_cap) total = 0xffffffffu;
    s_off = off; s_len = row_len[r] == 0xffffffffu ? 0 : row_len[r]; s_total = total;
    if (r == 0) hdr[0] = total;
    hdr[1 + r] = row_len[r];
  }
  __syncthreads();
  if (s_total == 0xffffffffu) return;
  const uint8_t *s = src + (size_t)r * row_cap;
  uint8_t *d = dst + s_off;
  const uint32_t n = s_len;
  // align the destination to 4 bytes, then move words (source alignment handled by byte gathers)
  uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)d & 3)) & 3));
  for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) d[i] = s[i];
  const uint32_t words = (n - head) >> 2;
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) {
    const uint8_t *q = s + head + 4 * i;
    ((uint32_t *)(d + head))[i] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
  }
  for (uint32_t i = head + 4 * words + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

}  // namespace

cudaError_t launch_pack_rows(int rows, const uint8_t *src, uint32_t row_cap, const uint32_t *row_len, uint8_t *dst,
                             uint32_t dst_cap, uint32_t *hdr, cudaStream_t s)
{
  k_pack_rows<<<rows, 128, 0, s>>>(rows, src, row_cap, row_len, dst, dst_cap, hdr);
  return cudaGetLastError();
}

cudaError_t launch_binarise(const FrameParams &fp, const CuInfo *cu, const int16_t *levels, uint32_t *recs, cudaStream_t s)
{
  const int quads = fp.ctb_cols * fp.ctb_rows * 4;
  k_binarise<<<(quads + kBinWarps - 1) / kBinWarps, 32 * kBinWarps, 0, s>>>(fp, cu, levels, recs);
  return cudaGetLastError();
}

cudaError_t launch_arith(const FrameParams &fp, const CuInfo *cu, uint32_t *recs, uint8_t *rows, uint32_t row_cap,
                         uint32_t *row_len, uint8_t *sync_ctx, int *sync_flag, unsigned long long *bins, bool fused, cudaStream_t s)
{
  // sync_flag[rows] is followed by progress[rows] in the encoder's small state: phase A's per-substream CTU count
  cudaError_t e = cudaMemsetAsync(sync_flag, 0, 2 * sizeof(int) * fp.ctb_rows, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(bins, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  const int subs = fp.no_wpp ? 1 : fp.ctb_rows;
  if (fused) {
    k_entropy_rows<<<2 * subs, 32, 0, s>>>(fp, cu, recs, sync_ctx, sync_flag, sync_flag + fp.ctb_rows, rows, row_cap, row_len, bins);
  } else {
    k_ctx_rows<<<subs, 32, 0, s>>>(fp, cu, recs, sync_ctx, sync_flag);
    k_arith_rows<<<subs, 32, 0, s>>>(fp, recs, rows, row_cap, row_len, bins);
  }
  return cudaGetLastError();
}

}  // namespace b200
