// CABAC parsing of the B200 HEVC decoder (sm_100a); SURVEY.md 8a row a7 (OpenHEVC replacement).
//
// One warp per WPP substream (CTU row); the entry points of the slice header give every row its
// own byte range, so all rows of a picture parse concurrently, each two CTUs behind the row above
// (context hand-over after the second CTU, H.265 9.3.1, and the above / above-right neighbours
// that merge / AMVP derivation reads).  Parsing cannot be split from arithmetic decoding -- what
// to read next depends on what was just decoded -- so the warp runs the parser redundantly in all
// lanes on one shared context table (loads are broadcasts), with no branch that depends on the lane
// index; stores of identical values from all lanes coalesce into one transaction.
//
// Scope: CUs 8..64 (inter 2Nx2N; intra 2Nx2N and NxN with explicit chroma modes), transform trees down
// to 4x4 luma blocks, I and P slices, several reference pictures with temporal candidates, SAO (with
// merge candidates), cu_qp_delta per CTU, sign data hiding (scaling lists act on the dequantiser, not here);
// no PCM / AMP / transform skip.  Anything else is reported through `status` and the picture is rejected by the host.
#include "hevc_device.cuh"
#include "hevc_kernels.h"

namespace b200 {

namespace {

// sig_coeff_flag context increments (9.3.4.2.5) by scan position, two bits per position:
// c_sig_pat[scan_idx][prevCsbf]; for 4x4 blocks four bits per position from ctxIdxMap.
static __constant__ uint32_t c_sig_pat[3][4] = {{0x00000556u, 0x01090926u, 0x0010619au, 0xaaaaaaaau}, {0x00010516u, 0x000055aau, 0x06060606u, 0xaaaaaaaau}, {0x00010516u, 0x06060606u, 0x000055aau, 0xaaaaaaaau}};
// inverse up-right diagonal scans: scan index of position (x, y), row-major
static __constant__ uint8_t c_inv_diag2[4] = {0, 2, 1, 3};
static __constant__ uint8_t c_inv_diag4[16] = {0, 2, 5, 9, 1, 4, 8, 12, 3, 7, 11, 14, 6, 10, 13, 15};
static __constant__ uint8_t c_inv_diag8[64] = {0, 2, 5, 9, 14, 20, 27, 35, 1, 4, 8, 13, 19, 26, 34, 42, 3, 7, 12, 18, 25, 33, 41, 48, 6, 11, 17, 24, 32, 40, 47, 53, 10, 16, 23, 31, 39, 46, 52, 57, 15, 22, 30, 38, 45, 51, 56, 60, 21, 29, 37, 44, 50, 55, 59, 62, 28, 36, 43, 49, 54, 58, 61, 63};
static __constant__ unsigned long long c_sig_pat4[3] = {0x8885875467436120ull, 0x8877886654325410ull, 0x8855884476317620ull};

// Context table: one entry per context, shared by the lanes (they all run the same decoder: loads are
// broadcasts, stores write the same value).  An entry carries what a bin needs in ONE load -- .x = the four
// rangeTabLps bytes of its state, .y = (pStateIdx << 1) | valMps -- and the entry a context moves to comes
// from `next` (indexed by .y * 2 + is-LPS), fetched while the bin's result is already in use.  The updated
// entry is written back at the start of the next decodeDecision, behind that bin's own table load, so that
// neither the transition lookup nor the store sits on the value / range dependency chain; a bin in the same
// context takes it from the registers.  (A lone warp pays the full latency of every dependent load.)
struct Reader {
  const uint8_t *p, *end;   // next byte to fetch INTO the window / end of the substream
  // The next 64 bytes of the substream sit in two registers per lane, one byte per lane each: a byte is
  // handed out by a shuffle, and a fresh half is fetched (one coalesced 32-byte load) a whole half ahead of
  // its first use -- instead of one dependent global load per byte.
  uint32_t win, win_next;
  int widx, lane;
  uint32_t range, value;
  int bits_needed;
  uint2 *ctx;            // [CTX_COUNT + 1]; entry CTX_COUNT takes the write-back while nothing is pending
  const uint2 *next;     // [128 * 2]
  uint2 pend;            // the update of context pend_idx, not yet in the table
  uint32_t pend_idx;
  int err;
};

__device__ __forceinline__ uint2 ctx_entry(uint32_t state_mps) { return make_uint2(c_range_lps[state_mps >> 1], state_mps); }
// all pending updates into the table (before it is read as a whole: WPP hand-over)
__device__ __forceinline__ void ctx_flush(Reader &r) { r.ctx[r.pend_idx] = r.pend; __syncwarp(); }

__device__ __forceinline__ uint32_t window_load(const Reader &r)
{
  const uint8_t *q = r.p + r.lane;
  return q < r.end ? (uint32_t)__ldg(q) : 0u;      // past the end: zeros (a truncated substream decodes to an error, not a fault)
}

__device__ __forceinline__ uint32_t next_byte(Reader &r)
{
  const uint32_t b = __shfl_sync(0xffffffffu, r.win, r.widx);
  if (++r.widx == 32) {
    r.widx = 0;
    r.win = r.win_next;
    r.win_next = window_load(r);
    r.p += 32;
  }
  return b;
}

__device__ __forceinline__ void reader_start(Reader &r)
{
  r.widx = 0;
  r.win = window_load(r); r.p += 32;
  r.win_next = window_load(r); r.p += 32;
  r.range = 510;
  r.bits_needed = -8;
  r.value = next_byte(r) << 8;
  r.value |= next_byte(r);
}

__device__ __forceinline__ int dec_bin(Reader &r, int ctx_idx)
{
  // 9.3.4.3.2, written without branches up to the byte refill: a lone warp pays a pipeline refill
  // for every taken branch, and the MPS / LPS / renormalise cases are selects on the same registers.
  // The table is shared by the lanes of the warp, which all hold the same values and all do the same accesses;
  // the two barriers order one lane's write-back against the other lanes' loads (before and after it), which the
  // lockstep of a converged warp gives in practice and the memory model only with them (compute-sanitizer racecheck).
  const uint2 loaded = r.ctx[ctx_idx];
  __syncwarp();
  r.ctx[r.pend_idx] = r.pend;                                  // the previous bin's update, behind this bin's load
  __syncwarp();
  const uint2 e = (uint32_t)ctx_idx == r.pend_idx ? r.pend : loaded;
  const uint32_t lps = (e.x >> (((r.range >> 6) & 3) * 8)) & 0xff;
  const uint32_t rmps = r.range - lps, scaled = rmps << 7;
  const bool is_lps = r.value >= scaled;
  const int nb = is_lps ? __clz(lps) - 23 : (rmps < 256 ? 1 : 0);
  r.value = (r.value - (is_lps ? scaled : 0u)) << nb;
  r.range = (is_lps ? lps : rmps) << nb;
  const int bin = (int)((e.y & 1u) ^ (is_lps ? 1u : 0u));
  r.pend = r.next[e.y * 2 + (is_lps ? 1u : 0u)];
  r.pend_idx = (uint32_t)ctx_idx;
  r.bits_needed += nb;
  if (r.bits_needed >= 0) { r.value += next_byte(r) << r.bits_needed; r.bits_needed -= 8; }
  return bin;
}

__device__ __forceinline__ int dec_bypass(Reader &r)
{
  r.value += r.value;
  if (++r.bits_needed >= 0) { r.bits_needed = -8; r.value += next_byte(r); }
  uint32_t scaled = r.range << 7;
  if (r.value >= scaled) { r.value -= scaled; return 1; }
  return 0;
}

__device__ __forceinline__ uint32_t dec_bypass_bits(Reader &r, int n)
{
  uint32_t v = 0;
  for (int i = 0; i < n; i++) v = (v << 1) | (uint32_t)dec_bypass(r);
  return v;
}

__device__ __forceinline__ int dec_terminate(Reader &r)
{
  r.range -= 2;
  uint32_t scaled = r.range << 7;
  if (r.value >= scaled) return 1;
  if (scaled < (256u << 7)) {
    r.range = scaled >> 6;
    r.value += r.value;
    if (++r.bits_needed == 0) { r.bits_needed = -8; r.value += next_byte(r); }
  }
  return 0;
}

__device__ __forceinline__ void init_contexts_d(uint2 *ctx, int init_type, int qp, int lane)
{
  qp = clip3(0, 51, qp);
  for (int i = lane; i < CTX_COUNT; i += 32) {
    int iv = c_ctx_init[init_type][i];
    int m = (iv >> 4) * 5 - 45, n = ((iv & 15) << 3) - 16;
    int pre = clip3(1, 126, ((m * qp) >> 4) + n);
    int mps = pre <= 63 ? 0 : 1;
    int st = mps ? pre - 64 : 63 - pre;
    ctx[i] = ctx_entry((uint32_t)((st << 1) | mps));
  }
}

__device__ __forceinline__ void scan_pos_d(int scan_idx, int blk_log2, int i, int &x, int &y)
{
  int n = 1 << blk_log2;
  if (scan_idx == 1) { x = i & (n - 1); y = i >> blk_log2; return; }
  if (scan_idx == 2) { x = i >> blk_log2; y = i & (n - 1); return; }
  int v = blk_log2 == 2 ? c_diag4[i] : (blk_log2 == 1 ? c_diag2[i] : (blk_log2 == 3 ? c_diag8[i] : 0));
  x = v & 15; y = v >> 4;
}
// inverse scan: index of position (x,y) in the scan of a (1<<blk_log2)^2 array
__device__ __forceinline__ int scan_index_d(int scan_idx, int blk_log2, int x, int y)
{
  int n = 1 << blk_log2;
  if (scan_idx == 1) return y * n + x;
  if (scan_idx == 2) return x * n + y;
  if (blk_log2 == 2) return c_inv_diag4[y * 4 + x];
  if (blk_log2 == 3) return c_inv_diag8[y * 8 + x];
  if (blk_log2 == 1) return c_inv_diag2[y * 2 + x];
  return 0;
}

// Everything a CU reads from its neighbours (skip / depth contexts, merge and AMVP candidates,
// intra MPMs) lies in the current CTU, the CTU to its left, or the bottom unit line of the CTU row
// above between x = cx - 8 and cx + 71.  Those live in shared memory: a global (L2) round trip per
// neighbour cost more than the arithmetic decoding of the CU itself.
struct ParseCtx {
  const FrameParams &fp;  // the kernel's __grid_constant__ parameter: fields (ref_dist[] by index too) come from the constant bank, no local copy
  CuInfo *cu;            // global cu map (written by lane 0 for the later kernels and the row below)
  int16_t *levels;       // zeroed by the host before the launch; only non-zero levels are stored
  int lane;
  int max_mv;            // largest |mv component| seen (quarter samples)
  uint32_t *s_ctu;       // [2][64][4]: 8x8 units of the current (cur_buf) and the previous CTU, raster order
  uint32_t *s_above;     // [10][4]: units (cx/8 - 1 .. cx/8 + 8) of the line above the CTU row
  int cx, cy, cur_buf;           // cy: top of the CTU row being parsed
  // cu_qp_delta with one quantisation group per CTU (8.6.1): qp_cur is QpY of the CU being parsed
  // (the prediction until the CTU's delta is coded), reset to the slice QP at each row start (WPP)
  int qp_cur, delta_coded;
  int any_intra;                 // an intra CU was met in a P slice
  uint16_t *s_tu;                // [64] transform-tree scratch of the CU being parsed (see parse_cu)
};

__device__ __forceinline__ CuInfo load_cu(const ParseCtx &pc, int x, int y)
{
  const uint32_t *s;
  if (y < pc.cy) {
    s = pc.s_above + 4 * (((x - pc.cx) >> 3) + 1);
  } else {
    const int buf = x < pc.cx ? pc.cur_buf ^ 1 : pc.cur_buf;
    s = pc.s_ctu + 4 * (buf * 64 + (((y - pc.cy) >> 3) << 3) + ((x >> 3) & 7));
  }
  // fields unpacked with shifts: punning the words into the struct would send it through local memory
  const uint32_t w0 = s[0], w1 = s[1], w2 = s[2], w3 = s[3];
  CuInfo c;
  c.mvx = (int16_t)(w0 & 0xffff); c.mvy = (int16_t)(w0 >> 16);
  c.log2_size = (uint8_t)w1; c.pred_mode = (uint8_t)(w1 >> 8); c.intra_mode = (uint8_t)(w1 >> 16); c.cbf = (uint8_t)(w1 >> 24);
  c.skip = (uint8_t)w2; c.merge_idx = (uint8_t)(w2 >> 8); c.mvp_idx = (uint8_t)(w2 >> 16); c.qp = (uint8_t)(w2 >> 24);
  c.ref_idx = (uint8_t)w3; c.chroma_mode = (uint8_t)(w3 >> 8); c.tu_log2 = (uint8_t)(w3 >> 16); c.flags = (uint8_t)(w3 >> 24);
  return c;
}

__device__ __forceinline__ unsigned coding_order_p(const FrameParams &fp, int x, int y)
{
  return (unsigned)((y >> kCtbLog2) * fp.ctb_cols + (x >> kCtbLog2)) * 64u + (unsigned)xy_to_z((x >> 3) & 7, (y >> 3) & 7);
}

struct NbP { bool ok; int mvx, mvy, ref; };
__device__ __forceinline__ NbP nb_p(const ParseCtx &pc, unsigned cur, int xn, int yn)
{
  NbP n{false, 0, 0, 0};
  const FrameParams &fp = pc.fp;
  if (xn < 0 || yn < 0 || xn >= fp.w || yn >= fp.h) return n;
  if (coding_order_p(fp, xn, yn) >= cur) return n;
  CuInfo c = load_cu(pc, xn, yn);
  if (c.pred_mode != 0) return n;
  n.ok = true; n.mvx = c.mvx; n.mvy = c.mvy; n.ref = c.ref_idx;
  return n;
}
__device__ __forceinline__ bool same_p(const NbP &a, const NbP &b) { return a.mvx == b.mvx && a.mvy == b.mvy && a.ref == b.ref; }

// 8.5.3.2.7 / 8.5.3.2.8: a vector that spans td pictures rescaled to span tb pictures
__device__ __forceinline__ int scale_mv_d(int mv, int td, int tb)
{
  td = clip3(-128, 127, td); tb = clip3(-128, 127, tb);
  if (td == 0) return mv;                              // (a stream cannot name the current picture; keeps the division safe)
  const int tx = (16384 + (abs(td) >> 1)) / td;
  const int dsf = clip3(-4096, 4095, (tb * tx + 32) >> 6);
  const int p = dsf * mv;
  return clip3(-32768, 32767, (p < 0 ? -1 : 1) * ((abs(p) + 127) >> 8));
}
// a neighbour's vector as a predictor for reference index `ref`: as it is when it points to the same
// picture, else rescaled by the ratio of the POC distances
__device__ __forceinline__ void nb_for_ref(const FrameParams &fp, const NbP &n, int ref, int &mx, int &my)
{
  const int td = fp.ref_dist[n.ref & 15], tb = fp.ref_dist[ref & 15];
  mx = n.mvx; my = n.mvy;
  if (td != tb) { mx = scale_mv_d(mx, td, tb); my = scale_mv_d(my, td, tb); }
}

// 8.5.3.2.8 temporal luma motion vector prediction for reference index `ref`: bottom-right candidate
// (same CTB row, inside the picture), then the centre, from the collocated picture's 16x16 motion field
__device__ __forceinline__ bool temporal_mv_d(const FrameParams &fp, int x0, int y0, int n, int ref, int &mx, int &my)
{
  if (!fp.col_mvf) return false;
  const int w16 = (fp.w + 15) >> 4;
  for (int k = 0; k < 2; k++) {
    const int x = k ? x0 + (n >> 1) : x0 + n, y = k ? y0 + (n >> 1) : y0 + n;
    if (k == 0 && ((y0 >> kCtbLog2) != (y >> kCtbLog2) || y >= fp.h || x >= fp.w)) continue;
    const uint2 raw = __ldg((const uint2 *)(fp.col_mvf + (size_t)(y >> 4) * w16 + (x >> 4)));
    if (!((raw.y >> 16) & 0xff)) continue;             // intra or outside
    const int cmx = (int16_t)(raw.x & 0xffff), cmy = (int16_t)(raw.x >> 16), col_diff = (int16_t)(raw.y & 0xffff);
    const int cur_diff = fp.ref_dist[ref & 15];
    mx = cmx; my = cmy;
    if (col_diff != cur_diff) { mx = scale_mv_d(cmx, col_diff, cur_diff); my = scale_mv_d(cmy, col_diff, cur_diff); }
    return true;
  }
  return false;
}

// residual_coding (7.3.8.11): decoded levels go straight to the picture-shaped level plane
__device__ void parse_residual(Reader &r, const ParseCtx &pc, int16_t *plane, int pw, int x0, int y0, int log2n, int cidx,
                               int scan_idx)
{
  const int sb_log2 = log2n - 2, sbw = 1 << sb_log2;
  int offset, shift;
  if (cidx == 0) { offset = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; }
  else { offset = 15; shift = log2n - 2; }
  const int cmax = (log2n << 1) - 1;
  int pre[2];
  for (int d = 0; d < 2; d++) {
    int base = d ? CTX_LAST_Y : CTX_LAST_X, k = 0;
    while (k < cmax && dec_bin(r, base + offset + (k >> shift))) k++;
    pre[d] = k;
  }
  int last[2];
  for (int d = 0; d < 2; d++) {
    int g = pre[d];
    if (g > 3) {
      int nb = (g >> 1) - 1;
      last[d] = ((2 + (g & 1)) << nb) + (int)dec_bypass_bits(r, nb);
    } else {
      last[d] = g;
    }
  }
  int last_x = last[0], last_y = last[1];
  if (scan_idx == 2) { int t = last_x; last_x = last_y; last_y = t; }
  const int last_sb = scan_index_d(scan_idx, sb_log2, last_x >> 2, last_y >> 2);
  const int last_pos = scan_index_d(scan_idx, 2, last_x & 3, last_y & 3);
  unsigned long long csbf = 0;
  int c1 = 1;
  for (int i = last_sb; i >= 0; i--) {
    int xs, ys;
    scan_pos_d(scan_idx, sb_log2, i, xs, ys);
    int right = xs + 1 < sbw ? (int)((csbf >> (ys * 8 + xs + 1)) & 1) : 0;
    int below = ys + 1 < sbw ? (int)((csbf >> ((ys + 1) * 8 + xs)) & 1) : 0;
    int prev_csbf = right | (below << 1);
    int coded = 1, infer_dc = 0;
    if (i < last_sb && i > 0) {
      coded = dec_bin(r, CTX_CSBF + (prev_csbf ? 1 : 0) + (cidx ? 2 : 0));
      infer_dc = 1;
    }
    if (!coded) continue;
    csbf |= 1ull << (ys * 8 + xs);
    unsigned sig = 0;
    int start = 15;
    if (i == last_sb) { sig = 1u << last_pos; start = last_pos - 1; }
    // context of every sig_coeff_flag of this sub-block: base + a pattern looked up by scan position
    const int sig_base = CTX_SIG + (cidx ? 27 : 0);
    int sub_base;
    if (cidx == 0) sub_base = ((xs || ys) ? 3 : 0) + (log2n == 3 ? (scan_idx == 0 ? 9 : 15) : 21);
    else sub_base = log2n == 3 ? 9 : 12;
    const uint32_t pat = c_sig_pat[scan_idx][prev_csbf];
    const unsigned long long pat4 = c_sig_pat4[scan_idx];
    for (int p = start; p >= 0; p--) {
      if (p == 0 && infer_dc) { sig |= 1u; break; }
      int sctx;
      if (log2n == 2) sctx = (int)((pat4 >> (4 * p)) & 15);
      else if (p == 0 && i == 0) sctx = 0;                               // the DC coefficient
      else sctx = sub_base + (int)((pat >> (2 * p)) & 3);
      if (dec_bin(r, sig_base + sctx)) { sig |= 1u << p; infer_dc = 0; }
    }
    if (!sig) continue;
    int ctx_set = (i > 0 && cidx == 0) ? 2 : 0;
    if (c1 == 0) ctx_set++;
    c1 = 1;
    // The three passes below visit the significant positions only, from the highest scan position
    // down (sparse sub-blocks are the common case: looping over all 16 positions cost more than
    // the bins themselves).
    int num_g1 = 0, first_g1 = -1;
    unsigned g1 = 0;
    for (unsigned t = sig; t && num_g1 < 8; num_g1++) {
      const int p = 31 - __clz(t);
      t &= ~(1u << p);
      int f = dec_bin(r, CTX_GT1 + (cidx ? 16 : 0) + 4 * ctx_set + c1);
      if (f) { g1 |= 1u << p; c1 = 0; if (first_g1 < 0) first_g1 = p; }
      else if (c1 < 3 && c1 > 0) c1++;
    }
    int g2 = 0;
    if (first_g1 >= 0) g2 = dec_bin(r, CTX_GT2 + (cidx ? 4 : 0) + ctx_set);
    // sign data hiding (7.3.8.11): the sign of the first coefficient in scan order is not coded when the
    // significant coefficients of the group span more than three positions; it is the parity of the sum
    const int first_sig = __ffs(sig) - 1;
    const bool hidden = pc.fp.sign_hiding && (31 - __clz(sig)) - first_sig > 3;
    unsigned neg = 0;
    for (unsigned t = hidden ? sig & (sig - 1) : sig; t;) {
      const int p = 31 - __clz(t);
      t &= ~(1u << p);
      neg |= (unsigned)dec_bypass(r) << p;
    }
    int num_sig = 0, rice = 0, sum_abs = 0;
    for (unsigned t = sig; t;) {
      const int p = 31 - __clz(t);
      t &= ~(1u << p);
      int base = 1 + (num_sig < 8 ? (int)((g1 >> p) & 1) : 0) + (p == first_g1 ? g2 : 0);
      int thresh = num_sig < 8 ? (p == first_g1 ? 3 : 2) : 1;
      int absv = base;
      if (base == thresh) {
        // coeff_abs_level_remaining: TR prefix (<= 4 ones) + rice suffix, or EG(rice+1) escape
        int q = 0;
        while (q < 4 && dec_bypass(r)) q++;
        int rem;
        if (q < 4) {
          rem = (q << rice) + (int)dec_bypass_bits(r, rice);
        } else {
          int k = rice + 1, v = 0;
          while (k < 32 && dec_bypass(r)) { v += 1 << k; k++; }
          if (k >= 32) { r.err = 1; return; }
          v += (int)dec_bypass_bits(r, k);
          rem = (4 << rice) + v;
        }
        absv = base + rem;
        if (absv > (3 << rice)) rice = min(rice + 1, 4);
      }
      num_sig++;
      sum_abs += absv;
      int xp, yp;
      scan_pos_d(scan_idx, 2, p, xp, yp);
      {                                   // all lanes store the same value to the same address
        int v = min(absv, 32767);
        const bool negative = (hidden && p == first_sig) ? (sum_abs & 1) != 0 : ((neg >> p) & 1) != 0;
        plane[(size_t)(y0 + ys * 4 + yp) * pw + x0 + xs * 4 + xp] = (int16_t)(negative ? -v : v);
      }
    }
  }
}

__device__ __forceinline__ int scan_idx_for_d(int pred_mode, int intra_mode, int log2n, int cidx)
{
  if (pred_mode != 1) return 0;
  if (!(log2n == 2 || (log2n == 3 && cidx == 0))) return 0;
  if (intra_mode >= 6 && intra_mode <= 14) return 2;
  if (intra_mode >= 22 && intra_mode <= 30) return 1;
  return 0;
}

// luma intra prediction mode of the 4x4 block that covers luma sample (x, y) of a CU of the cu map
// (inter CUs carry mode 1, DC, as the candidate derivation of 8.4.2 asks)
__device__ __forceinline__ int intra_mode_of(const CuInfo &c, int x, int y)
{
  if (!(c.flags & 1) || c.pred_mode != 1) return c.intra_mode;
  const int part = (((y >> 2) & 1) << 1) | ((x >> 2) & 1);
  return part == 0 ? c.intra_mode : (part == 1 ? (c.mvx & 0xff) : (part == 2 ? ((c.mvx >> 8) & 0xff) : (c.mvy & 0xff)));
}

// cu_qp_delta_abs (9.3.3.10: prefix TR cMax 5, ctx 0 then ctx 1; suffix EG0) and sign, once per quantisation group
__device__ __forceinline__ void parse_cu_qp_delta(Reader &r, ParseCtx &pc)
{
  int a = 0;
  while (a < 5 && dec_bin(r, CTX_CU_QP_DELTA + (a ? 1 : 0))) a++;
  if (a == 5) {
    int k = 0, v = 0;
    while (k < 8 && dec_bypass(r)) { v += 1 << k; k++; }
    if (k >= 8) { r.err = 11; return; }
    a += v + (int)dec_bypass_bits(r, k);
  }
  const int d = (a && dec_bypass(r)) ? -a : a;
  if (d < -26 || d > 25) { r.err = 11; return; }
  pc.qp_cur = (pc.qp_cur + d + 52) % 52;
  pc.delta_coded = 1;
}

__device__ void parse_cu(Reader &r, ParseCtx &pc, int x0, int y0, int log2)
{
  const FrameParams &fp = pc.fp;
  const int n = 1 << log2;
  CuInfo cu;
  cu.mvx = 0; cu.mvy = 0; cu.log2_size = (uint8_t)log2; cu.pred_mode = 0; cu.intra_mode = 1; cu.cbf = 0;
  cu.skip = 0; cu.merge_idx = 0xff; cu.mvp_idx = 0; cu.qp = 0;
  cu.ref_idx = 0; cu.chroma_mode = 1; cu.tu_log2 = (uint8_t)min(log2, 5); cu.flags = 0;
  bool tu = false;
  bool intra = fp.is_idr != 0;
  if (!fp.is_idr) {
    int ctx = 0;
    if (x0 > 0) ctx += load_cu(pc, x0 - 1, y0).skip;
    if (y0 > 0) ctx += load_cu(pc, x0, y0 - 1).skip;
    cu.skip = (uint8_t)dec_bin(r, CTX_SKIP + ctx);
    int merge = cu.skip;
    if (!cu.skip) {
      intra = dec_bin(r, CTX_PRED_MODE) != 0;                          // intra CU in a P slice
      if (!intra) {
        if (!dec_bin(r, CTX_PART_MODE)) { r.err = 3; return; }         // only PART_2Nx2N
        merge = dec_bin(r, CTX_MERGE_FLAG);
      }
    }
    if (!intra) {
    const unsigned cur = coding_order_p(fp, x0, y0);
    NbP a1 = nb_p(pc, cur, x0 - 1, y0 + n - 1), b1 = nb_p(pc, cur, x0 + n - 1, y0 - 1);
    NbP b0 = nb_p(pc, cur, x0 + n, y0 - 1), a0 = nb_p(pc, cur, x0 - 1, y0 + n), b2 = nb_p(pc, cur, x0 - 1, y0 - 1);
    if (merge) {
      int midx = 0;
      if (fp.max_merge > 1 && dec_bin(r, CTX_MERGE_IDX)) {
        midx = 1;
        while (midx < fp.max_merge - 1 && dec_bypass(r)) midx++;
      }
      // merge candidates (8.5.3.2.2-5): spatial, temporal (reference index 0), zero candidates that
      // walk the reference indices
      // (the list is not stored: the candidate whose position equals merge_idx is kept as it goes by --
      // an array indexed by merge_idx would live in local memory)
      int cnt = 0, sel_x = 0, sel_y = 0, sel_ref = 0;
      bool use_b1 = b1.ok && !(a1.ok && same_p(b1, a1));
      bool use_b0 = b0.ok && !(b1.ok && same_p(b0, b1));
      bool use_a0 = a0.ok && !(a1.ok && same_p(a0, a1));
      bool use_b2 = b2.ok && !(a1.ok && same_p(b2, a1)) && !(b1.ok && same_p(b2, b1));
#define MERGE_PUSH(X, Y, R) do { if (cnt == midx) { sel_x = (X); sel_y = (Y); sel_ref = (R); } cnt++; } while (0)
      if (a1.ok) MERGE_PUSH(a1.mvx, a1.mvy, a1.ref);
      if (use_b1) MERGE_PUSH(b1.mvx, b1.mvy, b1.ref);
      if (use_b0) MERGE_PUSH(b0.mvx, b0.mvy, b0.ref);
      if (use_a0) MERGE_PUSH(a0.mvx, a0.mvy, a0.ref);
      if (use_b2 && cnt < 4) MERGE_PUSH(b2.mvx, b2.mvy, b2.ref);
      if (cnt <= midx) {                                  // not among the spatial candidates: temporal, then zero candidates
        int tx = 0, ty = 0;
        if (cnt < kMaxMerge && temporal_mv_d(fp, x0, y0, n, 0, tx, ty)) MERGE_PUSH(tx, ty, 0);
        for (int zero_idx = 0; cnt <= midx; zero_idx++) MERGE_PUSH(0, 0, zero_idx < fp.n_refs ? zero_idx : 0);
      }
#undef MERGE_PUSH
      cu.mvx = (int16_t)sel_x; cu.mvy = (int16_t)sel_y; cu.ref_idx = (uint8_t)sel_ref;
      cu.merge_idx = (uint8_t)midx;
      tu = !cu.skip;
    } else {
      // ref_idx_l0 (TR, cMax n_refs - 1: two context-coded bins, then bypass)
      int ref = 0;
      if (fp.n_refs > 1) {
        const int cmax = fp.n_refs - 1;
        while (ref < cmax && (ref < 2 ? dec_bin(r, CTX_REF_IDX + ref) : dec_bypass(r))) ref++;
      }
      cu.ref_idx = (uint8_t)ref;
      // mvd_coding (7.3.8.9)
      int gt0[2], gt1[2] = {0, 0}, mvd[2] = {0, 0};
      gt0[0] = dec_bin(r, CTX_MVD_GT0);
      gt0[1] = dec_bin(r, CTX_MVD_GT0);
      if (gt0[0]) gt1[0] = dec_bin(r, CTX_MVD_GT1);
      if (gt0[1]) gt1[1] = dec_bin(r, CTX_MVD_GT1);
      for (int k = 0; k < 2; k++) {
        if (!gt0[k]) continue;
        int a = 1;
        if (gt1[k]) {
          int kk = 1, v = 0;
          while (kk < 32 && dec_bypass(r)) { v += 1 << kk; kk++; }
          if (kk >= 32) { r.err = 4; return; }
          v += (int)dec_bypass_bits(r, kk);
          a = v + 2;
        }
        mvd[k] = dec_bypass(r) ? -a : a;
      }
      int pi = dec_bin(r, CTX_MVP_IDX);
      // AMVP (8.5.3.2.6-8): candidate A from (A0, A1), B from (B0, B1, B2) -- first a neighbour that
      // points to the same reference picture, then any neighbour, rescaled -- then temporal, then zero
      const int tdist = fp.ref_dist[ref & 15];
      const bool is_scaled = a0.ok || a1.ok;
      bool fa = false, fb = false;
      int ax = 0, ay = 0, bx = 0, by = 0;
      if (a0.ok && fp.ref_dist[a0.ref & 15] == tdist) { ax = a0.mvx; ay = a0.mvy; fa = true; }
      else if (a1.ok && fp.ref_dist[a1.ref & 15] == tdist) { ax = a1.mvx; ay = a1.mvy; fa = true; }
      else if (a0.ok) { nb_for_ref(fp, a0, ref, ax, ay); fa = true; }
      else if (a1.ok) { nb_for_ref(fp, a1, ref, ax, ay); fa = true; }
      if (b0.ok && fp.ref_dist[b0.ref & 15] == tdist) { bx = b0.mvx; by = b0.mvy; fb = true; }
      else if (b1.ok && fp.ref_dist[b1.ref & 15] == tdist) { bx = b1.mvx; by = b1.mvy; fb = true; }
      else if (b2.ok && fp.ref_dist[b2.ref & 15] == tdist) { bx = b2.mvx; by = b2.mvy; fb = true; }
      if (!is_scaled && fb) { ax = bx; ay = by; fa = true; }
      if (!is_scaled) {
        fb = false;
        if (b0.ok) { nb_for_ref(fp, b0, ref, bx, by); fb = true; }
        else if (b1.ok) { nb_for_ref(fp, b1, ref, bx, by); fb = true; }
        else if (b2.ok) { nb_for_ref(fp, b2, ref, bx, by); fb = true; }
      }
      int k = 0, pvx = 0, pvy = 0;                          // the predictor at position mvp_l0_flag, kept as it goes by
      if (fa) { if (k == pi) { pvx = ax; pvy = ay; } k++; }
      if (fb && !(fa && ax == bx && ay == by)) { if (k == pi) { pvx = bx; pvy = by; } k++; }
      if (k <= pi) {
        int tx = 0, ty = 0;
        if (temporal_mv_d(fp, x0, y0, n, ref, tx, ty)) { if (k == pi) { pvx = tx; pvy = ty; } k++; }
      }
      cu.mvx = (int16_t)(pvx + mvd[0]); cu.mvy = (int16_t)(pvy + mvd[1]);
      cu.mvp_idx = (uint8_t)pi;
      tu = dec_bin(r, CTX_RQT_ROOT_CBF) != 0;
    }
    pc.max_mv = max(pc.max_mv, max(abs((int)cu.mvx), abs((int)cu.mvy)));
    if (fp.mv_edges && !(mv_allowed(fp, x0, n, cu.mvx) && mv_allowed_v(fp, y0, n, cu.mvy))) { r.err = 12; return; }   // motion across an interior tile edge
    }
  }
  int part_mode_v[4] = {1, 1, 1, 1};                                     // luma modes of the prediction blocks
  if (intra) {
    cu.pred_mode = 1;
    // part_mode: only the smallest CUs choose between PART_2Nx2N and PART_NxN (four 4x4 prediction blocks)
    const bool nxn = log2 == 3 && !dec_bin(r, CTX_PART_MODE);
    const int parts = nxn ? 4 : 1;
    int prev[4];
    for (int b = 0; b < parts; b++) prev[b] = dec_bin(r, CTX_PREV_INTRA_LUMA);
    for (int b = 0; b < parts; b++) {
      const int xb = x0 + 4 * (b & 1), yb = y0 + 4 * (b >> 1);
      // candidate modes (8.4.2): inter neighbours carry intra_mode 1 (DC) in the cu map, as the rule asks;
      // the blocks of this CU itself are not in the map yet
      int ca = 1, cb = 1;
      if (b & 1) ca = part_mode_v[b - 1];
      else if (xb > 0) ca = intra_mode_of(load_cu(pc, xb - 1, yb), xb - 1, yb);
      if (b & 2) cb = part_mode_v[b - 2];
      else if (yb > 0 && (yb & (kCtb - 1))) cb = intra_mode_of(load_cu(pc, xb, yb - 1), xb, yb - 1);
      int cand[3];
      if (ca == cb) {
        if (ca < 2) { cand[0] = 0; cand[1] = 1; cand[2] = 26; }
        else { cand[0] = ca; cand[1] = 2 + ((ca + 29) % 32); cand[2] = 2 + ((ca - 2 + 1) % 32); }
      } else {
        cand[0] = ca; cand[1] = cb;
        cand[2] = (ca != 0 && cb != 0) ? 0 : ((ca != 1 && cb != 1) ? 1 : 26);
      }
      int mode;
      if (prev[b]) {
        int idx = 0;
        if (dec_bypass(r)) idx = dec_bypass(r) ? 2 : 1;
        mode = cand[idx];
      } else {
        if (cand[0] > cand[1]) { int t = cand[0]; cand[0] = cand[1]; cand[1] = t; }
        if (cand[0] > cand[2]) { int t = cand[0]; cand[0] = cand[2]; cand[2] = t; }
        if (cand[1] > cand[2]) { int t = cand[1]; cand[1] = cand[2]; cand[2] = t; }
        mode = (int)dec_bypass_bits(r, 5);
        for (int i = 0; i < 3; i++) if (mode >= cand[i]) mode++;
      }
      part_mode_v[b] = mode;
    }
    if (!nxn) part_mode_v[1] = part_mode_v[2] = part_mode_v[3] = part_mode_v[0];
    cu.intra_mode = (uint8_t)part_mode_v[0];
    if (nxn) {
      cu.flags = 1;
      cu.mvx = (int16_t)(part_mode_v[1] | (part_mode_v[2] << 8)); cu.mvy = (int16_t)part_mode_v[3];
    }
    // intra_chroma_pred_mode (8.4.3): derived from the luma mode (of the first block), or planar / vertical /
    // horizontal / DC, where the one that equals the luma mode stands for mode 34
    int cmode = part_mode_v[0];
    if (dec_bin(r, CTX_INTRA_CHROMA)) {
      const int k = (int)dec_bypass_bits(r, 2);
      const int m = k == 0 ? 0 : (k == 1 ? 26 : (k == 2 ? 10 : 1));
      cmode = m == part_mode_v[0] ? 34 : m;
    }
    cu.chroma_mode = (uint8_t)cmode;
    if (!fp.is_idr) pc.any_intra = 1;
    tu = true;
  }
  // ---- transform tree (7.3.8.8), depth first with a small explicit stack.  Per 8x8 unit of the CU the
  // size of the transform unit covering it and that unit's coded block flags go to pc.s_tu (raster
  // inside the CU, 8 units per row): cbf | tu_log2 << 8.
  const int n8 = n >> 3, units = n8 * n8;
  const size_t ysz = (size_t)fp.w * fp.h;
  for (int base = 0; base < units; base += 32) {               // branch-free: surplus lanes repeat the last unit
    const int u = min(base + pc.lane, units - 1);
    pc.s_tu[((u >> (log2 - 3)) << 3) + (u & (n8 - 1))] = (uint16_t)(min(log2, 5) << 8);
  }
  __syncwarp();
  int root = 0;
  if (tu) {
    const int nxn = cu.flags & 1;                                // IntraSplitFlag: the first split is inferred
    const int max_depth = (cu.pred_mode == 1 ? fp.tr_depth_intra : fp.tr_depth_inter) + nxn;
    struct Node { short x, y; signed char l2, depth, child, cb, cr; } st[5];
    int sp = 0;
    st[0] = {(short)x0, (short)y0, (signed char)log2, 0, -1, 1, 1};
    while (sp >= 0 && !r.err) {
      Node &nd = st[sp];
      if (nd.child < 0) {
        // node header: split_transform_flag (parsed or inferred), then the chroma flags of the node
        int split;
        if (nd.l2 <= 5 && nd.l2 > 2 && nd.depth < max_depth && !(nxn && nd.depth == 0)) split = dec_bin(r, CTX_SPLIT_TRANSFORM + 5 - nd.l2);
        else split = nd.l2 > 5 || (nxn && nd.depth == 0);      // larger than the largest transform block / NxN: inferred
        const int par_cb = nd.cb, par_cr = nd.cr;
        const int cb = par_cb ? dec_bin(r, CTX_CBF_CHROMA + nd.depth) : 0;
        const int cr = par_cr ? dec_bin(r, CTX_CBF_CHROMA + nd.depth) : 0;
        nd.cb = (signed char)cb; nd.cr = (signed char)cr;
        if (split && nd.l2 == 3) {
          // four 4x4 luma transform units (7.3.8.10 with log2TrafoSize 2); the two 4x4 chroma blocks of the
          // 8x8 node follow the fourth.  The unit keeps cbf bit 0 = any luma block, bits 4..7 = the blocks.
          int cbf = (cb << 1) | (cr << 2);
          for (int b = 0; b < 4 && !r.err; b++) {
            const int lu = dec_bin(r, CTX_CBF_LUMA);           // trafoDepth > 0
            if (fp.ctu_qp && (lu | cb | cr) && !pc.delta_coded) parse_cu_qp_delta(r, pc);
            if (!lu || r.err) continue;
            cbf |= 1 | (16 << b);
            parse_residual(r, pc, pc.levels, fp.w, nd.x + 4 * (b & 1), nd.y + 4 * (b >> 1), 2, 0,
                           scan_idx_for_d(cu.pred_mode, part_mode_v[b], 2, 0));
          }
          for (int k = 1; k < 3 && !r.err; k++)
            if ((cbf >> k) & 1)
              parse_residual(r, pc, pc.levels + ysz + (k == 2 ? ysz / 4 : 0), fp.w >> 1, nd.x >> 1, nd.y >> 1, 2, k,
                             scan_idx_for_d(cu.pred_mode, cu.chroma_mode, 2, k));
          root |= cbf;
          pc.s_tu[(((nd.y - y0) >> 3) << 3) + ((nd.x - x0) >> 3)] = (uint16_t)(cbf | (2 << 8));   // every lane: same value
          __syncwarp();
          sp--;
          continue;
        }
        if (split) { nd.child = 0; continue; }
        // leaf: transform_unit
        int lu = 1;
        if (cu.pred_mode == 1 || nd.depth != 0 || cb || cr) lu = dec_bin(r, CTX_CBF_LUMA + (nd.depth == 0 ? 1 : 0));
        const int cbf = lu | (cb << 1) | (cr << 2);
        root |= cbf;
        if (fp.ctu_qp && cbf && !pc.delta_coded) { parse_cu_qp_delta(r, pc); if (r.err) return; }
        {
          const int t8 = 1 << (nd.l2 - 3), tunits = t8 * t8;
          const int bx = (nd.x - x0) >> 3, by = (nd.y - y0) >> 3;
          for (int base = 0; base < tunits; base += 32) {
            const int u = min(base + pc.lane, tunits - 1);
            pc.s_tu[((by + (u >> (nd.l2 - 3))) << 3) + bx + (u & (t8 - 1))] = (uint16_t)(cbf | (nd.l2 << 8));
          }
          __syncwarp();
        }
        for (int k = 0; k < 3 && !r.err; k++) {
          if (!((cbf >> k) & 1)) continue;
          const int sft = k ? 1 : 0;
          int16_t *plane = pc.levels + (k == 0 ? 0 : ysz + (k == 2 ? ysz / 4 : 0));
          parse_residual(r, pc, plane, fp.w >> sft, nd.x >> sft, nd.y >> sft, nd.l2 - sft, k,
                         scan_idx_for_d(cu.pred_mode, k ? cu.chroma_mode : part_mode_v[0], nd.l2 - sft, k));
        }
        sp--;
        continue;
      }
      if (nd.child < 4) {
        const int q = nd.child++, h = 1 << (nd.l2 - 1);
        const short cx2 = (short)(nd.x + (q & 1) * h), cy2 = (short)(nd.y + (q >> 1) * h);
        const signed char l2 = (signed char)(nd.l2 - 1), dp = (signed char)(nd.depth + 1), pcb = nd.cb, pcr = nd.cr;
        sp++;
        st[sp] = {cx2, cy2, l2, dp, -1, pcb, pcr};
      } else {
        sp--;
      }
    }
    if (r.err) return;
  }
  if (fp.ctu_qp) cu.qp = (uint8_t)pc.qp_cur;
  cu.flags |= (uint8_t)(root ? 2 : 0);                       // bit 1: the CU has a coded residual
  // publish the cu map entries: shared memory for the CUs that follow in this row, global memory for
  // the row below and the reconstruction kernels (lane u takes unit u of the CU); cbf and the
  // transform unit size are per unit
  {
    // the entry's four words, packed with shifts (no pointer punning: the struct stays in registers)
    uint32_t s[4];
    s[0] = (uint32_t)(uint16_t)cu.mvx | ((uint32_t)(uint16_t)cu.mvy << 16);
    s[1] = (uint32_t)cu.log2_size | ((uint32_t)cu.pred_mode << 8) | ((uint32_t)cu.intra_mode << 16) | ((uint32_t)cu.cbf << 24);
    s[2] = (uint32_t)cu.skip | ((uint32_t)cu.merge_idx << 8) | ((uint32_t)cu.mvp_idx << 16) | ((uint32_t)cu.qp << 24);
    s[3] = (uint32_t)cu.ref_idx | ((uint32_t)cu.chroma_mode << 8) | ((uint32_t)cu.tu_log2 << 16) | ((uint32_t)cu.flags << 24);
    __syncwarp();
    // Branch-free on purpose: a lane-dependent branch here left the warp split into two groups that
    // ran the rest of the (lane-redundant) parse one after the other, doubling the time.  Lanes with
    // no unit of their own repeat the store of unit 0 (same address, same value).
    for (int base = 0; base < units; base += 32) {
      int u = base + pc.lane;
      int i = u & (n8 - 1), j = u >> (log2 - 3);
      const bool mine = u < units && x0 + 8 * i < fp.w && y0 + 8 * j < fp.h;
      i = mine ? i : 0; j = mine ? j : 0;
      const unsigned tuv = pc.s_tu[(j << 3) + i];
      // CuInfo bytes: word 1 = log2_size | pred_mode << 8 | intra_mode << 16 | cbf << 24; word 3 = ref_idx | chroma_mode << 8 | tu_log2 << 16 | flags << 24
      const uint32_t w1 = (s[1] & 0x00ffffffu) | ((tuv & 0xffu) << 24);
      const uint32_t w3 = (s[3] & 0xff00ffffu) | ((tuv >> 8) << 16);
      uint32_t *t = pc.s_ctu + 4 * (pc.cur_buf * 64 + ((((y0 - pc.cy) >> 3) + j) << 3) + ((x0 >> 3) & 7) + i);
      t[0] = s[0]; t[1] = w1; t[2] = s[2]; t[3] = w3;
      uint32_t *d = (uint32_t *)(pc.cu + (size_t)((y0 >> 3) + j) * fp.w8 + (x0 >> 3) + i);
      __stcg(d, s[0]); __stcg(d + 1, w1); __stcg(d + 2, s[2]); __stcg(d + 3, w3);
    }
    __syncwarp();
  }
}

// sao() of one CTU (7.3.8.3) -> fp.sao[ctu].  `left`: the previous CTU's parameters (sao_merge_left_flag).
__device__ void parse_sao(Reader &r, const FrameParams &fp, int row, int col, SaoCtu &left)
{
  SaoCtu p;
  bool merged = false;
  if (col > 0 && dec_bin(r, CTX_SAO_MERGE)) { p = left; merged = true; }
  if (!merged && row > 0 && dec_bin(r, CTX_SAO_MERGE)) {
    // the CTU above was completed at least two CTUs ago (WPP progress wait) or by this warp (no_wpp)
    const uint32_t *s = (const uint32_t *)(fp.sao + (size_t)(row - 1) * fp.ctb_cols + col);
    uint32_t *d = (uint32_t *)&p;
    for (int i = 0; i < 5; i++) d[i] = __ldcg(s + i);
    merged = true;
  }
  if (!merged) {
    uint32_t *z = (uint32_t *)&p;
    for (int i = 0; i < 5; i++) z[i] = 0;
    for (int comp = 0; comp < 3; comp++) {
      const int g = comp ? 1 : 0;
      if (!((fp.sao_flags >> g) & 1)) continue;
      if (comp < 2) {
        int type = 0;
        if (dec_bin(r, CTX_SAO_TYPE)) type = dec_bypass(r) ? 2 : 1;
        p.type[g] = (uint8_t)type;
      }
      const int type = p.type[g];
      if (!type) continue;
      int a[4];
      for (int k = 0; k < 4; k++) {
        int v = 0;
        while (v < 7 && dec_bypass(r)) v++;
        a[k] = v;
      }
      if (type == 1) {
        for (int k = 0; k < 4; k++)
          if (a[k] && dec_bypass(r)) a[k] = -a[k];
        p.band_pos[comp] = (uint8_t)dec_bypass_bits(r, 5);
      } else {
        if (comp == 0) p.eo_class[0] = (uint8_t)dec_bypass_bits(r, 2);
        if (comp == 1) p.eo_class[1] = (uint8_t)dec_bypass_bits(r, 2);
        a[2] = -a[2]; a[3] = -a[3];                    // categories 3 and 4 (local maxima) take negative offsets
      }
      for (int k = 0; k < 4; k++) p.offset[comp][k] = (int8_t)a[k];
    }
  }
  left = p;
  {                                                    // every lane stores the same five words
    const uint32_t *s = (const uint32_t *)&p;
    uint32_t *d = (uint32_t *)(fp.sao + (size_t)row * fp.ctb_cols + col);
    for (int i = 0; i < 5; i++) __stcg(d + i, s[i]);
  }
}

// bases[r] = byte offset of substream r inside `data`, bases[rows] = end.  status[0] receives the
// first error code (0 = ok), status[1] the largest |mv| component.
__global__ void __launch_bounds__(32)
k_parse_rows(const __grid_constant__ FrameParams fp, const uint8_t *__restrict__ data, const uint32_t *__restrict__ bases, CuInfo *cu,
             int16_t *levels, uint8_t *sync_ctx, int *sync_flag, int *progress, int *status)
{
  __shared__ uint2 s_ctx[CTX_COUNT + 1];
  __shared__ uint2 s_next[128 * 2];
  __shared__ uint32_t s_ctu[2 * 64 * 4];
  __shared__ uint32_t s_above[10 * 4];
  // WPP: one substream per CTU row (row_first == row_last == blockIdx.x).  no_wpp (a tile without
  // entropy_coding_sync): one warp reads every row from a single substream; bases[0..1] bound it.
  const int lane = threadIdx.x;
  const int row_first = fp.no_wpp ? 0 : blockIdx.x, row_last = fp.no_wpp ? fp.ctb_rows - 1 : blockIdx.x;
  int row = row_first;
  for (int i = lane; i < 128; i += 32) {                  // i = (pStateIdx << 1) | valMps
    const uint32_t st = (uint32_t)i >> 1, mps = (uint32_t)i & 1;
    s_next[2 * i] = ctx_entry((min(st + 1, 62u) << 1) | mps);                                  // after an MPS
    s_next[2 * i + 1] = ctx_entry(((uint32_t)c_trans_lps[min(st, 63u)] << 1) | (mps ^ (st == 0 ? 1u : 0u)));   // after an LPS
  }
  if (lane == 0) s_ctx[CTX_COUNT] = make_uint2(0, 0);
  Reader r;
  r.p = data + bases[fp.no_wpp ? 0 : row]; r.end = data + bases[fp.no_wpp ? 1 : row + 1]; r.ctx = s_ctx; r.next = s_next; r.err = 0; r.lane = lane;
  r.pend = make_uint2(0, 0); r.pend_idx = CTX_COUNT;
  __shared__ uint16_t s_tu[64];
  ParseCtx pc{fp, cu, levels, lane, 0, s_ctu, s_above, 0, row * kCtb, 0, fp.qp, 0, 0, s_tu};
  SaoCtu sao_left;
  { uint32_t *z = (uint32_t *)&sao_left; for (int i = 0; i < 5; i++) z[i] = 0; }
  if (row == 0 || fp.ctb_cols < 2) {
    init_contexts_d(r.ctx, fp.init_type, fp.qp, lane);
  } else {
    // All lanes poll (one broadcast load per iteration).  Nothing in this kernel branches on the
    // lane index: a leader-only branch followed by __syncwarp() has left the warp split into groups
    // that then ran the lane-redundant parser one after the other (measured: 16 instead of 32
    // threads per instruction, twice the time).
    {
      volatile int *f = sync_flag;
      while (f[row - 1] == 0) __nanosleep(100);
      __threadfence();
    }
    for (int i = lane; i < CTX_COUNT; i += 32) r.ctx[i] = ctx_entry(__ldcg(sync_ctx + (size_t)(row - 1) * CTX_COUNT + i));
  }
  __syncwarp();
  reader_start(r);
  for (; row <= row_last && !r.err; row++) {
    pc.cy = row * kCtb;
    for (int col = 0; col < fp.ctb_cols && !r.err; col++) {
      if (row > 0) {                       // above and above-right CTUs must be parsed (cu map reads)
        if (!fp.no_wpp) {                  // (one warp does all rows without WPP: nothing to wait for)
          volatile int *p = progress;
          const int need = min(col + 2, fp.ctb_cols);
          for (;;) {
            const int have = p[row - 1];
            if (have >= need || have < 0) break;
            __nanosleep(64);
          }
          __threadfence();
        }
        {
          const int slot = min(lane, 9);
          const int ux = min(max(col * 8 - 1 + slot, 0), fp.w8 - 1);
          const uint32_t *s = (const uint32_t *)(cu + (size_t)(row * 8 - 1) * fp.w8 + ux);
          uint32_t *d = s_above + 4 * slot;
          d[0] = __ldcg(s); d[1] = __ldcg(s + 1); d[2] = __ldcg(s + 2); d[3] = __ldcg(s + 3);
        }
      }
      const int cx = col * kCtb, cy = row * kCtb;
      pc.cx = cx; pc.cur_buf = col & 1;
      pc.delta_coded = 0;                  // new quantisation group; qp_cur carries over as qPY_PREV
      __syncwarp();
      if (fp.sao_flags) parse_sao(r, fp, row, col, sao_left);
      for (int z = 0; z < 64 && !r.err;) {
        int x0 = cx + 8 * z_to_x(z), y0 = cy + 8 * z_to_y(z);
        if (x0 >= fp.w || y0 >= fp.h) { z++; continue; }
        int log2 = 3;
        for (int L = 6; L > 3; L--) {
          if (z & ((1 << (2 * (L - 3))) - 1)) continue;
          const int nn = 1 << L;
          int split;
          if (x0 + nn <= fp.w && y0 + nn <= fp.h) {
            int ctx = 0, depth = 6 - L;
            if (x0 > 0) ctx += (kCtbLog2 - load_cu(pc, x0 - 1, y0).log2_size) > depth;
            if (y0 > 0) ctx += (kCtbLog2 - load_cu(pc, x0, y0 - 1).log2_size) > depth;
            split = dec_bin(r, CTX_SPLIT_CU + ctx);
          } else {
            split = 1;
          }
          if (!split) { log2 = L; break; }
        }
        parse_cu(r, pc, x0, y0, log2);
        z += 1 << (2 * (log2 - 3));
      }
      if (col == 1 && row + 1 < fp.ctb_rows && !fp.no_wpp) {
        // every lane holds the same table; all store it (same addresses, same values)
        ctx_flush(r);
        for (int i = lane; i < CTX_COUNT; i += 32) sync_ctx[(size_t)row * CTX_COUNT + i] = (uint8_t)r.ctx[i].y;
        __threadfence();
        __syncwarp();
        *(volatile int *)&sync_flag[row] = 1;
      }
      if (fp.ctu_qp) fp.ctu_qp[row * fp.ctb_cols + col] = (uint8_t)pc.qp_cur;     // what the CTU's residuals are scaled with
      const bool last = row == fp.ctb_rows - 1 && col == fp.ctb_cols - 1 && !fp.more_tiles;
      const bool end_sub = col == fp.ctb_cols - 1 && (!fp.no_wpp || row == fp.ctb_rows - 1);
      int eos = dec_terminate(r);                                        // end_of_slice_segment_flag
      if (eos != (last ? 1 : 0)) r.err = 8;
      if (end_sub && !last && !dec_terminate(r)) r.err = 9;              // end_of_subset_one_bit
      __threadfence();
      __syncwarp();                        // every lane's cu map stores are fenced before any lane publishes
      *(volatile int *)&progress[row] = r.err ? -1 : col + 1;
    }
  }                                      // rows of this substream
  if (lane == 0) {
    if (r.err) {
      atomicCAS(&status[0], 0, r.err);
      for (int i = row_first; i <= row_last; i++) {
        atomicExch(&progress[i], -1);          // release the rows below
        atomicExch(&sync_flag[i], 1);
      }
    }
    atomicMax(&status[1], pc.max_mv);
    if (pc.any_intra && fp.any_intra) *fp.any_intra = 1;
  }
}

}  // namespace

// Motion field of the parsed picture at 16x16 granularity, for the temporal candidates of later pictures
__global__ void __launch_bounds__(256)
k_store_mvf(FrameParams fp, const CuInfo *__restrict__ cu, MvField *__restrict__ out)
{
  const int w16 = (fp.w + 15) >> 4, h16 = (fp.h + 15) >> 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w16 * h16) return;
  const int y16 = i / w16, x16 = i - y16 * w16;
  const CuInfo c = cu[(size_t)(2 * y16) * fp.w8 + 2 * x16];
  MvField m;
  m.inter = !fp.is_idr && c.pred_mode == 0 && c.log2_size >= 3;
  m.mvx = c.mvx; m.mvy = c.mvy; m.poc_diff = m.inter ? fp.ref_dist[c.ref_idx & 15] : 0; m.pad = 0;
  out[i] = m;
}

cudaError_t launch_store_mvf(const FrameParams &fp, const CuInfo *cu, MvField *out, cudaStream_t s)
{
  const int n = ((fp.w + 15) >> 4) * ((fp.h + 15) >> 4);
  k_store_mvf<<<(n + 255) / 256, 256, 0, s>>>(fp, cu, out);
  return cudaGetLastError();
}

cudaError_t launch_parse(const FrameParams &fp, const uint8_t *data, const uint32_t *bases, CuInfo *cu, int16_t *levels,
                         uint8_t *sync_ctx, int *sync_flag, int *progress, int *status, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(sync_flag, 0, sizeof(int) * fp.ctb_rows, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(progress, 0, sizeof(int) * fp.ctb_rows, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(status, 0, sizeof(int) * 2, s);
  if (e != cudaSuccess) return e;
  k_parse_rows<<<fp.no_wpp ? 1 : fp.ctb_rows, 32, 0, s>>>(fp, data, bases, cu, levels, sync_ctx, sync_flag, progress, status);
  return cudaGetLastError();
}

}  // namespace b200
