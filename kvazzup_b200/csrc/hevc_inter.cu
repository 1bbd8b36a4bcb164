// P-picture kernels of the B200 HEVC encoder (sm_100a).
//
//   k_me_ctu        one CTA per CTU: exhaustive full-sample search at 8x8 granularity with the
//                   SADs summed to 16x16 / 32x32 in registers (warp shuffles), bottom-up
//                   partition decision, half- then quarter-sample refinement.  Source block in
//                   registers, reference window staged in shared memory, SAD by VABSDIFF4,
//                   interpolation by DP4A.  Replaces Kvazaar's search_inter / kvz_sad /
//                   kvz_filter_inter path.
//   k_inter_recon   one CTA per CTU: motion compensation (luma 8-tap, chroma 4-tap), residual,
//                   forward DCT, quantisation, dequantisation, inverse DCT, reconstruction
//                   (rows K3, K5, K6).
//   k_inter_modes   one thread per 8x8 unit: merge candidates, skip decision, AMVP predictor
//                   choice from the final motion field (row K10) -- keeps the entropy coder
//                   free of any decision logic.
#include <algorithm>

#include "hevc_intra.cuh"
#include "hevc_kernels.h"

namespace b200 {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int window_margin(int range) { return (range + 4 + 3) & ~3; }

struct MeShared {
  unsigned key8[64], key16[16], key32[4];
  unsigned eff16[16];
  uint8_t use16[16], use32[4];
  uint8_t org[64], log2[64];       // per unit (z-order): origin unit of its CU, CU log2 size
  short mvx[64], mvy[64];          // per origin unit: best mv so far (quarter samples)
  short cmx[64], cmy[64];          // per origin unit: centre of the current refinement step
  unsigned best[64];               // per origin unit: best cost so far
  unsigned acc8[64][8];            // per origin unit: SAD (or SATD) accumulators of the eight candidates of a step
  unsigned acc_c[64];              // kSatd: SATD at the centre of the first step (the full-sample vector)
  unsigned short pen[65 * 65];     // mv penalty of every full-sample candidate (index = (dy+R)*side + dx+R)
  // intra candidates of a P picture (fp.intra_in_p): per 16x16 block (z-order) "try it" / chosen mode
  // (-1 = stays inter); per 8x8 unit: intra mode + 1 of the intra CU covering it (0 = inter)
  uint8_t try_intra[16];
  int8_t intra_mode16[16];
  uint8_t intra_unit[64];
  // two-level search (fp.me_coarse): best key and resulting centre (full samples) of the coarse level
  // per 32x32 quadrant; per origin unit, the set whose window the CU's vector came from
  unsigned ckey[4], czero[4];
  short ctr_x[4], ctr_y[4];
  uint8_t wset[64];
};

// Intra CUs in P pictures: a 16x16 block whose best inter cost exceeds kIntraTryCost gets the 35-mode
// source-based intra search; intra wins when 1.5 x its cost (prediction from reconstructed neighbours
// is worse than from the source ones the search uses) plus kIntraOverheadBits of signalling is smaller.
constexpr unsigned kIntraTryCost = 4096;
constexpr int kIntraOverheadBits = 24;

// quarter-resolution picture for the coarse level of the motion search: every sample the rounded mean
// of a 4x4 block of the luma plane.  One thread per coarse sample.
__global__ void __launch_bounds__(256)
k_down4(const uint8_t *__restrict__ plane, int w, int h, uint8_t *__restrict__ out)
{
  const int wq = w >> 2, hq = h >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= wq * hq) return;
  const int y = i / wq, x = i - y * wq;
  unsigned s = 8;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned v = __ldg((const uint32_t *)(plane + (size_t)(4 * y + j) * w + 4 * x));
    s += (v & 0xff) + ((v >> 8) & 0xff) + ((v >> 16) & 0xff) + (v >> 24);
  }
  out[i] = (uint8_t)(s >> 4);
}

// Search centres of the full-sample level.  Set 0: the zero vector, searched on the CTU-wide window.
// Set 1 (fp.me_coarse > 0): per 32x32 quadrant, the vector the coarse level found on the
// quarter-resolution pictures (+-me_coarse coarse samples = 4 * me_coarse luma samples), searched on
// a window of the quadrant's own; skipped where it is the zero vector again.  Around every centre
// the same +-R window is searched and the mv penalty counts from the centre.
//
// kSatd (fp.subme_satd): the fractional refinement compares the Hadamard SATD of the
// residual, 8x8 tiles with (sum |h| + 2) >> 2 each as oracle/hevc_prims.c orc_satd, instead of the SAD.  The
// four threads of a unit hold two columns x eight rows each: the first horizontal butterfly and the eight-point
// vertical transform are in registers, the other two horizontal stages are shuffles among the four lanes.
template <bool kSatd>
__global__ void __launch_bounds__(kThreads, kSatd ? 3 : 4)     // SAD: <= 64 registers, four CTAs per SM hold the 510 CTUs of a 1080p picture in one wave
k_me_ctu(FrameParams fp, const uint8_t *__restrict__ src, const uint8_t *__restrict__ ref, CuInfo *__restrict__ cu)
{
  extern __shared__ uint32_t s_dyn[];
  __shared__ MeShared sh;
  __shared__ ModeShared shm;
  const int R = fp.search_range;
  const int M = window_margin(R);
  const int WS = kCtb + 2 * M, WSW = (WS >> 2) + 1;
  const int WSq = 32 + 2 * M, WSWq = (WSq >> 2) + 1;          // quadrant windows (set 1)
  uint32_t *s_ref = s_dyn;                     // WS rows x WSW words
  uint32_t *s_src = s_dyn + WS * WSW;          // 64 rows x 16 words
  uint32_t *s_qwin = s_src + 64 * 16;          // 4 x (WSq rows x WSWq words), only with me_coarse
  const int ctb_x = blockIdx.x % fp.ctb_cols, ctb_y = blockIdx.x / fp.ctb_cols;
  const int cx = ctb_x * kCtb, cy = ctb_y * kCtb;
  const int t = threadIdx.x;
  const int lambda_q4 = lambda_q4_at(fp, cx, cy);
  const int side = 2 * R + 1;

  load_window(ref, fp.w, fp.h, cx - M, cy - M, WS, WSW, s_ref);
  for (int i = t; i < 64 * 16; i += kThreads) {
    int y = cy + (i >> 4), x = cx + 4 * (i & 15);
    s_src[i] = (y < fp.h && x < fp.w) ? __ldg((const uint32_t *)(src + (size_t)y * fp.w + x)) : 0u;
  }
  if (t < 64) { sh.key8[t] = 0xffffffffu; sh.acc_c[t] = 0; for (int k = 0; k < 8; k++) sh.acc8[t][k] = 0; }
  for (int c = t; c < side * side; c += kThreads) {
    int dy = c / side - R, dx = c - (dy + R) * side - R;
    sh.pen[c] = (unsigned short)mv_penalty(lambda_q4, dx * 4, dy * 4);
  }
  if (t < 16) sh.key16[t] = 0xffffffffu;
  if (t < 4) { sh.key32[t] = 0xffffffffu; sh.ckey[t] = 0xffffffffu; sh.czero[t] = 0xffffffffu; sh.ctr_x[t] = 0; sh.ctr_y[t] = 0; }
  __syncthreads();

  // ---- coarse level: one displacement per 32x32 quadrant (8x8 coarse samples, cut at the picture edge) ----
  if (fp.me_coarse > 0) {
    const int Rc = fp.me_coarse, wq = fp.w >> 2, hq = fp.h >> 2;
    const int CW = 16 + 2 * Rc, CWW = (CW >> 2) + 1;           // coarse reference window, Rc is a multiple of 4
    uint32_t *s_cref = s_qwin;                                 // the quadrant windows are loaded afterwards
    uint32_t *s_csrc = s_cref + CW * CWW;                      // 16 rows x 4 words
    // (byte loads: the quarter-resolution plane's row pitch is a multiple of 2 only)
    for (int i = t; i < CW * CWW; i += kThreads) {
      const int wy = i / CWW, wi = i - wy * CWW;
      const uint8_t *rowp = fp.ref_q + (size_t)clip3(0, hq - 1, (cy >> 2) - Rc + wy) * wq;
      const int x = (cx >> 2) - Rc + 4 * wi;
      s_cref[i] = (uint32_t)__ldg(rowp + clip3(0, wq - 1, x)) | ((uint32_t)__ldg(rowp + clip3(0, wq - 1, x + 1)) << 8) |
                  ((uint32_t)__ldg(rowp + clip3(0, wq - 1, x + 2)) << 16) | ((uint32_t)__ldg(rowp + clip3(0, wq - 1, x + 3)) << 24);
    }
    if (t < 64) {
      const int y = (cy >> 2) + (t >> 2), x = (cx >> 2) + 4 * (t & 3);
      // widths are multiples of 8, so a coarse row is a whole number of 2-sample pairs: assemble bytes
      uint32_t v = 0;
      for (int j = 0; j < 4; j++)
        if (y < hq && x + j < wq) v |= (uint32_t)__ldg(fp.src_q + (size_t)y * wq + x + j) << (8 * j);
      s_csrc[t] = v;
    }
    __syncthreads();
    const int cside = 2 * Rc + 1;
    // per quadrant: extent inside the picture, as byte masks of its two words and a row count
    unsigned m0[4], m1[4];
    int rows[4];
    for (int q = 0; q < 4; q++) {
      const int bw = min(8, wq - ((cx >> 2) + 8 * (q & 1))), bh = min(8, hq - ((cy >> 2) + 8 * (q >> 1)));
      rows[q] = max(bh, 0);
      m0[q] = bw >= 4 ? 0xffffffffu : (bw > 0 ? (1u << (8 * bw)) - 1 : 0u);
      m1[q] = bw >= 8 ? 0xffffffffu : (bw > 4 ? (1u << (8 * (bw - 4))) - 1 : 0u);
      if (bw <= 0) rows[q] = 0;
    }
    unsigned best[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    for (int c = t; c < cside * cside; c += kThreads) {
      const int dy = c / cside - Rc, dx = c - (dy + Rc) * cside - Rc;
      const unsigned pen = mv_penalty(lambda_q4, dx * 16, dy * 16);
      const int xo = Rc + dx, sft = (xo & 3) * 8;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (rows[q] == 0) continue;
        if (fp.mv_edges && !(mv_allowed(fp, cx + 32 * (q & 1), min(32, fp.w - (cx + 32 * (q & 1))), dx * 16) &&
                             mv_allowed_v(fp, cy + 32 * (q >> 1), min(32, fp.h - (cy + 32 * (q >> 1))), dy * 16))) continue;
        unsigned sad = 0;
        for (int r = 0; r < rows[q]; r++) {
          const uint32_t *rw = s_cref + (Rc + dy + 8 * (q >> 1) + r) * CWW + ((xo + 8 * (q & 1)) >> 2);
          const unsigned a = rw[0], b = rw[1], cc = rw[2];
          const unsigned r0 = __funnelshift_r(a, b, sft), r1 = __funnelshift_r(b, cc, sft);
          const uint32_t *sw = s_csrc + (8 * (q >> 1) + r) * 4 + 2 * (q & 1);
          sad = sad4_acc(sw[0] & m0[q], r0 & m0[q], sad);
          sad = sad4_acc(sw[1] & m1[q], r1 & m1[q], sad);
        }
        best[q] = min(best[q], ((16 * sad + pen) << 13) | (unsigned)c);
        if (dx == 0 && dy == 0) sh.czero[q] = 16 * sad + pen;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const unsigned b = __reduce_min_sync(0xffffffffu, best[q]);
      if ((t & 31) == 0 && b != 0xffffffffu) atomicMin(&sh.ckey[q], b);
    }
    __syncthreads();
    // a second centre only where it clearly beats staying put (below 3/4 of the zero displacement's
    // cost): smooth content matches equally well at many coarse displacements
    if (t < 4 && sh.ckey[t] != 0xffffffffu && (sh.ckey[t] >> 13) < sh.czero[t] - (sh.czero[t] >> 2)) {
      const int c = (int)(sh.ckey[t] & 8191u);
      const int dy = c / cside - Rc, dx = c - (dy + Rc) * cside - Rc;
      // ... and only where the zero-centred window does not cover it anyway (the vector behind a
      // coarse displacement lies within half a coarse step, 2 samples, of it)
      if (max(abs(4 * dx), abs(4 * dy)) + 2 > R) { sh.ctr_x[t] = (short)(4 * dx); sh.ctr_y[t] = (short)(4 * dy); }
    }
    __syncthreads();
    // windows of set 1, centred on each quadrant's coarse vector (a multiple of 4: aligned loads)
    for (int q = 0; q < 4; q++) {
      if (sh.ctr_x[q] == 0 && sh.ctr_y[q] == 0) continue;      // uniform: shared memory
      load_window(ref, fp.w, fp.h, cx + 32 * (q & 1) + sh.ctr_x[q] - M, cy + 32 * (q >> 1) + sh.ctr_y[q] - M, WSq, WSWq,
                  s_qwin + q * WSq * WSWq);
    }
    __syncthreads();
  }

  // ---- full-sample search: lane <-> 8x8 block (z-order), warp pair <-> candidate subset ----
  {
    const int lane = t & 31, wid = t >> 5;
    const int z = ((wid & 1) << 5) | lane;          // z-order index of this thread's 8x8 block
    const int q = wid >> 1;                         // candidates q, q+4, q+8, ...
    const int bx = z_to_x(z), by = z_to_y(z);
    const bool v8 = cx + 8 * bx < fp.w && cy + 8 * by < fp.h;
    const bool v16 = cx + 16 * (bx >> 1) + 16 <= fp.w && cy + 16 * (by >> 1) + 16 <= fp.h;
    const bool v32 = cx + 32 * (bx >> 2) + 32 <= fp.w && cy + 32 * (by >> 2) + 32 <= fp.h;
    uint32_t s[16];
#pragma unroll
    for (int r = 0; r < 8; r++) { s[2 * r] = s_src[(by * 8 + r) * 16 + bx * 2]; s[2 * r + 1] = s_src[(by * 8 + r) * 16 + bx * 2 + 1]; }
    // Candidates are taken four at a time along x, starting on a 4-byte boundary of the window:
    // the three words loaded per row then serve all four (byte shifts 0..3 are compile-time), so a
    // row costs 3 LDS + 6 SHF + 8 VABSDIFF4 for four candidates.  dx runs from -Rr (R rounded up to
    // a multiple of 4); candidates beyond +R are masked out.
    const int Rr = (R + 3) & ~3, groups = (Rr + R) / 4 + 1, items = side * groups;
    unsigned k8 = 0xffffffffu, k16 = 0xffffffffu, k32 = 0xffffffffu;
    const int nsets = fp.me_coarse > 0 ? 2 : 1;
    for (int set = 0; set < nsets; set++) {
      // where this lane's block sits in the window of the set, and the set's centre for it
      const int qd = z >> 4;                               // quadrant of the block (lanes 0-15 / 16-31 differ)
      const int mcx = set ? sh.ctr_x[qd] : 0, mcy = set ? sh.ctr_y[qd] : 0;
      const bool live = set == 0 || mcx != 0 || mcy != 0;
      if (set && __all_sync(0xffffffffu, !live)) continue;  // both quadrants of this warp skip set 1
      const uint32_t *wbase = set ? s_qwin + qd * WSq * WSWq : s_ref;
      const int wp = set ? WSWq : WSW;
      const int lx = M + (set ? (bx & 3) : bx) * 8, ly = M + (set ? (by & 3) : by) * 8;
      const unsigned cbase = (unsigned)(set * side * side);
      for (int it = q; it < items; it += 4) {
        const int dyi = it / groups, g = it - dyi * groups;
        const int dy = dyi - R, dx0 = 4 * g - Rr;
        const uint32_t *row = wbase + (ly + dy) * wp + ((lx + dx0) >> 2);
        unsigned sad0 = 0, sad1 = 0, sad2 = 0, sad3 = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const unsigned a = row[0], b = row[1], cc = row[2];
          const unsigned s0 = s[2 * r], s1 = s[2 * r + 1];
          sad0 = sad4_acc(s0, a, sad0);
          sad0 = sad4_acc(s1, b, sad0);
          sad1 = sad4_acc(s0, __funnelshift_r(a, b, 8), sad1);
          sad1 = sad4_acc(s1, __funnelshift_r(b, cc, 8), sad1);
          sad2 = sad4_acc(s0, __funnelshift_r(a, b, 16), sad2);
          sad2 = sad4_acc(s1, __funnelshift_r(b, cc, 16), sad2);
          sad3 = sad4_acc(s0, __funnelshift_r(a, b, 24), sad3);
          sad3 = sad4_acc(s1, __funnelshift_r(b, cc, 24), sad3);
          row += wp;
        }
        const unsigned sads[4] = {sad0, sad1, sad2, sad3};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int dx = dx0 + j;
          const bool in_range = dx >= -R && dx <= R;          // uniform across the warp
          unsigned sad = v8 ? sads[j] : 0;
          unsigned s16 = sad + __shfl_xor_sync(0xffffffffu, sad, 1);
          s16 += __shfl_xor_sync(0xffffffffu, s16, 2);
          unsigned s32 = s16 + __shfl_xor_sync(0xffffffffu, s16, 4);
          s32 += __shfl_xor_sync(0xffffffffu, s32, 8);
          if (!in_range) continue;
          const unsigned c = (unsigned)((dy + R) * side + dx + R);
          const unsigned pen = sh.pen[c];
          bool a8 = v8 && live, a16 = v16 && live, a32 = v32 && live;
          if (fp.mv_edges) {                                   // tile mode only (uniform branch)
            const int mvx = 4 * (mcx + dx), mvy = 4 * (mcy + dy);
            a8 = a8 && mv_allowed(fp, cx + 8 * bx, 8, mvx) && mv_allowed_v(fp, cy + 8 * by, 8, mvy);
            a16 = a16 && mv_allowed(fp, cx + 16 * (bx >> 1), 16, mvx) && mv_allowed_v(fp, cy + 16 * (by >> 1), 16, mvy);
            a32 = a32 && mv_allowed(fp, cx + 32 * (bx >> 2), 32, mvx) && mv_allowed_v(fp, cy + 32 * (by >> 2), 32, mvy);
          }
          if (a8) k8 = min(k8, ((sad + pen) << 13) | (cbase + c));
          if (a16) k16 = min(k16, ((s16 + pen) << 13) | (cbase + c));
          if (a32) k32 = min(k32, ((s32 + pen) << 13) | (cbase + c));
        }
      }
    }
    atomicMin(&sh.key8[z], k8);
    if ((lane & 3) == 0) atomicMin(&sh.key16[z >> 2], k16);
    if ((lane & 15) == 0) atomicMin(&sh.key32[z >> 4], k32);
  }
  __syncthreads();

  // ---- bottom-up partition decision ----
  const unsigned ovh = (unsigned)((lambda_q4 * kCuOverheadBits) >> 4);
  if (t < 16) {
    unsigned sum8 = 0;
    for (int i = 0; i < 4; i++) {
      unsigned k = sh.key8[4 * t + i];
      if (k != 0xffffffffu) sum8 += (k >> 13) + ovh;
    }
    unsigned k16 = sh.key16[t];
    bool use = k16 != 0xffffffffu && (k16 >> 13) + ovh <= sum8;
    sh.use16[t] = use;
    sh.eff16[t] = use ? (k16 >> 13) + ovh : sum8;
    sh.intra_mode16[t] = -1;
    // intra candidate: blocks that lie wholly inside the picture and predict badly
    sh.try_intra[t] = fp.intra_in_p && k16 != 0xffffffffu && sh.eff16[t] > kIntraTryCost;
  }
  __syncthreads();
  if (fp.intra_in_p) {
    for (int b = 0; b < 16; b++) {
      if (!sh.try_intra[b]) continue;                          // uniform: shared memory
      intra_search_cu(shm, fp, src, cx + 8 * z_to_x(4 * b), cy + 8 * z_to_y(4 * b), 4);
      if (t == 0) {
        const unsigned ic = shm.best_cost + shm.best_cost / 2 + (unsigned)((lambda_q4 * kIntraOverheadBits) >> 4) + ovh;
        if (ic < sh.eff16[b]) { sh.eff16[b] = ic; sh.intra_mode16[b] = (int8_t)shm.best_mode; }
      }
      __syncthreads();
    }
  }
  if (t < 4) {
    unsigned sum16 = sh.eff16[4 * t] + sh.eff16[4 * t + 1] + sh.eff16[4 * t + 2] + sh.eff16[4 * t + 3];
    unsigned k32 = sh.key32[t];
    sh.use32[t] = k32 != 0xffffffffu && (k32 >> 13) + ovh <= sum16;
  }
  __syncthreads();
  if (t < 64) {
    unsigned key;
    int org, l2;
    int imode = 0;
    if (sh.use32[t >> 4]) { key = sh.key32[t >> 4]; org = t & ~15; l2 = 5; }
    else if (sh.intra_mode16[t >> 2] >= 0) { key = 0; org = 0xff; l2 = 4; imode = sh.intra_mode16[t >> 2] + 1; }
    else if (sh.use16[t >> 2]) { key = sh.key16[t >> 2]; org = t & ~3; l2 = 4; }
    else { key = sh.key8[t]; org = t; l2 = 3; }
    if (key == 0xffffffffu) { org = 0xff; l2 = 0; }
    sh.intra_unit[t] = (uint8_t)imode;                       // intra units take no part in the refinement (org = 0xff)
    sh.org[t] = (uint8_t)org;
    sh.log2[t] = (uint8_t)l2;
    if (org == t) {
      int c = (int)(key & 8191u);
      const int set = c >= side * side ? 1 : 0;
      c -= set * side * side;
      int dy = c / side - R, dx = c - (dy + R) * side - R;
      if (set) { dx += sh.ctr_x[t >> 4]; dy += sh.ctr_y[t >> 4]; }
      sh.wset[t] = (uint8_t)set;
      sh.mvx[t] = (short)(dx * 4); sh.mvy[t] = (short)(dy * 4);
      sh.cmx[t] = (short)(dx * 4); sh.cmy[t] = (short)(dy * 4);
      sh.best[t] = key >> 13;                    // SAD at the full-sample mv + its mv penalty
    }
  }
  __syncthreads();

  // ---- half- then quarter-sample refinement: 4 threads per 8x8 unit, 2 columns x 8 rows each.
  // The eight candidates of a step share one centre, so they are evaluated back to back into
  // eight accumulators per CU and compared once per step (same result as comparing one by one:
  // the first strictly smaller cost in candidate order wins).  Samples come from the window of the
  // set that won the CU; the mv penalty keeps counting from that set's centre.
  {
    const int z = t >> 2, qc = t & 3;
    const int ux = z_to_x(z), uy = z_to_y(z);
    const int org = sh.org[z];
    const uint8_t *srcb = (const uint8_t *)s_src;
    unsigned sv[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint8_t *sp = srcb + (uy * 8 + r) * 64 + ux * 8 + 2 * qc;
      sv[r] = (unsigned)sp[0] | ((unsigned)sp[1] << 8);
    }
    const int set = org != 0xff ? sh.wset[org] : 0, qd = z >> 4;
    const uint32_t *wbase = set ? s_qwin + qd * WSq * WSWq : s_ref;
    const int wp = set ? WSWq : WSW;
    // window coordinates of the unit for a zero vector
    const int wx0 = M + (set ? (ux & 3) * 8 - sh.ctr_x[qd] : ux * 8) + 2 * qc, wy0 = M + (set ? (uy & 3) * 8 - sh.ctr_y[qd] : uy * 8);
    const unsigned live = __ballot_sync(0xffffffffu, org != 0xff);      // the four lanes of a unit are in or out together
    for (int step = 2; step >= 1; step--) {
      if (org != 0xff) {
        const int cmx = sh.cmx[org], cmy = sh.cmy[org];
        for (int k = 0; k < ((kSatd && step == 2) ? 9 : 8); k++) {          // kSatd: candidate 8 of the first step = the centre itself
          const int ox = k == 8 ? 0 : (k < 3 ? k - 1 : (k == 3 ? -1 : (k == 4 ? 1 : k - 6)));
          const int oy = k == 8 ? 0 : (k < 3 ? -1 : (k < 5 ? 0 : 1));
          const int mx = cmx + ox * step, my = cmy + oy * step;
          unsigned pr[8];
          mc_luma_2x8(wbase, wp, wx0 + (mx >> 2), wy0 + (my >> 2), mx & 3, my & 3, pr);
          if (!kSatd) {
            unsigned sad = 0;
#pragma unroll
            for (int r = 0; r < 8; r++) sad = sad4_acc(sv[r], pr[r], sad);
            atomicAdd(&sh.acc8[org][k], sad);
          } else {
            int a[8], b[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
              const int d0 = (int)(sv[r] & 0xff) - (int)(pr[r] & 0xff), d1 = (int)(sv[r] >> 8) - (int)(pr[r] >> 8);
              a[r] = d0 + d1; b[r] = d0 - d1;
            }
#pragma unroll
            for (int len = 1; len < 8; len <<= 1)
#pragma unroll
              for (int i = 0; i < 8; i += 2 * len)
#pragma unroll
                for (int j = i; j < i + len; j++) {
                  const int ua = a[j], va = a[j + len], ub = b[j], vb = b[j + len];
                  a[j] = ua + va; a[j + len] = ua - va; b[j] = ub + vb; b[j + len] = ub - vb;
                }
            unsigned sum = 0;
#pragma unroll
            for (int r = 0; r < 8; r++) {
#pragma unroll
              for (int m = 1; m <= 2; m <<= 1) {
                const int oa = __shfl_xor_sync(live, a[r], m), ob = __shfl_xor_sync(live, b[r], m);
                a[r] = (qc & m) ? oa - a[r] : a[r] + oa;
                b[r] = (qc & m) ? ob - b[r] : b[r] + ob;
              }
              sum += (unsigned)(abs(a[r]) + abs(b[r]));
            }
            sum += __shfl_xor_sync(live, sum, 1);
            sum += __shfl_xor_sync(live, sum, 2);
            if (qc == 0) atomicAdd(k == 8 ? &sh.acc_c[org] : &sh.acc8[org][k], (sum + 2) >> 2);
          }
        }
      }
      __syncthreads();
      if (t < 64 && sh.org[t] == t) {
        const int pcx = sh.wset[t] ? 4 * sh.ctr_x[t >> 4] : 0, pcy = sh.wset[t] ? 4 * sh.ctr_y[t >> 4] : 0;
        if (kSatd && step == 2) sh.best[t] = sh.acc_c[t] + mv_penalty(lambda_q4, sh.cmx[t] - pcx, sh.cmy[t] - pcy);
        for (int k = 0; k < 8; k++) {
          const int ox = (k < 3 ? k - 1 : (k == 3 ? -1 : (k == 4 ? 1 : k - 6)));
          const int oy = k < 3 ? -1 : (k < 5 ? 0 : 1);
          const int mx = sh.cmx[t] + ox * step, my = sh.cmy[t] + oy * step;
          unsigned cost = sh.acc8[t][k] + mv_penalty(lambda_q4, mx - pcx, my - pcy);
          const bool ok = !fp.mv_edges || (mv_allowed(fp, cx + 8 * z_to_x(t), 1 << sh.log2[t], mx) &&
                                           mv_allowed_v(fp, cy + 8 * z_to_y(t), 1 << sh.log2[t], my));
          if (ok && cost < sh.best[t]) { sh.best[t] = cost; sh.mvx[t] = (short)mx; sh.mvy[t] = (short)my; }
          sh.acc8[t][k] = 0;
        }
        sh.cmx[t] = sh.mvx[t]; sh.cmy[t] = sh.mvy[t];
      }
      __syncthreads();
    }
  }

  if (fp.me_stats && t == 0) {
    int live = 0, tries = 0, chosen = 0;
    for (int q = 0; q < 4; q++) live += sh.ctr_x[q] != 0 || sh.ctr_y[q] != 0;
    for (int b = 0; b < 16; b++) { tries += sh.try_intra[b]; chosen += sh.intra_mode16[b] >= 0; }
    atomicAdd(&fp.me_stats[0], 1ull);
    if (live) atomicAdd(&fp.me_stats[1], (unsigned long long)live);
    if (tries) atomicAdd(&fp.me_stats[2], (unsigned long long)tries);
    if (chosen) atomicAdd(&fp.me_stats[3], (unsigned long long)chosen);
  }

  // ---- cu map ----
  if (t < 64) {
    int ux = z_to_x(t), uy = z_to_y(t);
    int x8 = (cx >> 3) + ux, y8 = (cy >> 3) + uy;
    int org = sh.org[t];
    if (sh.intra_unit[t] && x8 < fp.w8 && y8 < fp.h8) {
      // reconstructed by the intra pass that follows the inter reconstruction (k_intra_frame)
      CuInfo ci;
      ci.mvx = 0; ci.mvy = 0; ci.log2_size = 4; ci.pred_mode = 1; ci.intra_mode = (uint8_t)(sh.intra_unit[t] - 1); ci.cbf = 0;
      ci.skip = 0; ci.merge_idx = 0xff; ci.mvp_idx = 0; ci.qp = 0;
      ci.ref_idx = 0; ci.chroma_mode = ci.intra_mode; ci.tu_log2 = ci.log2_size < 5 ? ci.log2_size : 5; ci.flags = 0;
      cu[(size_t)y8 * fp.w8 + x8] = ci;
      if (fp.any_intra) *fp.any_intra = 1;
    } else if (org != 0xff && x8 < fp.w8 && y8 < fp.h8) {
      CuInfo ci;
      ci.mvx = sh.mvx[org]; ci.mvy = sh.mvy[org];
      ci.log2_size = sh.log2[t]; ci.pred_mode = 0; ci.intra_mode = 0; ci.cbf = 0;
      ci.skip = 0; ci.merge_idx = 0xff; ci.mvp_idx = 0; ci.qp = 0;
      ci.ref_idx = 0; ci.chroma_mode = ci.intra_mode; ci.tu_log2 = ci.log2_size < 5 ? ci.log2_size : 5; ci.flags = 0;
      cu[(size_t)y8 * fp.w8 + x8] = ci;
    }
  }
}

// ------------------------------------------------------------------------------------------------

struct ReconShared {
  uint8_t org[64], log2[64];
  short mvx[64], mvy[64];
  uint8_t ref[64];                 // per unit: reference index of its CU
  int nz[64];
  uint8_t cbf[64];
  DctWords w;
};

// Reference samples are staged per 8x8 unit (luma: 15 rows x 16 bytes around the unit at its own
// vector; chroma: 7 rows x 8 bytes), straight from the unit's reference picture -- so any vector and
// any reference index work, whatever the search range was (a decoder meets arbitrary vectors).
constexpr int kPatchWordsY = 15 * 4, kPatchWordsC = 7 * 2;

// kDecode = false: encoder (source given, levels and cbf produced).
// kDecode = true : decoder (levels and cbf given by the parser; `src` unused).
template <bool kDecode>
__global__ void __launch_bounds__(kThreads)
k_inter_recon(FrameParams fp, const uint8_t *__restrict__ src, const RefList refs,
              uint8_t *__restrict__ rec, int16_t *__restrict__ levels, CuInfo *__restrict__ cu)
{
  extern __shared__ uint32_t s_dyn[];
  __shared__ ReconShared sh;
  uint32_t *s_ref = s_dyn;                                   // 64 unit patches (luma size; reused for chroma)
  uint8_t *s_src = (uint8_t *)(s_dyn + 64 * kPatchWordsY);   // 64 x 64
  uint8_t *s_pred = s_src + 4096;                            // 64 x 64
  uint8_t *s_rec = s_pred + 4096;                            // 64 x 64
  int16_t *s_a = (int16_t *)(s_rec + 4096);                  // 64 rows, pitch 66
  int16_t *s_b = s_a + 64 * 66;                              // 64 rows, pitch 66
  const int ctb_x = blockIdx.x % fp.ctb_cols, ctb_y = blockIdx.x / fp.ctb_cols;
  const int cx = ctb_x * kCtb, cy = ctb_y * kCtb;
  const int t = threadIdx.x;
  const size_t ysz = (size_t)fp.w * fp.h;

  build_dct_words(sh.w);
  if (t < 64) {
    int ux = z_to_x(t), uy = z_to_y(t);
    int x8 = (cx >> 3) + ux, y8 = (cy >> 3) + uy;
    int org = 0xff, l2 = 0;
    if (x8 < fp.w8 && y8 < fp.h8) {
      CuInfo ci = cu[(size_t)y8 * fp.w8 + x8];
      // the transform blocks are what the residual pipeline works on: the CU itself in the encoder, the
      // transform unit covering the 8x8 unit in the decoder (cbf and size per unit from the parser)
      l2 = kDecode ? ci.tu_log2 : ci.log2_size;
      int n8 = l2 > 3 ? 1 << (l2 - 3) : 1;
      org = xy_to_z(ux & ~(n8 - 1), uy & ~(n8 - 1));
      sh.mvx[t] = ci.mvx; sh.mvy[t] = ci.mvy;
      sh.ref[t] = ci.ref_idx < refs.n ? ci.ref_idx : 0;
      sh.cbf[t] = kDecode ? ci.cbf : 0;
      if (ci.pred_mode != 0) org = 0xff;                   // not an inter CU: left to the intra pass (k_intra_frame)
    } else {
      sh.cbf[t] = 0;
    }
    sh.org[t] = (uint8_t)org; sh.log2[t] = (uint8_t)l2;
  }
  __syncthreads();

  for (int c = 0; c < 3; c++) {
    const int cs = c ? 1 : 0;                                // chroma shift
    const int T = kCtb >> cs, pw = fp.w >> cs, ph = fp.h >> cs;
    const uint8_t *psrc = src + (c == 0 ? 0 : ysz + (c == 2 ? ysz / 4 : 0));
    const size_t poff = c == 0 ? 0 : ysz + (c == 2 ? ysz / 4 : 0);
    uint8_t *prec = rec + poff;
    int16_t *plev = levels + poff;
    const int px0 = cx >> cs, py0 = cy >> cs;
    // reference patches: one (unit, row) task per thread and pass
    {
      const int prow = c ? 7 : 15, pw4 = c ? 2 : 4, us = c ? 4 : 8;
      for (int i = t; i < 64 * prow; i += kThreads) {
        const int z = i / prow, r = i - z * prow;
        if (sh.org[z] == 0xff) continue;
        const int mx = sh.mvx[z], my = sh.mvy[z];
        const int x = px0 + z_to_x(z) * us + (c ? (mx >> 3) - 1 : (mx >> 2) - 3);
        const int y = py0 + z_to_y(z) * us + (c ? (my >> 3) - 1 : (my >> 2) - 3) + r;
        load_patch_row(refs.pic[sh.ref[z]] + poff, pw, ph, x, y, pw4, s_ref + z * (c ? kPatchWordsC : kPatchWordsY) + r * pw4);
      }
    }
    if (!kDecode) {
      for (int i = t; i < T * T / 4; i += kThreads) {
        int y = i / (T / 4), xw = i - y * (T / 4);
        int gy = py0 + y, gx = px0 + 4 * xw;
        ((uint32_t *)s_src)[i] = (gy < ph && gx < pw) ? __ldg((const uint32_t *)(psrc + (size_t)gy * pw + gx)) : 0u;
      }
    }
    // decoder: bit 0 = the block has levels; a unit split into four 4x4 luma blocks has one bit per block
    if (t < 64) sh.nz[t] = kDecode ? ((c == 0 && sh.log2[t] == 2) ? (sh.cbf[t] >> 4) : ((sh.cbf[t] >> c) & 1)) : 0;
    __syncthreads();
    // motion compensation: 4 threads per unit, each from its unit's own patch
    {
      const int z = t >> 2, qc = t & 3;
      const int ux = z_to_x(z), uy = z_to_y(z);
      const int org = sh.org[z];
      if (org != 0xff) {
        int mx = sh.mvx[z], my = sh.mvy[z];
        if (c == 0) {
          unsigned pr[8];
          mc_luma_2x8(s_ref + z * kPatchWordsY, 4, 3 + 2 * qc, 3, mx & 3, my & 3, pr);
#pragma unroll
          for (int r = 0; r < 8; r++) {
            uint8_t *d = s_pred + (uy * 8 + r) * 64 + ux * 8 + 2 * qc;
            d[0] = (uint8_t)(pr[r] & 0xff); d[1] = (uint8_t)(pr[r] >> 8);
          }
        } else {
          unsigned pr[4];
          mc_chroma_1x4(s_ref + z * kPatchWordsC, 2, 1 + qc, 1, mx & 7, my & 7, pr);
#pragma unroll
          for (int r = 0; r < 4; r++) s_pred[(uy * 4 + r) * 32 + ux * 4 + qc] = (uint8_t)pr[r];
        }
      }
    }
    __syncthreads();
    TileGeom g{T, c ? 5 : 6, c ? 2 : 3, cs};
    TqParams q{c ? qp_c_at(fp, cx, cy, c) : qp_at(fp, cx, cy), fp.is_idr, fp.scaling, 3 + c};
    if (!kDecode) {
      forward_tq(g, q, s_src, s_pred, sh.org, sh.log2, sh.w, s_a, s_b, sh.nz);
      // levels (s_b) -> HBM, 4 per thread; s_b then becomes the scratch tile of the inverse path
      for (int i = t; i < T * T / 4; i += kThreads) {
        int y = i / (T / 4), xw = i - y * (T / 4);
        int gy = py0 + y, gx = px0 + 4 * xw;
        if (gy < ph && gx < pw) {
          const uint32_t *lp = (const uint32_t *)(s_b + y * (T + 2) + 4 * xw);      // pitch T+2: 4-byte aligned only
          *(uint2 *)(plev + (size_t)gy * pw + gx) = make_uint2(lp[0], lp[1]);
        }
      }
    } else {
      // dequantise the parsed levels (H.265 8.6.3) straight into the coefficient tile, transposed
      // inside each block (what the inverse vertical pass contracts over)
      const int qper = q.qp / 6, lscale = c_level_scale[q.qp % 6];
      for (int p = t; p < T * T; p += kThreads) {
        int y = p >> g.tlog2, x = p & (T - 1);
        TbPos tb;
        if (tb_at(g, sh.org, sh.log2, x, y, tb) && ((sh.nz[tb.org] >> tb.sub) & 1)) {
          int lvl = plev[(size_t)(py0 + y) * pw + px0 + x];
          s_a[(tb.oy + (x - tb.ox)) * (T + 2) + tb.ox + (y - tb.oy)] =
              dequant_level(lvl, tb.log2n, qper, sl_factor(q.sl, tb.log2n, q.matrix, x - tb.ox, y - tb.oy) * lscale);
        }
      }
    }
    __syncthreads();
    inverse_recon(g, s_pred, sh.org, sh.log2, sh.w, s_a, s_b, sh.nz, s_rec);
    if (!kDecode) {
      for (int i = t; i < T * T / 4; i += kThreads) {
        int y = i / (T / 4), xw = i - y * (T / 4);
        int gy = py0 + y, gx = px0 + 4 * xw;
        if (gy < ph && gx < pw) *(uint32_t *)(prec + (size_t)gy * pw + gx) = ((const uint32_t *)s_rec)[i];
      }
    } else {
      // only inter CUs own their samples here (a decoder may meet intra CUs in later revisions)
      for (int p = t; p < T * T; p += kThreads) {
        int y = p >> g.tlog2, x = p & (T - 1);
        TbPos tb;
        if (tb_at(g, sh.org, sh.log2, x, y, tb)) prec[(size_t)(py0 + y) * pw + px0 + x] = s_rec[p];
      }
    }
    if (!kDecode && t < 64 && sh.org[t] == t && sh.nz[t]) sh.cbf[t] |= (uint8_t)(1 << c);
    __syncthreads();
  }
  if (!kDecode && t < 64) {
    int ux = z_to_x(t), uy = z_to_y(t);
    int x8 = (cx >> 3) + ux, y8 = (cy >> 3) + uy;
    int org = sh.org[t];
    if (org != 0xff) cu[(size_t)y8 * fp.w8 + x8].cbf = sh.cbf[org];
  }
}

// ------------------------------------------------------------------------------------------------
// merge / skip / AMVP decisions from the final motion field (H.265 8.5.3.2.2-8.5.3.2.7, P slice,
// one reference picture, no temporal candidates).  One thread per 8x8 unit; every unit of a CU
// computes the CU's result and writes its own entry.

__device__ __forceinline__ unsigned coding_order(const FrameParams &fp, int x, int y)
{
  return (unsigned)((y >> kCtbLog2) * fp.ctb_cols + (x >> kCtbLog2)) * 64u + (unsigned)xy_to_z((x >> 3) & 7, (y >> 3) & 7);
}

struct Nb { bool ok; int mvx, mvy; };
__device__ __forceinline__ Nb inter_nb(const FrameParams &fp, const CuInfo *cu, unsigned cur_order, int xn, int yn)
{
  Nb n{false, 0, 0};
  if (xn < 0 || yn < 0 || xn >= fp.w || yn >= fp.h) return n;
  if (coding_order(fp, xn, yn) >= cur_order) return n;
  const CuInfo *c = &cu[(size_t)(yn >> 3) * fp.w8 + (xn >> 3)];
  if (c->pred_mode != 0) return n;
  n.ok = true; n.mvx = c->mvx; n.mvy = c->mvy;
  return n;
}
__device__ __forceinline__ bool same_mv(const Nb &a, const Nb &b) { return a.mvx == b.mvx && a.mvy == b.mvy; }

__global__ void __launch_bounds__(kThreads)
k_inter_modes(FrameParams fp, CuInfo *__restrict__ cu)
{
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= fp.w8 * fp.h8) return;
  int x8 = u % fp.w8, y8 = u / fp.w8;
  CuInfo me = cu[u];
  if (me.pred_mode != 0 || me.log2_size < 3) return;
  const int n = 1 << me.log2_size;
  const int x0 = (x8 * 8) & ~(n - 1), y0 = (y8 * 8) & ~(n - 1);
  const unsigned cur = coding_order(fp, x0, y0);
  Nb a1 = inter_nb(fp, cu, cur, x0 - 1, y0 + n - 1);
  Nb b1 = inter_nb(fp, cu, cur, x0 + n - 1, y0 - 1);
  Nb b0 = inter_nb(fp, cu, cur, x0 + n, y0 - 1);
  Nb a0 = inter_nb(fp, cu, cur, x0 - 1, y0 + n);
  Nb b2 = inter_nb(fp, cu, cur, x0 - 1, y0 - 1);
  // merge list with the standard's pairwise pruning (comparisons use availability, not the pruned flags)
  int mvx[kMaxMerge], mvy[kMaxMerge], cnt = 0;
  bool use_b1 = b1.ok && !(a1.ok && same_mv(b1, a1));
  bool use_b0 = b0.ok && !(b1.ok && same_mv(b0, b1));
  bool use_a0 = a0.ok && !(a1.ok && same_mv(a0, a1));
  bool use_b2 = b2.ok && !(a1.ok && same_mv(b2, a1)) && !(b1.ok && same_mv(b2, b1));
  if (a1.ok) { mvx[cnt] = a1.mvx; mvy[cnt++] = a1.mvy; }
  if (use_b1) { mvx[cnt] = b1.mvx; mvy[cnt++] = b1.mvy; }
  if (use_b0) { mvx[cnt] = b0.mvx; mvy[cnt++] = b0.mvy; }
  if (use_a0) { mvx[cnt] = a0.mvx; mvy[cnt++] = a0.mvy; }
  if (use_b2 && cnt < 4) { mvx[cnt] = b2.mvx; mvy[cnt++] = b2.mvy; }
  while (cnt < kMaxMerge) { mvx[cnt] = 0; mvy[cnt++] = 0; }
  int midx = -1;
  for (int i = kMaxMerge - 1; i >= 0; i--)
    if (mvx[i] == me.mvx && mvy[i] == me.mvy) midx = i;
  me.merge_idx = (uint8_t)(midx >= 0 ? midx : 0xff);
  me.skip = (uint8_t)(midx >= 0 && me.cbf == 0);
  me.mvp_idx = 0;
  if (midx < 0) {
    Nb a = a0.ok ? a0 : a1;
    Nb b = b0.ok ? b0 : (b1.ok ? b1 : b2);
    int cx[2], cy[2], k = 0;
    if (a.ok) { cx[k] = a.mvx; cy[k++] = a.mvy; }
    if (b.ok && !(a.ok && same_mv(a, b))) { cx[k] = b.mvx; cy[k++] = b.mvy; }
    while (k < 2) { cx[k] = 0; cy[k++] = 0; }
    int bits0 = mv_comp_bits(me.mvx - cx[0]) + mv_comp_bits(me.mvy - cy[0]);
    int bits1 = mv_comp_bits(me.mvx - cx[1]) + mv_comp_bits(me.mvy - cy[1]);
    me.mvp_idx = (uint8_t)(bits1 < bits0);
    // the predictor itself is re-derived by the entropy coder; stash the mvd for it instead
  }
  cu[u] = me;
}

}  // namespace

static size_t me_smem(int range, int coarse)
{
  int M = (range + 4 + 3) & ~3, WS = kCtb + 2 * M, WSW = (WS >> 2) + 1;
  size_t words = (size_t)WS * WSW + 64 * 16;
  if (coarse > 0) {
    const int WSq = 32 + 2 * M, WSWq = (WSq >> 2) + 1, CW = 16 + 2 * coarse, CWW = (CW >> 2) + 1;
    words += std::max<size_t>((size_t)4 * WSq * WSWq, (size_t)CW * CWW + 64);    // the coarse window lives where the quadrant windows go
  }
  return words * 4;
}
static size_t recon_smem()
{
  return (size_t)64 * kPatchWordsY * 4 + 3 * 4096 + 2 * 64 * 66 * 2;
}

// The opt-in limit of dynamic shared memory is a per-function, process-wide attribute: set it once
// to the largest size any stream may ask for (several encoders / decoders with different search
// ranges launch these kernels concurrently from different host threads).
constexpr int kMaxDynSmem = 200 * 1024;
static void allow_big_smem()
{
  static const bool once = [] {
    cudaFuncSetAttribute(k_me_ctu<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    cudaFuncSetAttribute(k_me_ctu<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    cudaFuncSetAttribute(k_inter_recon<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    cudaFuncSetAttribute(k_inter_recon<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    return true;
  }();
  (void)once;
}

cudaError_t launch_inter_me(const FrameParams &fp, const uint8_t *src, const uint8_t *ref, CuInfo *cu, cudaStream_t s)
{
  if (fp.me_coarse < 0 || (fp.me_coarse & 3) || fp.me_coarse > 32 || (fp.me_coarse > 0 && fp.search_range > 16)) return cudaErrorInvalidValue;
  size_t sm = me_smem(fp.search_range, fp.me_coarse);
  allow_big_smem();
  if (fp.subme_satd) k_me_ctu<true><<<fp.ctb_cols * fp.ctb_rows, kThreads, sm, s>>>(fp, src, ref, cu);
  else k_me_ctu<false><<<fp.ctb_cols * fp.ctb_rows, kThreads, sm, s>>>(fp, src, ref, cu);
  return cudaGetLastError();
}

cudaError_t launch_inter_recon(const FrameParams &fp, const uint8_t *src, const uint8_t *ref, uint8_t *rec,
                               int16_t *levels, CuInfo *cu, cudaStream_t s)
{
  allow_big_smem();
  RefList refs{};
  refs.pic[0] = ref; refs.n = 1;
  k_inter_recon<false><<<fp.ctb_cols * fp.ctb_rows, kThreads, recon_smem(), s>>>(fp, src, refs, rec, levels, cu);
  return cudaGetLastError();
}

// Decoder reconstruction of the inter CUs of a P picture: motion compensation from the reference
// picture each CU names (any vector: reference samples are fetched per 8x8 unit with clamped
// coordinates) + dequantisation + inverse transform.
cudaError_t launch_inter_decode(const FrameParams &fp, const RefList &refs, uint8_t *rec, const int16_t *levels,
                                const CuInfo *cu, cudaStream_t s)
{
  if (refs.n < 1 || refs.n > 16) return cudaErrorInvalidValue;
  allow_big_smem();
  k_inter_recon<true><<<fp.ctb_cols * fp.ctb_rows, kThreads, recon_smem(), s>>>(fp, nullptr, refs, rec, (int16_t *)levels, (CuInfo *)cu);
  return cudaGetLastError();
}

cudaError_t launch_down4(const uint8_t *plane, int w, int h, uint8_t *out, cudaStream_t s)
{
  const int n = (w >> 2) * (h >> 2);
  k_down4<<<(n + 255) / 256, 256, 0, s>>>(plane, w, h, out);
  return cudaGetLastError();
}

cudaError_t launch_inter_modes(const FrameParams &fp, CuInfo *cu, cudaStream_t s)
{
  int units = fp.w8 * fp.h8;
  k_inter_modes<<<(units + kThreads - 1) / kThreads, kThreads, 0, s>>>(fp, cu);
  return cudaGetLastError();
}

}  // namespace b200
