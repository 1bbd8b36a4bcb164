// Sample adaptive offset (H.265 8.7.3) of the B200 HEVC codec (sm_100a); SURVEY.md 8a-K row K7.
//
// One CTA per CTU, one launch per picture, after deblocking.  SAO reads the DEBLOCKED samples of the
// CTU and of a one-sample ring around it (the neighbours of edge classification) and writes the
// output picture, so no CTU depends on another CTU's result and the whole picture runs in parallel.
//
//   encoder (kDecide): statistics of (source - deblocked) per edge class / category and per band,
//     accumulated in per-warp histograms with one packed shared-memory atomic per sample and bin;
//     best offset per bin and its rate-distortion term by one thread per bin (48 bins per plane);
//     the choice between off / four edge classes / band offset per component group (luma, chroma)
//     exactly as oracle/hevc_enc.c sao_decide_group; the CTU's parameters go to HBM for the
//     binariser; then the offsets are applied.
//   decoder: parameters come from the parser; apply only.
#include "hevc_device.cuh"
#include "hevc_kernels.h"

namespace b200 {

namespace {

constexpr int kSaoThreads = 256, kSaoWarps = kSaoThreads / 32;

// Tiles hold the deblocked samples of a plane of the CTU with a one-sample ring; the interior starts
// at byte column 4 so that every group of four samples is one aligned word (ring: column 3 and
// column T + 4).  Row pitch T + 8.
struct SaoShared {
  uint32_t tile_y[66 * 18];                    // 66 rows x 72 bytes
  uint32_t tile_c[2][34 * 10];                 // 34 rows x 40 bytes
  unsigned hist[kSaoWarps][32];                // per warp: the 32 bands; packed count << 20 | sum (20-bit two's complement)
  int eo_cnt[16], eo_sum[16];                  // edge bins (class * 4 + category index 0..3) of the plane in work
  long long eo_term[3][16];                    // per plane: SSE change + lambda * bits of the bin's best offset
  long long band_term[3][32];
  int8_t eo_off[3][16], band_off[3][32];
  SaoCtu prm;
};

// best offset of a bin given count and sum of (source - deblocked); returns the change in SSE
__device__ __forceinline__ long long best_offset(int count, int sum, int lo, int hi, int &off)
{
  off = 0;
  if (count == 0) return 0;
  int o = (sum >= 0 ? sum + count / 2 : sum - count / 2) / count;
  o = clip3(lo, hi, o);
  long long best = 0;
  for (int k = o; k != 0; k += (k > 0 ? -1 : 1)) {          // smaller magnitudes cost fewer bits: check them all
    const long long d = (long long)count * k * k - 2LL * k * sum;
    if (d < best) { best = d; off = k; }
  }
  return best;
}

struct PlaneGeom { int pw, ph, x0, y0, x1, y1, T, pitchw; };

__device__ __forceinline__ PlaneGeom plane_geom(const FrameParams &fp, int cx, int cy, int c)
{
  PlaneGeom g;
  const int sh = c ? 1 : 0;
  g.pw = fp.w >> sh; g.ph = fp.h >> sh;
  g.x0 = cx >> sh; g.y0 = cy >> sh;
  g.x1 = min(g.pw, (cx + kCtb) >> sh); g.y1 = min(g.ph, (cy + kCtb) >> sh);
  g.T = kCtb >> sh; g.pitchw = (g.T + 8) >> 2;
  return g;
}

// The four samples of word `xw` of tile row `row` (0 = first row of the CTU) and their two neighbours
// of edge class `cls`, as packed bytes.  class 0: a = left, b = right; 1: above, below; 2: above-left,
// below-right; 3: above-right, below-left (8.7.3.2, Table 8-13).
__device__ __forceinline__ void edge_neighbours(const uint32_t *tile, int pitchw, int row, int xw, int cls, uint32_t &c, uint32_t &a, uint32_t &b)
{
  const uint32_t *r1 = tile + (row + 1) * pitchw + xw + 1;      // centre word (interior starts at word 1)
  c = r1[0];
  if (cls == 0) { a = __funnelshift_r(r1[-1], r1[0], 24); b = __funnelshift_r(r1[0], r1[1], 8); return; }
  const uint32_t *r0 = r1 - pitchw, *r2 = r1 + pitchw;
  if (cls == 1) { a = r0[0]; b = r2[0]; }
  else if (cls == 2) { a = __funnelshift_r(r0[-1], r0[0], 24); b = __funnelshift_r(r2[0], r2[1], 8); }
  else { a = __funnelshift_r(r0[0], r0[1], 8); b = __funnelshift_r(r2[-1], r2[0], 24); }
}

// edgeIdx of four samples at once: 2 + sign(c - a) + sign(c - b) per byte (0..4; 2 = no offset)
__device__ __forceinline__ uint32_t edge_index4(uint32_t c, uint32_t a, uint32_t b)
{
  const uint32_t one = 0x01010101u;
  return 0x02020202u + (__vcmpgtu4(c, a) & one) + (__vcmpgtu4(c, b) & one) - (__vcmpltu4(c, a) & one) - (__vcmpltu4(c, b) & one);
}

// Byte mask (0xff per sample) of the samples of a word whose class-`cls` neighbours both lie inside the
// picture; x = plane column of the word's first sample, y = plane row.
__device__ __forceinline__ uint32_t inside_mask4(int x, int y, int pw, int ph, int cls)
{
  if (cls != 0 && (y == 0 || y == ph - 1)) return 0u;
  uint32_t m = 0xffffffffu;
  if (cls != 1) {
    if (x == 0) m &= 0xffffff00u;
    if (x + 4 == pw) m &= 0x00ffffffu;
  }
  return m;
}

// four byte values 0..7 -> a PRMT selector (nibble per byte)
__device__ __forceinline__ uint32_t nibbles_of(uint32_t e)
{
  const uint32_t t = e | (e >> 4);
  return (t & 0xffu) | ((t >> 8) & 0xff00u);
}

template <bool kDecide>
__global__ void __launch_bounds__(kSaoThreads)
k_sao_ctu(FrameParams fp, const uint8_t *__restrict__ src, const uint8_t *__restrict__ dbk, uint8_t *__restrict__ out,
          SaoCtu *params)
{
  __shared__ SaoShared sh;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int ctu = blockIdx.x;
  const int cx = (ctu % fp.ctb_cols) * kCtb, cy = (ctu / fp.ctb_cols) * kCtb;
  const size_t ysz = (size_t)fp.w * fp.h;

  // ---- deblocked samples of the three planes, with their ring, into shared memory (coordinates
  // clamped to the picture: what lies outside is never used, inside_mask4 sees to that) ----
  for (int c = 0; c < 3; c++) {
    const PlaneGeom g = plane_geom(fp, cx, cy, c);
    const uint8_t *pl = dbk + (c == 0 ? 0 : ysz + (c == 2 ? ysz / 4 : 0));
    uint32_t *tile = c == 0 ? sh.tile_y : sh.tile_c[c - 1];
    const int rows = g.T + 2;
    for (int i = t; i < rows * g.pitchw; i += kSaoThreads) {
      const int ty = i / g.pitchw, wi = i - ty * g.pitchw;
      const int y = min(max(g.y0 - 1 + ty, 0), g.ph - 1), x = g.x0 - 4 + 4 * wi;
      const uint8_t *rowp = pl + (size_t)y * g.pw;
      uint32_t v;
      if (x >= 0 && x + 4 <= g.pw) v = __ldg((const uint32_t *)(rowp + x));
      else v = (uint32_t)rowp[min(max(x, 0), g.pw - 1)] | ((uint32_t)rowp[min(max(x + 1, 0), g.pw - 1)] << 8) |
               ((uint32_t)rowp[min(max(x + 2, 0), g.pw - 1)] << 16) | ((uint32_t)rowp[min(max(x + 3, 0), g.pw - 1)] << 24);
      tile[i] = v;
    }
  }
  if (!kDecide && t == 0) sh.prm = params[ctu];
  __syncthreads();

  if (kDecide) {
    const int lq = lambda_q4_at(fp, cx, cy);
    const long long lam = (lq * lq + 128) >> 8;                              // lambda, SSE domain
    for (int c = 0; c < 3; c++) {
      const PlaneGeom g = plane_geom(fp, cx, cy, c);
      const uint8_t *ps = src + (c == 0 ? 0 : ysz + (c == 2 ? ysz / 4 : 0));
      const uint32_t *tile = c == 0 ? sh.tile_y : sh.tile_c[c - 1];
      for (int i = t; i < kSaoWarps * 32; i += kSaoThreads) (&sh.hist[0][0])[i] = 0;
      if (t < 16) { sh.eo_cnt[t] = 0; sh.eo_sum[t] = 0; }
      __syncthreads();
      // Statistics, four samples per step.  Edge bins: per class the edgeIdx of the four samples as
      // packed bytes, then per category a byte mask, the count from its population and the sum of
      // (source - deblocked) as two masked byte sums -- accumulated in registers, reduced at the end.
      // Bands: one packed shared-memory atomic per sample into the warp's own histogram (a warp sees
      // <= 512 samples per plane, so count and |sum| fit 12 + 20 bits).
      const int wd4 = (g.x1 - g.x0) >> 2, ht = g.y1 - g.y0;
      int cnt[16], sum[16];
#pragma unroll
      for (int k = 0; k < 16; k++) { cnt[k] = 0; sum[k] = 0; }
      for (int i = t; i < wd4 * ht; i += kSaoThreads) {
        const int row = i / wd4, xw = i - row * wd4;
        const int x = g.x0 + 4 * xw, y = g.y0 + row;
        const uint32_t sw = __ldg((const uint32_t *)(ps + (size_t)y * g.pw + x));
        uint32_t cw = 0;
#pragma unroll
        for (int cls = 0; cls < 4; cls++) {
          uint32_t a, b;
          edge_neighbours(tile, g.pitchw, row, xw, cls, cw, a, b);
          const uint32_t in = inside_mask4(x, y, g.pw, g.ph, cls);
          const uint32_t e = (edge_index4(cw, a, b) & in) | (0x02020202u & ~in);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint32_t m = __vcmpeq4(e, 0x01010101u * (uint32_t)(k < 2 ? k : k + 1));     // categories 1, 2, 3, 4 = edgeIdx 0, 1, 3, 4
            cnt[cls * 4 + k] += __popc(m) >> 3;
            sum[cls * 4 + k] += (int)__vsadu4(sw & m, 0u) - (int)__vsadu4(cw & m, 0u);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int v = (cw >> (8 * j)) & 0xff, d = (int)((sw >> (8 * j)) & 0xff) - v;
          atomicAdd(&sh.hist[warp][v >> 3], (1u << 20) + (unsigned)d);
        }
      }
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const int cs = __reduce_add_sync(0xffffffffu, cnt[k]), ss = __reduce_add_sync(0xffffffffu, sum[k]);
        if (lane == 0 && cs) { atomicAdd(&sh.eo_cnt[k], cs); atomicAdd(&sh.eo_sum[k], ss); }
      }
      __syncthreads();
      if (t < 48) {
        int o;
        if (t < 16) {
          const int k = (t & 3) + 1;
          const long long d = best_offset(sh.eo_cnt[t], sh.eo_sum[t], k <= 2 ? 0 : -7, k <= 2 ? 7 : 0, o);
          sh.eo_term[c][t] = d + lam * (abs(o) + 1);
          sh.eo_off[c][t] = (int8_t)o;
        } else {
          int bc = 0, bs = 0;
          for (int w = 0; w < kSaoWarps; w++) {
            const unsigned h = sh.hist[w][t - 16];
            const int s2 = ((int)(h << 12)) >> 12;
            bs += s2;
            bc += (int)((h - (unsigned)s2) >> 20);
          }
          const long long d = best_offset(bc, bs, -7, 7, o);
          sh.band_term[c][t - 16] = d + lam * (abs(o) + 1 + (o != 0));
          sh.band_off[c][t - 16] = (int8_t)o;
        }
      }
      __syncthreads();
    }
    // decision per component group: thread 0 luma, thread 32 chroma (two warps, side by side)
    if (t == 0 || t == 32) {
      const int group = t ? 1 : 0, c_first = group, c_last = group ? 2 : 0;
      long long best_cost = 0;                                               // "off": no change, one bin
      int best_type = 0, best_cls = 0;
      for (int cls = 0; cls < 4; cls++) {
        long long cost = lam * 4;
        for (int c = c_first; c <= c_last; c++)
          for (int k = 0; k < 4; k++) cost += sh.eo_term[c][cls * 4 + k];
        if (cost < best_cost) { best_cost = cost; best_type = 2; best_cls = cls; }
      }
      int band[3] = {0, 0, 0};
      {
        long long cost = lam * 2;
        for (int c = c_first; c <= c_last; c++) {
          long long bg = 0;
          int bs = 0;
          for (int st = 0; st <= 28; st++) {
            const long long gsum = sh.band_term[c][st] + sh.band_term[c][st + 1] + sh.band_term[c][st + 2] + sh.band_term[c][st + 3];
            if (st == 0 || gsum < bg) { bg = gsum; bs = st; }
          }
          cost += bg + lam * 5;
          band[c] = bs;
        }
        if (cost < best_cost) { best_cost = cost; best_type = 1; }
      }
      sh.prm.type[group] = (uint8_t)best_type;
      sh.prm.eo_class[group] = (uint8_t)(best_type == 2 ? best_cls : 0);
      for (int c = c_first; c <= c_last; c++) {
        sh.prm.band_pos[c] = (uint8_t)(best_type == 1 ? band[c] : 0);
        for (int k = 0; k < 4; k++)
          sh.prm.offset[c][k] = best_type == 2 ? sh.eo_off[c][best_cls * 4 + k] : (best_type == 1 ? sh.band_off[c][band[c] + k] : 0);
      }
      if (t == 0) sh.prm.pad = 0;
    }
    __syncthreads();
    if (t == 0) params[ctu] = sh.prm;
  }

  // ---- apply: four samples per step.  The offsets of a plane sit in two byte tables (positive and
  // negative parts) that PRMT indexes with the four edgeIdx / band indices at once; saturating byte
  // add and subtract clip to 0..255. ----
  for (int c = 0; c < 3; c++) {
    const PlaneGeom g = plane_geom(fp, cx, cy, c);
    uint8_t *po = out + (c == 0 ? 0 : ysz + (c == 2 ? ysz / 4 : 0));
    const uint32_t *tile = c == 0 ? sh.tile_y : sh.tile_c[c - 1];
    const int grp = c ? 1 : 0;
    const int type = sh.prm.type[grp], cls = sh.prm.eo_class[grp], bpos = sh.prm.band_pos[c];
    // table index: edge offset -> edgeIdx (0, 1: categories 1, 2; 2: none; 3, 4: categories 3, 4); band
    // offset -> band index 0..3, 4 = none
    uint32_t pos_lo = 0, pos_hi = 0, neg_lo = 0, neg_hi = 0;
    for (int k = 0; k < 4; k++) {
      const int o = sh.prm.offset[c][k], idx = type == 2 ? (k < 2 ? k : k + 1) : k;
      const uint32_t p = (uint32_t)max(o, 0), n = (uint32_t)max(-o, 0);
      if (idx < 4) { pos_lo |= p << (8 * idx); neg_lo |= n << (8 * idx); }
      else { pos_hi |= p; neg_hi |= n; }
    }
    const int wd4 = (g.x1 - g.x0) >> 2, ht = g.y1 - g.y0;
    for (int i = t; i < wd4 * ht; i += kSaoThreads) {
      const int row = i / wd4, xw = i - row * wd4;
      const int x = g.x0 + 4 * xw, y = g.y0 + row;
      uint32_t cw, idx4;
      if (type == 2) {
        uint32_t a, b;
        edge_neighbours(tile, g.pitchw, row, xw, cls, cw, a, b);
        const uint32_t in = inside_mask4(x, y, g.pw, g.ph, cls);
        idx4 = (edge_index4(cw, a, b) & in) | (0x02020202u & ~in);
      } else {
        cw = tile[(row + 1) * g.pitchw + xw + 1];
        const uint32_t d = __vsub4((cw >> 3) & 0x1f1f1f1fu, 0x01010101u * (uint32_t)bpos) & 0x1f1f1f1fu;
        const uint32_t lt = __vcmpltu4(d, 0x04040404u);
        idx4 = (d & lt) | (0x04040404u & ~lt);
      }
      uint32_t word = cw;
      if (type) {
        const uint32_t sel = nibbles_of(idx4);
        word = __vsubus4(__vaddus4(cw, __byte_perm(pos_lo, pos_hi, sel)), __byte_perm(neg_lo, neg_hi, sel));
      }
      *(uint32_t *)(po + (size_t)y * g.pw + x) = word;
    }
  }
}

}  // namespace

cudaError_t launch_sao_encode(const FrameParams &fp, const uint8_t *src, const uint8_t *dbk, uint8_t *out, SaoCtu *params,
                              cudaStream_t s)
{
  k_sao_ctu<true><<<fp.ctb_cols * fp.ctb_rows, kSaoThreads, 0, s>>>(fp, src, dbk, out, params);
  return cudaGetLastError();
}

cudaError_t launch_sao_decode(const FrameParams &fp, const uint8_t *dbk, uint8_t *out, const SaoCtu *params, cudaStream_t s)
{
  k_sao_ctu<false><<<fp.ctb_cols * fp.ctb_rows, kSaoThreads, 0, s>>>(fp, nullptr, dbk, out, (SaoCtu *)params);
  return cudaGetLastError();
}

}  // namespace b200
