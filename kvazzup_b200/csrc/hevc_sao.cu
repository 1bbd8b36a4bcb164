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

struct SaoShared {
  uint8_t tile_y[66 * 68];                     // deblocked samples with a one-sample ring, row pitch 68
  uint8_t tile_c[2][34 * 36];                  // pitch 36
  unsigned hist[kSaoWarps][48];                // per warp: 16 edge bins (class * 4 + category - 1) + 32 bands;
                                               // packed count << 20 | sum (20-bit two's complement)
  long long eo_term[3][16];                    // per plane: SSE change + lambda * bits of the bin's best offset
  long long band_term[3][32];
  int8_t eo_off[3][16], band_off[3][32];
  SaoCtu prm;
};

__device__ __forceinline__ int sgn(int v) { return (v > 0) - (v < 0); }

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

struct PlaneGeom { int pw, ph, x0, y0, x1, y1, T, pitch; };

__device__ __forceinline__ PlaneGeom plane_geom(const FrameParams &fp, int cx, int cy, int c)
{
  PlaneGeom g;
  const int sh = c ? 1 : 0;
  g.pw = fp.w >> sh; g.ph = fp.h >> sh;
  g.x0 = cx >> sh; g.y0 = cy >> sh;
  g.x1 = min(g.pw, (cx + kCtb) >> sh); g.y1 = min(g.ph, (cy + kCtb) >> sh);
  g.T = kCtb >> sh; g.pitch = g.T + 4;
  return g;
}

// edge category 1..4 (0 = none) of the sample at tile position (tx, ty) (ring offset included)
// (`inside`: both neighbours of the class lie in the picture -- otherwise the category is 0, 8.7.3.2)
__device__ __forceinline__ int edge_category(const uint8_t *tile, int pitch, int tx, int ty, int cls, bool inside)
{
  if (!inside) return 0;
  // class 0: a = (-1, 0), b = (1, 0); 1: (0,-1),(0,1); 2: (-1,-1),(1,1); 3: (1,-1),(-1,1)
  const int ax = cls == 0 ? -1 : (cls == 1 ? 0 : (cls == 2 ? -1 : 1)), ay = cls == 0 ? 0 : -1;
  const int v = tile[ty * pitch + tx];
  const int e = 2 + sgn(v - tile[(ty + ay) * pitch + tx + ax]) + sgn(v - tile[(ty - ay) * pitch + tx - ax]);
  return e == 2 ? 0 : (e < 2 ? e + 1 : e);
}

// are the two neighbours of class `cls` of plane sample (x, y) inside the picture?
__device__ __forceinline__ bool nb_inside(int x, int y, int pw, int ph, int cls)
{
  const int ax = cls == 0 ? -1 : (cls == 1 ? 0 : (cls == 2 ? -1 : 1)), ay = cls == 0 ? 0 : -1;
  const int x_a = x + ax, y_a = y + ay, x_b = x - ax, y_b = y - ay;
  return x_a >= 0 && x_a < pw && y_a >= 0 && y_a < ph && x_b >= 0 && x_b < pw && y_b >= 0 && y_b < ph;
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

  // ---- deblocked samples of the three planes, with their ring, into shared memory ----
  for (int c = 0; c < 3; c++) {
    const PlaneGeom g = plane_geom(fp, cx, cy, c);
    const uint8_t *pl = dbk + (c == 0 ? 0 : ysz + (c == 2 ? ysz / 4 : 0));
    uint8_t *tile = c == 0 ? sh.tile_y : sh.tile_c[c - 1];
    const int side = g.T + 2;
    for (int i = t; i < side * side; i += kSaoThreads) {
      const int ty = i / side, tx = i - ty * side;
      const int x = min(max(g.x0 - 1 + tx, 0), g.pw - 1), y = min(max(g.y0 - 1 + ty, 0), g.ph - 1);
      tile[ty * g.pitch + tx] = __ldg(pl + (size_t)y * g.pw + x);
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
      const uint8_t *tile = c == 0 ? sh.tile_y : sh.tile_c[c - 1];
      for (int i = t; i < kSaoWarps * 48; i += kSaoThreads) (&sh.hist[0][0])[i] = 0;
      __syncthreads();
      // a warp takes whole rows (32 consecutive samples per pass): <= 512 samples per warp and
      // plane, so count (<= 512) and |sum| (<= 130560) fit the packed 12 + 20 bit accumulators
      const int wd = g.x1 - g.x0, ht = g.y1 - g.y0;
      for (int row = warp; row < ht; row += kSaoWarps) {
        for (int xx = lane; xx < wd; xx += 32) {
          const int x = g.x0 + xx, y = g.y0 + row;
          const int v = tile[(row + 1) * g.pitch + xx + 1];
          const int diff = (int)__ldg(ps + (size_t)y * g.pw + x) - v;
          const unsigned add = (1u << 20) + (unsigned)diff;
          atomicAdd(&sh.hist[warp][16 + (v >> 3)], add);
#pragma unroll
          for (int cls = 0; cls < 4; cls++) {
            const int k = edge_category(tile, g.pitch, xx + 1, row + 1, cls, nb_inside(x, y, g.pw, g.ph, cls));
            if (k) atomicAdd(&sh.hist[warp][cls * 4 + k - 1], add);
          }
        }
      }
      __syncthreads();
      if (t < 48) {
        int cnt = 0, sum = 0;
        for (int w = 0; w < kSaoWarps; w++) {
          const unsigned h = sh.hist[w][t];
          const int s = ((int)(h << 12)) >> 12;
          sum += s;
          cnt += (int)((h - (unsigned)s) >> 20);
        }
        int o;
        if (t < 16) {
          const int k = (t & 3) + 1;
          const long long d = best_offset(cnt, sum, k <= 2 ? 0 : -7, k <= 2 ? 7 : 0, o);
          sh.eo_term[c][t] = d + lam * (abs(o) + 1);
          sh.eo_off[c][t] = (int8_t)o;
        } else {
          const long long d = best_offset(cnt, sum, -7, 7, o);
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

  // ---- apply ----
  for (int c = 0; c < 3; c++) {
    const PlaneGeom g = plane_geom(fp, cx, cy, c);
    uint8_t *po = out + (c == 0 ? 0 : ysz + (c == 2 ? ysz / 4 : 0));
    const uint8_t *tile = c == 0 ? sh.tile_y : sh.tile_c[c - 1];
    const int grp = c ? 1 : 0;
    const int type = sh.prm.type[grp], cls = sh.prm.eo_class[grp], bpos = sh.prm.band_pos[c];
    const int wd4 = (g.x1 - g.x0) >> 2, ht = g.y1 - g.y0;
    for (int i = t; i < wd4 * ht; i += kSaoThreads) {
      const int row = i / wd4, xw = i - row * wd4;
      uint32_t word = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int xx = 4 * xw + j;
        const int v = tile[(row + 1) * g.pitch + xx + 1];
        int o = 0;
        if (type == 2) {
          const int k = edge_category(tile, g.pitch, xx + 1, row + 1, cls, nb_inside(g.x0 + xx, g.y0 + row, g.pw, g.ph, cls));
          if (k) o = sh.prm.offset[c][k - 1];
        } else if (type == 1) {
          const int k = ((v >> 3) - bpos) & 31;
          if (k < 4) o = sh.prm.offset[c][k];
        }
        word |= (uint32_t)clip8(v + o) << (8 * j);
      }
      *(uint32_t *)(po + (size_t)(g.y0 + row) * g.pw + g.x0 + 4 * xw) = word;
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
