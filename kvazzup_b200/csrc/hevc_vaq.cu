// Variance adaptive quantisation ("vaq", Kvazaar's --vaq <strength>; the reference sets it at
// kvazaarfilter.cpp:280-284 when the user setting is 1..20).  Every picture, each CTU's QP moves by
//     dqp = strength * 0.1 * (ln(max(var_ctu, 4)) - ln(max(var_picture, 4))),
// var = luma variance + the two chroma variances, so that flat areas (where the eye sees blocking)
// are quantised finer and busy ones coarser.  Integer arithmetic throughout (the oracle's
// orc_vaq_offsets is the same, bit for bit): variances as exact fractions
// (n * sum x^2 - (sum x)^2) / n^2, log2 in Q8 by repeated squaring of a 32-bit mantissa,
// dqp = clip(round_half_away(strength * 71 * (L_ctu - L_picture) / 2^18), -12, 12).
//
// Two launches per picture on the stream that consumes the input picture:
//   k_vaq_stats  one CTA per CTU: sum and sum of squares of its three planes (DP4A), 6 words per CTU.
//                Reads the picture once: 1.5 B/px, HBM bound in isolation, but the picture was just
//                written by the colour conversion or the H2D copy and the motion search reads it
//                next, so in the pipeline it is an L2 hit (3 MB at 1080p).
//   k_vaq_qp     one CTA: picture totals, then ctu_qp[i] = clip(ctu_qp[i] - kVaqBias + dqp_i, 0, 51).
#include "hevc_common.h"
#include "hevc_kernels.h"

namespace b200 {
namespace {

typedef unsigned __int128 u128;

// floor(256 * log2(v)), v > 0
__device__ int log2_q8(u128 v)
{
  const unsigned long long hi = (unsigned long long)(v >> 64), lo = (unsigned long long)v;
  const int msb = hi ? 127 - __clzll((long long)hi) : 63 - __clzll((long long)lo);
  unsigned long long m = msb >= 31 ? (unsigned long long)(v >> (msb - 31)) : (unsigned long long)(v << (31 - msb));
  // m in [2^31, 2^32)
  int frac = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    m = (m * m) >> 31;                      // [2^31, 2^33)
    frac <<= 1;
    if (m >> 32) { frac |= 1; m >>= 1; }
  }
  return msb * 256 + frac;
}

// n^2 * (var Y + var U + var V) of a region of n luma and n / 4 samples per chroma plane
__device__ u128 var_numer(unsigned long long n, const unsigned long long *s, const unsigned long long *ss)
{
  const unsigned long long nc = n / 4;
  const u128 y = (u128)n * ss[0] - (u128)s[0] * s[0];
  const u128 u = (u128)nc * ss[1] - (u128)s[1] * s[1];
  const u128 v = (u128)nc * ss[2] - (u128)s[2] * s[2];
  return y + 16 * (u + v);
}

__global__ void __launch_bounds__(256) k_vaq_stats(const uint8_t *__restrict__ src, int w, int h, int ctb_cols,
                                                   uint32_t *__restrict__ stats)
{
  __shared__ uint32_t part[8][6];
  const int ctu = blockIdx.x, cx = ctu % ctb_cols, cy = ctu / ctb_cols, t = threadIdx.x;
  const int x0 = cx * 64, y0 = cy * 64, bw = min(64, w - x0), bh = min(64, h - y0);
  const int cw = w >> 1;
  uint32_t acc[6] = {0, 0, 0, 0, 0, 0};
  // luma: 8 units of 8 samples per row (widths are multiples of 8), two units per thread
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int u = t + 256 * k, ux = (u & 7) * 8, uy = u >> 3;
    if (ux < bw && uy < bh) {
      const uint2 v = *reinterpret_cast<const uint2 *>(src + (size_t)(y0 + uy) * w + x0 + ux);
      acc[0] = __dp4a(v.x, 0x01010101u, acc[0]); acc[0] = __dp4a(v.y, 0x01010101u, acc[0]);
      acc[3] = __dp4a(v.x, v.x, acc[3]); acc[3] = __dp4a(v.y, v.y, acc[3]);
    }
  }
  // chroma: 8 units of 4 samples per row, one unit per thread and plane
  {
    const int ux = (t & 7) * 4, uy = t >> 3;
    if (ux < (bw >> 1) && uy < (bh >> 1)) {
      const uint8_t *pu = src + (size_t)w * h + (size_t)((y0 >> 1) + uy) * cw + (x0 >> 1) + ux;
      const uint32_t a = *reinterpret_cast<const uint32_t *>(pu);
      const uint32_t b = *reinterpret_cast<const uint32_t *>(pu + (size_t)cw * (h >> 1));
      acc[1] = __dp4a(a, 0x01010101u, 0u); acc[4] = __dp4a(a, a, 0u);
      acc[2] = __dp4a(b, 0x01010101u, 0u); acc[5] = __dp4a(b, b, 0u);
    }
  }
#pragma unroll
  for (int k = 0; k < 6; k++) acc[k] = __reduce_add_sync(0xffffffffu, acc[k]);
  if ((t & 31) == 0)
    for (int k = 0; k < 6; k++) part[t >> 5][k] = acc[k];
  __syncthreads();
  if (t < 6) {
    uint32_t a = 0;
    for (int wi = 0; wi < 8; wi++) a += part[wi][t];
    stats[ctu * 6 + t] = a;
  }
}

__global__ void __launch_bounds__(1024) k_vaq_qp(const uint32_t *__restrict__ stats, int w, int h, int ctb_cols, int ctb_rows,
                                                 int strength, uint8_t *__restrict__ ctu_qp)
{
  __shared__ unsigned long long part[32][6];
  __shared__ int s_lf;
  const int t = threadIdx.x, ctus = ctb_cols * ctb_rows;
  unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = t; i < ctus; i += 1024)
    for (int k = 0; k < 6; k++) acc[k] += stats[i * 6 + k];
  for (int k = 0; k < 6; k++)
    for (int d = 16; d; d >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
  if ((t & 31) == 0)
    for (int k = 0; k < 6; k++) part[t >> 5][k] = acc[k];
  __syncthreads();
  if (t == 0) {
    unsigned long long tot[6];
    for (int k = 0; k < 6; k++) { tot[k] = 0; for (int wi = 0; wi < 32; wi++) tot[k] += part[wi][k]; }
    const unsigned long long fn = (unsigned long long)w * h;
    u128 fnum = var_numer(fn, tot, tot + 3);
    if (fnum < (u128)4 * fn * fn) fnum = (u128)4 * fn * fn;      // the floor of the CTUs: a flat picture moves nothing
    s_lf = log2_q8(fnum) - 2 * log2_q8(fn);
  }
  __syncthreads();
  const int lf = s_lf;
  for (int i = t; i < ctus; i += 1024) {
    const int cx = i % ctb_cols, cy = i / ctb_cols;
    const unsigned long long n = (unsigned long long)min(64, w - cx * 64) * min(64, h - cy * 64);
    unsigned long long v[6];
    for (int k = 0; k < 6; k++) v[k] = stats[i * 6 + k];
    u128 num = var_numer(n, v, v + 3);
    if (num < (u128)4 * n * n) num = (u128)4 * n * n;
    const int lc = log2_q8(num) - 2 * log2_q8(n);
    const int tt = strength * 71 * (lc - lf);
    int d = tt >= 0 ? (tt + (1 << 17)) >> 18 : -((-tt + (1 << 17)) >> 18);
    d = min(max(d, -12), 12);
    ctu_qp[i] = (uint8_t)min(max((int)ctu_qp[i] - kVaqBias + d, 0), 51);
  }
}

}  // namespace

cudaError_t launch_vaq(const FrameParams &fp, const uint8_t *src, int strength, uint32_t *stats, uint8_t *ctu_qp, cudaStream_t s)
{
  const int ctus = fp.ctb_cols * fp.ctb_rows;
  k_vaq_stats<<<ctus, 256, 0, s>>>(src, fp.w, fp.h, fp.ctb_cols, stats);
  k_vaq_qp<<<1, 1024, 0, s>>>(stats, fp.w, fp.h, fp.ctb_cols, fp.ctb_rows, strength, ctu_qp);
  return cudaGetLastError();
}

}  // namespace b200
