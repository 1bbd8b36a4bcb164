// Colour-conversion kernels (sm_100a) and their C-ABI entry points.
//
//   I420 -> RGB32   replaces reference src/media/processing/yuvconversions.cpp:72-420
//   half_rgb        replaces yuvconversions.cpp:852-867
//   flip_rgb        replaces yuvconversions.cpp:869-921
//   * -> I420       replaces libyuv::ConvertToI420 as called at
//                   src/media/processing/libyuvconverter.cpp:120-127
//
// All of these are HBM-bound byte kernels: every datum is read once and
// written once, loads/stores are vectorised so that each warp instruction
// touches whole 128-byte lines, and the grid is a multiple of the SM count
// with a grid-stride loop so the tail wave stays short.
#include "runtime.h"

#include <string.h>

namespace {

using namespace b200;

constexpr int kThreads = 256;

inline int grid_for(size_t items, int per_sm = 8)
{
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  size_t want = (items + kThreads - 1) / kThreads;
  size_t cap = (size_t)sms * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// ---------------------------------------------------------------------------
// I420 -> RGB32 (B,G,R,0).  R = clamp(Y + v + (v>>2) + (v>>3) + (v>>5)),
// G = clamp(Y - ((u>>2)+(u>>4)+(u>>5)) - ((v>>1)+(v>>3)+(v>>4)+(v>>5))),
// B = clamp(Y + u + (u>>1) + (u>>2) + (u>>6)), u=U-128, v=V-128, arithmetic
// shifts (yuvconversions.cpp:133-141).
//
// The three chroma offsets are computed once per chroma sample and broadcast
// into both 16-bit lanes of a word; __viaddmin_s16x2_relu then does
// "add luma, clamp to [0,255]" for two horizontally adjacent pixels in one
// instruction, and two PRMTs per pixel assemble the B,G,R,0 word.

// Item index -> frame / position: 64-bit division costs about as much as the rest of an item, and
// every realistic batch fits 32 bits, so take the short division whenever both operands allow it.
__device__ __forceinline__ size_t div_idx(size_t a, size_t b)
{
  return ((a | b) >> 32) == 0 ? (size_t)((unsigned)a / (unsigned)b) : a / b;
}

struct ChromaOff { unsigned r2, g2, b2; };  // offsets duplicated in both s16 lanes

__device__ __forceinline__ ChromaOff chroma_offsets(int U, int V)
{
  int u = U - 128, v = V - 128;
  int r = v + (v >> 2) + (v >> 3) + (v >> 5);
  int g = -(((u >> 2) + (u >> 4) + (u >> 5)) + ((v >> 1) + (v >> 3) + (v >> 4) + (v >> 5)));
  int b = u + (u >> 1) + (u >> 2) + (u >> 6);
  ChromaOff o;
  o.r2 = __byte_perm((unsigned)r, 0, 0x1010);
  o.g2 = __byte_perm((unsigned)g, 0, 0x1010);
  o.b2 = __byte_perm((unsigned)b, 0, 0x1010);
  return o;
}

// y2 = two luma samples in s16x2 lanes; returns the two B,G,R,0 words.
__device__ __forceinline__ uint2 two_pixels(unsigned y2, const ChromaOff &c)
{
  const unsigned k255 = 0x00FF00FFu;
  unsigned r = __viaddmin_s16x2_relu(y2, c.r2, k255);
  unsigned g = __viaddmin_s16x2_relu(y2, c.g2, k255);
  unsigned b = __viaddmin_s16x2_relu(y2, c.b2, k255);
  unsigned bg = __byte_perm(b, g, 0x6240);        // b0 g0 b1 g1
  uint2 px;
  px.x = __byte_perm(bg, r, 0x5410);              // b0 g0 r0 0(high byte of r0 lane)
  px.y = __byte_perm(bg, r, 0x7632);              // b1 g1 r1 0
  return px;
}

__device__ __forceinline__ uint4 four_pixels(unsigned y4, const ChromaOff &c0, const ChromaOff &c1)
{
  unsigned y01 = __byte_perm(y4, 0, 0x4140);      // y0 0 y1 0
  unsigned y23 = __byte_perm(y4, 0, 0x4342);      // y2 0 y3 0
  uint2 a = two_pixels(y01, c0);
  uint2 b = two_pixels(y23, c1);
  return make_uint4(a.x, a.y, b.x, b.y);
}

// One item = 2 rows x 4 pixels.  Requires w % 4 == 0 and h % 2 == 0.
__device__ __forceinline__ void i420_to_rgb32_item(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                   int w, size_t ysz, int gpr, int f, int rp, int g)
{
  const uint8_t *fy = in + (size_t)f * (ysz + (ysz >> 1));
  const uint8_t *fu = fy + ysz;
  const uint8_t *fv = fu + (ysz >> 2);
  size_t yoff = (size_t)(2 * rp) * w + 4 * g;
  unsigned ya = __ldg((const unsigned *)(fy + yoff));
  unsigned yb = __ldg((const unsigned *)(fy + yoff + w));
  size_t coff = (size_t)rp * (w >> 1) + 2 * g;
  unsigned uu = __ldg((const unsigned short *)(fu + coff));
  unsigned vv = __ldg((const unsigned short *)(fv + coff));
  ChromaOff c0 = chroma_offsets(uu & 0xFF, vv & 0xFF);
  ChromaOff c1 = chroma_offsets(uu >> 8, vv >> 8);
  uint4 *o = (uint4 *)(out + ((size_t)f * ysz + yoff) * 4);
  __stcs(o, four_pixels(ya, c0, c1));
  __stcs(o + gpr, four_pixels(yb, c0, c1));    // next row: + w*4 bytes = gpr uint4
}

__global__ void __launch_bounds__(kThreads)
k_i420_to_rgb32_v4(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                   int w, int h, int n_frames)
{
  const int gpr = w >> 2;                       // 4-pixel groups per row
  const size_t items_per_frame = (size_t)gpr * (h >> 1);
  const size_t total = items_per_frame * n_frames;
  const size_t ysz = (size_t)w * h;
  if (total < 0x80000000ull) {
    // 32-bit index arithmetic: the two divisions below are a fifth of the work of an item when
    // done in 64 bits
    const unsigned ipf = (unsigned)items_per_frame, tot = (unsigned)total, stride = gridDim.x * blockDim.x;
    for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < tot; it += stride) {
      const unsigned f = it / ipf, r = it - f * ipf;
      const unsigned rp = r / (unsigned)gpr;
      i420_to_rgb32_item(in, out, w, ysz, gpr, (int)f, (int)rp, (int)(r - rp * gpr));
    }
    return;
  }
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)(it / items_per_frame);
    int r = (int)(it - (size_t)f * items_per_frame);
    int rp = r / gpr;                           // row pair
    i420_to_rgb32_item(in, out, w, ysz, gpr, f, rp, r - rp * gpr);
  }
}

// Any even w,h: one thread per 2x2 block, scalar accesses.
__global__ void __launch_bounds__(kThreads)
k_i420_to_rgb32_generic(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                        int w, int h, int n_frames)
{
  const int cw = w >> 1, ch = h >> 1;
  const size_t per = (size_t)cw * ch, total = per * n_frames;
  const size_t ysz = (size_t)w * h, fin = ysz + (ysz >> 1);
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    int r = (int)(it - (size_t)f * per);
    int cy = r / cw, cx = r - cy * cw;
    const uint8_t *fy = in + (size_t)f * fin;
    ChromaOff c = chroma_offsets(fy[ysz + (size_t)cy * cw + cx], fy[ysz + (ysz >> 2) + (size_t)cy * cw + cx]);
    for (int dy = 0; dy < 2; dy++) {
      size_t o = (size_t)(2 * cy + dy) * w + 2 * cx;
      unsigned y2 = (unsigned)fy[o] | ((unsigned)fy[o + 1] << 16);
      uint2 px = two_pixels(y2, c);
      unsigned *dst = (unsigned *)(out + ((size_t)f * ysz + o) * 4);
      dst[0] = px.x;
      dst[1] = px.y;
    }
  }
}

// ---------------------------------------------------------------------------
// half_rgb: keep pixel (2x,2y).  One thread produces 2 output pixels from one
// 16-byte load (4 input pixels).  Requires w % 4 == 0.
__global__ void __launch_bounds__(kThreads)
k_half_rgb(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int w, int h, int n_frames)
{
  const int ow = w >> 1, oh = (h + 1) >> 1;
  const int gpr = ow >> 1;
  const size_t per = (size_t)gpr * oh, total = per * n_frames;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    int r = (int)(it - (size_t)f * per);
    int y = r / gpr, g = r - y * gpr;
    uint4 v = __ldcs((const uint4 *)(in + (size_t)f * w * h + (size_t)(2 * y) * w + 4 * g));
    *(uint2 *)(out + (size_t)f * ow * oh + (size_t)y * ow + 2 * g) = make_uint2(v.x, v.z);
  }
}

__global__ void __launch_bounds__(kThreads)
k_half_rgb_generic(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int w, int h, int n_frames)
{
  const int ow = w >> 1, oh = (h + 1) >> 1;
  const size_t per = (size_t)ow * oh, total = per * n_frames;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    int r = (int)(it - (size_t)f * per);
    int y = r / ow, x = r - y * ow;
    out[(size_t)f * per + r] = in[(size_t)f * w * h + (size_t)(2 * y) * w + 2 * x];
  }
}

// flip_rgb: mirror.  One thread moves 4 pixels (16 bytes) when w % 4 == 0.
__global__ void __launch_bounds__(kThreads)
k_flip_rgb_v4(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int w, int h,
              int hor, int ver, int n_frames)
{
  const int gpr = w >> 2;
  const size_t per = (size_t)gpr * h, total = per * n_frames;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    int r = (int)(it - (size_t)f * per);
    int y = r / gpr, g = r - y * gpr;
    int sy = ver ? h - 1 - y : y;
    int sg = hor ? gpr - 1 - g : g;
    uint4 v = __ldcs((const uint4 *)(in + (size_t)f * w * h + (size_t)sy * w + 4 * sg));
    if (hor) v = make_uint4(v.w, v.z, v.y, v.x);
    __stcs((uint4 *)(out + (size_t)f * w * h + (size_t)y * w + 4 * g), v);
  }
}

__global__ void __launch_bounds__(kThreads)
k_flip_rgb_generic(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int w, int h,
                   int hor, int ver, int n_frames)
{
  const size_t per = (size_t)w * h, total = per * n_frames;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    int r = (int)(it - (size_t)f * per);
    int y = r / w, x = r - y * w;
    int sy = ver ? h - 1 - y : y, sx = hor ? w - 1 - x : x;
    out[(size_t)f * per + r] = in[(size_t)f * per + (size_t)sy * w + sx];
  }
}

// ---------------------------------------------------------------------------
// Camera formats -> I420.

// dot product of four unsigned bytes (a) with four signed bytes (b), accumulated
__device__ __forceinline__ int dp4a_u8s8(unsigned a, unsigned b, int c)
{
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ unsigned rgb_y(int r, int g, int b) { return (unsigned)(66 * r + 129 * g + 25 * b + 0x1080) >> 8; }
__device__ __forceinline__ unsigned rgb_u(int r, int g, int b) { return ((unsigned)(112 * b - 74 * g - 38 * r + 0x8080) >> 8) & 0xFF; }
__device__ __forceinline__ unsigned rgb_v(int r, int g, int b) { return ((unsigned)(112 * r - 94 * g - 18 * b + 0x8080) >> 8) & 0xFF; }

// Packed 4:2:2, one item = 2 rows x 8 pixels: two 16-byte loads, two 8-byte
// luma stores, one 4-byte store each for U and V.  Requires w % 8 == 0, h even.
// kYFirst: YUY2 (Y0 U Y1 V) else UYVY (U Y0 V Y1).
template <bool kYFirst>
__global__ void __launch_bounds__(kThreads)
k_packed422_to_i420_v8(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h, int n_frames)
{
  const int gpr = w >> 3;
  const size_t per = (size_t)gpr * (h >> 1), total = per * n_frames;
  const size_t ysz = (size_t)w * h, fout = ysz + (ysz >> 1), fin = ysz * 2;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    int r = (int)(it - (size_t)f * per);
    int rp = r / gpr, g = r - rp * gpr;
    const uint8_t *src = in + (size_t)f * fin + (size_t)(2 * rp) * w * 2 + (size_t)g * 16;
    uint4 a = __ldcs((const uint4 *)src);
    uint4 b = __ldcs((const uint4 *)(src + (size_t)w * 2));
    // luma: bytes 0,2 of each word (YUY2) or 1,3 (UYVY)
    const unsigned ysel = kYFirst ? 0x6420 : 0x7531;
    const unsigned csel = kYFirst ? 0x7531 : 0x6420;   // -> u0 u1 v0 v1 after a second shuffle
    uint2 ya = make_uint2(__byte_perm(a.x, a.y, ysel), __byte_perm(a.z, a.w, ysel));
    uint2 yb = make_uint2(__byte_perm(b.x, b.y, ysel), __byte_perm(b.z, b.w, ysel));
    // chroma bytes of words (x,y): u0 v0 u1 v1 ; rounded average of the two rows
    unsigned ca0 = __byte_perm(a.x, a.y, csel), ca1 = __byte_perm(a.z, a.w, csel);
    unsigned cb0 = __byte_perm(b.x, b.y, csel), cb1 = __byte_perm(b.z, b.w, csel);
    unsigned c0 = __vavgu4(ca0, cb0), c1 = __vavgu4(ca1, cb1);   // (a+b+1)>>1 per byte
    unsigned u4 = __byte_perm(c0, c1, 0x6420);
    unsigned v4 = __byte_perm(c0, c1, 0x7531);
    uint8_t *fy = out + (size_t)f * fout;
    size_t yoff = (size_t)(2 * rp) * w + 8 * g;
    *(uint2 *)(fy + yoff) = ya;
    *(uint2 *)(fy + yoff + w) = yb;
    size_t coff = (size_t)rp * (w >> 1) + 4 * g;
    *(unsigned *)(fy + ysz + coff) = u4;
    *(unsigned *)(fy + ysz + (ysz >> 2) + coff) = v4;
  }
}

// Planar / semi-planar sources, one item = 16 luma bytes copied, plus for the
// first quarter of the items 8 chroma samples produced.  Requires w % 16 == 0.
// mode 0: NV12, 1: NV21, 2: I422.
template <int kMode>
__global__ void __launch_bounds__(kThreads)
k_planar_to_i420_v16(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h, int n_frames)
{
  const size_t ysz = (size_t)w * h;
  const size_t fin = kMode == 2 ? ysz * 2 : ysz + (ysz >> 1);
  const size_t fout = ysz + (ysz >> 1);
  const size_t yitems = ysz >> 4;
  const int cgpr = w >> 4;                                 // 8-sample chroma groups per chroma row
  const size_t citems = (size_t)cgpr * (h >> 1);
  const size_t per = yitems + citems, total = per * n_frames;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    size_t r = it - (size_t)f * per;
    const uint8_t *src = in + (size_t)f * fin;
    uint8_t *dst = out + (size_t)f * fout;
    if (r < yitems) {
      __stcs((uint4 *)dst + r, __ldcs((const uint4 *)src + r));
      continue;
    }
    r -= yitems;
    int cy = (int)div_idx(r, (size_t)cgpr), g = (int)(r - (size_t)cy * cgpr);
    size_t coff = (size_t)cy * (w >> 1) + 8 * g;
    uint2 u8, v8;
    if (kMode == 2) {
      const uint8_t *pu = src + ysz + (size_t)(2 * cy) * (w >> 1) + 8 * g;
      const uint8_t *pv = pu + ysz / 2;
      uint2 a = __ldcs((const uint2 *)pu), b = __ldcs((const uint2 *)(pu + (w >> 1)));
      uint2 c = __ldcs((const uint2 *)pv), d = __ldcs((const uint2 *)(pv + (w >> 1)));
      u8 = make_uint2(__vavgu4(a.x, b.x), __vavgu4(a.y, b.y));
      v8 = make_uint2(__vavgu4(c.x, d.x), __vavgu4(c.y, d.y));
    } else {
      uint4 uv = __ldcs((const uint4 *)(src + ysz + (size_t)cy * w + 16 * g));
      uint2 e = make_uint2(__byte_perm(uv.x, uv.y, 0x6420), __byte_perm(uv.z, uv.w, 0x6420));
      uint2 o = make_uint2(__byte_perm(uv.x, uv.y, 0x7531), __byte_perm(uv.z, uv.w, 0x7531));
      u8 = kMode == 0 ? e : o;
      v8 = kMode == 0 ? o : e;
    }
    *(uint2 *)(dst + ysz + coff) = u8;
    *(uint2 *)(dst + ysz + (ysz >> 2) + coff) = v8;
  }
}

// Four pixels of two rows, one word per pixel (bytes in memory order; the byte that is not R, G or B meets
// a zero coefficient) -> four luma samples per row and two chroma pairs.
__device__ __forceinline__ void rgb_quads_to_i420(const unsigned (&wa)[4], const unsigned (&wb)[4], int ro, int go, int bo,
                                                  uint8_t *fy, size_t ysz, int w, int rp, int g)
{
    // One DP4A per output sample: the pixel word (bytes in memory order) times a coefficient word
    // with 66 / 129 / 25 (unsigned) or the signed chroma weights at the R, G, B byte positions; the
    // unused byte of a pixel (alpha, or the next pixel's first byte for 24-bit input) meets a 0.
    const int rs = ro * 8, gs = go * 8, bs = bo * 8;
    const unsigned cy = (66u << rs) | (129u << gs) | (25u << bs);
    const unsigned cu = ((unsigned)(uint8_t)(-38) << rs) | ((unsigned)(uint8_t)(-74) << gs) | (112u << bs);
    const unsigned cv = (112u << rs) | ((unsigned)(uint8_t)(-94) << gs) | ((unsigned)(uint8_t)(-18) << bs);
    // the sample is byte 1 of each sum (>> 8): PRMT gathers four of them into a word (the sums stay
    // below 2^16, so nothing else has to be masked)
    const unsigned ya = __byte_perm(__byte_perm(__dp4a(wa[0], cy, 0x1080u), __dp4a(wa[1], cy, 0x1080u), 0x0051),
                                    __byte_perm(__dp4a(wa[2], cy, 0x1080u), __dp4a(wa[3], cy, 0x1080u), 0x0051), 0x5410);
    const unsigned yb = __byte_perm(__byte_perm(__dp4a(wb[0], cy, 0x1080u), __dp4a(wb[1], cy, 0x1080u), 0x0051),
                                    __byte_perm(__dp4a(wb[2], cy, 0x1080u), __dp4a(wb[3], cy, 0x1080u), 0x0051), 0x5410);
    unsigned us[2], vs[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      // rounded 2x2 box average of every byte lane, two 16-bit lanes at a time
      const unsigned k = 0x00FF00FFu;
      const unsigned p00 = wa[2 * i], p01 = wa[2 * i + 1], p10 = wb[2 * i], p11 = wb[2 * i + 1];
      const unsigned ev = (p00 & k) + (p01 & k) + (p10 & k) + (p11 & k) + 0x00020002u;
      const unsigned od = __byte_perm(p00, 0, 0x4341) + __byte_perm(p01, 0, 0x4341) + __byte_perm(p10, 0, 0x4341) +
                          __byte_perm(p11, 0, 0x4341) + 0x00020002u;
      // (sum >> 2) of the four 16-bit lanes back into byte order: even lanes to bytes 0, 2, odd lanes to 1, 3
      const unsigned avg = __byte_perm((ev >> 2) & k, (od >> 2) & k, 0x6240);
      us[i] = (unsigned)dp4a_u8s8(avg, cu, 0x8080);
      vs[i] = (unsigned)dp4a_u8s8(avg, cv, 0x8080);
    }
    const unsigned u2 = __byte_perm(us[0], us[1], 0x0051), v2 = __byte_perm(vs[0], vs[1], 0x0051);
    size_t yoff = (size_t)(2 * rp) * w + 4 * g;
    *(unsigned *)(fy + yoff) = ya;
    *(unsigned *)(fy + yoff + w) = yb;
    size_t coff = (size_t)rp * (w >> 1) + 2 * g;
    *(unsigned short *)(fy + ysz + coff) = (unsigned short)u2;
    *(unsigned short *)(fy + ysz + (ysz >> 2) + coff) = (unsigned short)v2;
}

// Packed RGB (3 or 4 bytes per pixel), one item = 2 rows x 4 pixels.
// ro/go/bo are byte offsets of R,G,B inside a pixel.  Requires w % 4 == 0.
template <int kBpp>
__global__ void __launch_bounds__(kThreads)
k_rgb_to_i420_v4(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h, int n_frames,
                 int ro, int go, int bo)
{
  const int gpr = w >> 2;
  const size_t per = (size_t)gpr * (h >> 1), total = per * n_frames;
  const size_t ysz = (size_t)w * h, fout = ysz + (ysz >> 1), fin = ysz * kBpp;
  const bool small = total < 0x80000000ull;      // 32-bit index arithmetic when it fits (the divisions are a good part of an item)
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f, rp, g;
    if (small) {
      const unsigned i32 = (unsigned)it, uf = i32 / (unsigned)per, r = i32 - uf * (unsigned)per, urp = r / (unsigned)gpr;
      f = (int)uf; rp = (int)urp; g = (int)(r - urp * (unsigned)gpr);
    } else {
      f = (int)div_idx(it, per);
      const int r = (int)(it - (size_t)f * per);
      rp = r / gpr; g = r - rp * gpr;
    }
    const uint8_t *src = in + (size_t)f * fin + ((size_t)(2 * rp) * w + 4 * g) * kBpp;
    unsigned wa[4], wb[4];
    if (kBpp == 4) {
      uint4 a = __ldcs((const uint4 *)src), b = __ldcs((const uint4 *)(src + (size_t)w * 4));
      wa[0] = a.x; wa[1] = a.y; wa[2] = a.z; wa[3] = a.w;
      wb[0] = b.x; wb[1] = b.y; wb[2] = b.z; wb[3] = b.w;
    } else {
      // 12 bytes = 4 pixels; split into one word per pixel (top byte unused)
      const unsigned *pa = (const unsigned *)src, *pb = (const unsigned *)(src + (size_t)w * 3);
      unsigned a0 = __ldcs(pa), a1 = __ldcs(pa + 1), a2 = __ldcs(pa + 2);
      unsigned b0 = __ldcs(pb), b1 = __ldcs(pb + 1), b2 = __ldcs(pb + 2);
      wa[0] = a0; wa[1] = __byte_perm(a0, a1, 0x0543); wa[2] = __byte_perm(a1, a2, 0x0432); wa[3] = a2 >> 8;
      wb[0] = b0; wb[1] = __byte_perm(b0, b1, 0x0543); wb[2] = __byte_perm(b1, b2, 0x0432); wb[3] = b2 >> 8;
    }
    rgb_quads_to_i420(wa, wb, ro, go, bo, out + (size_t)f * fout, ysz, w, rp, g);
  }
}

// 24-bit RGB with w % 16 == 0: a thread's four pixels are 12 bytes, so direct loads are 4-byte words at a
// 12-byte stride (every warp-wide load touches three times the sectors it uses).  Here every warp stages
// two rows x 128 pixels (2 x 384 bytes) through its own slice of shared memory with coalesced 16-byte
// loads -- no CTA-wide barrier, warps run independently -- and the 12-byte reads that follow are
// conflict-free (word stride 3 is coprime with the 32 banks).
constexpr int kRgb24TilePx = 128;
__global__ void __launch_bounds__(kThreads)
k_rgb24_to_i420_staged(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h, int n_frames,
                       int ro, int go, int bo)
{
  __shared__ uint4 s_rows[kThreads / 32][2][kRgb24TilePx * 3 / 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tpr = (w + kRgb24TilePx - 1) / kRgb24TilePx;            // tiles per row pair
  const unsigned per = (unsigned)tpr * (unsigned)(h >> 1), total = per * (unsigned)n_frames;
  const size_t ysz = (size_t)w * h, fout = ysz + (ysz >> 1), fin = ysz * 3;
  const unsigned wstride = gridDim.x * (kThreads / 32);
  uint4 (*sr)[kRgb24TilePx * 3 / 16] = s_rows[warp];
  for (unsigned it = blockIdx.x * (kThreads / 32) + warp; it < total; it += wstride) {
    const unsigned f = it / per, r = it - f * per;
    const int rp = (int)(r / (unsigned)tpr), tile = (int)(r - (unsigned)rp * tpr);
    const int x0 = tile * kRgb24TilePx, tw = min(kRgb24TilePx, w - x0), nvec = tw * 3 / 16;     // nvec <= 24
    const uint8_t *src = in + (size_t)f * fin + ((size_t)(2 * rp) * w + x0) * 3;
    __syncwarp();                                         // the previous tile has been read
    if (lane < nvec) {
      sr[0][lane] = __ldg((const uint4 *)src + lane);
      sr[1][lane] = __ldg((const uint4 *)(src + (size_t)w * 3) + lane);
    }
    __syncwarp();
    if (4 * lane < tw) {
      const unsigned *pa = (const unsigned *)sr[0] + 3 * lane, *pb = (const unsigned *)sr[1] + 3 * lane;
      const unsigned a0 = pa[0], a1 = pa[1], a2 = pa[2], b0 = pb[0], b1 = pb[1], b2 = pb[2];
      unsigned wa[4], wb[4];
      wa[0] = a0; wa[1] = __byte_perm(a0, a1, 0x0543); wa[2] = __byte_perm(a1, a2, 0x0432); wa[3] = a2 >> 8;
      wb[0] = b0; wb[1] = __byte_perm(b0, b1, 0x0543); wb[2] = __byte_perm(b1, b2, 0x0432); wb[3] = b2 >> 8;
      rgb_quads_to_i420(wa, wb, ro, go, bo, out + (size_t)f * fout, ysz, w, rp, (x0 >> 2) + lane);
    }
  }
}

// Any even w,h, any supported format: one thread per 2x2 block, byte accesses.
// fmt: 0 YUY2, 1 UYVY, 2 NV12, 3 NV21, 4 I422, 5 RGB (bpp, ro, go, bo), 6 I420 copy.
__global__ void __launch_bounds__(kThreads)
k_to_i420_generic(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h, int n_frames,
                  int fmt, int bpp, int ro, int go, int bo, size_t fin)
{
  const int cw = w >> 1, ch = h >> 1;
  const size_t per = (size_t)cw * ch, total = per * n_frames;
  const size_t ysz = (size_t)w * h, fout = ysz + (ysz >> 1);
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    int r = (int)(it - (size_t)f * per);
    int cy = r / cw, cx = r - cy * cw;
    const uint8_t *src = in + (size_t)f * fin;
    uint8_t *dst = out + (size_t)f * fout;
    unsigned yy[4], U, V;
    if (fmt <= 1) {
      const uint8_t *a = src + ((size_t)(2 * cy) * cw + cx) * 4, *b = a + (size_t)cw * 4;
      int yo = fmt == 0 ? 0 : 1, uo = fmt == 0 ? 1 : 0, vo = fmt == 0 ? 3 : 2;
      yy[0] = a[yo]; yy[1] = a[yo + 2]; yy[2] = b[yo]; yy[3] = b[yo + 2];
      U = (a[uo] + b[uo] + 1) >> 1;
      V = (a[vo] + b[vo] + 1) >> 1;
    } else if (fmt <= 4 || fmt == 6) {
      const uint8_t *a = src + (size_t)(2 * cy) * w + 2 * cx;
      yy[0] = a[0]; yy[1] = a[1]; yy[2] = a[w]; yy[3] = a[w + 1];
      if (fmt == 4) {
        const uint8_t *pu = src + ysz + (size_t)(2 * cy) * cw + cx, *pv = pu + ysz / 2;
        U = (pu[0] + pu[cw] + 1) >> 1;
        V = (pv[0] + pv[cw] + 1) >> 1;
      } else if (fmt == 6) {
        U = src[ysz + (size_t)cy * cw + cx];
        V = src[ysz + (ysz >> 2) + (size_t)cy * cw + cx];
      } else {
        const uint8_t *puv = src + ysz + ((size_t)cy * cw + cx) * 2;
        U = fmt == 2 ? puv[0] : puv[1];
        V = fmt == 2 ? puv[1] : puv[0];
      }
    } else {
      const uint8_t *a = src + ((size_t)(2 * cy) * w + 2 * cx) * bpp, *b = a + (size_t)w * bpp;
      const uint8_t *p[4] = {a, a + bpp, b, b + bpp};
      int sr = 2, sg = 2, sb = 2;
      for (int i = 0; i < 4; i++) {
        yy[i] = rgb_y(p[i][ro], p[i][go], p[i][bo]);
        sr += p[i][ro]; sg += p[i][go]; sb += p[i][bo];
      }
      U = rgb_u(sr >> 2, sg >> 2, sb >> 2);
      V = rgb_v(sr >> 2, sg >> 2, sb >> 2);
    }
    size_t yoff = (size_t)(2 * cy) * w + 2 * cx;
    dst[yoff] = (uint8_t)yy[0]; dst[yoff + 1] = (uint8_t)yy[1];
    dst[yoff + w] = (uint8_t)yy[2]; dst[yoff + w + 1] = (uint8_t)yy[3];
    dst[ysz + (size_t)cy * cw + cx] = (uint8_t)U;
    dst[ysz + (ysz >> 2) + (size_t)cy * cw + cx] = (uint8_t)V;
  }
}

// ---------------------------------------------------------------------------
// Self-view path fused: I420 -> RGB32, optional 2x point decimation
// (HalfRGBFilter, halfrgbfilter.cpp:21-43) and optional mirroring (Filter::normalizeOrientation,
// filter.cpp:263-294) in one pass -- the RGB32 picture is written once instead of three times.
// Output pixel (ox,oy) takes source pixel (sx,sy) = (half ? 2 : 1) * (flip ? last - o : o).
// One thread produces 4 output pixels (one 16-byte store).  Requires ow % 4 == 0.
__global__ void __launch_bounds__(kThreads)
k_selfview(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h, int half, int hor, int ver, int n_frames)
{
  const int ow = half ? w >> 1 : w, oh = half ? (h + 1) >> 1 : h;
  const int gpr = ow >> 2;
  const size_t per = (size_t)gpr * oh, total = per * n_frames;
  const size_t ysz = (size_t)w * h, fin = ysz + (ysz >> 1);
  const int step = half ? 2 : 1;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (size_t)gridDim.x * blockDim.x) {
    int f = (int)div_idx(it, per);
    int r = (int)(it - (size_t)f * per);
    int oy = r / gpr, g = r - oy * gpr;
    const uint8_t *fy = in + (size_t)f * fin;
    const uint8_t *fu = fy + ysz, *fv = fu + (ysz >> 2);
    const int sy = step * (ver ? oh - 1 - oy : oy);
    unsigned px[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int ox = 4 * g + i;
      int sx = step * (hor ? ow - 1 - ox : ox);
      int Y = __ldg(fy + (size_t)sy * w + sx);
      size_t co = (size_t)(sy >> 1) * (w >> 1) + (sx >> 1);
      ChromaOff c = chroma_offsets(__ldg(fu + co), __ldg(fv + co));
      uint2 two = two_pixels((unsigned)Y, c);
      px[i] = two.x;
    }
    __stcs((uint4 *)(out + ((size_t)f * ow * oh + (size_t)oy * ow + 4 * g) * 4), make_uint4(px[0], px[1], px[2], px[3]));
  }
}

struct FmtInfo { int fmt, bpp, ro, go, bo; bool ok; };

FmtInfo fmt_info(uint32_t fourcc)
{
  switch (fourcc) {
  case B200_FOURCC_YUY2: case B200_FOURCC_YUYV: return {0, 2, 0, 0, 0, true};
  case B200_FOURCC_UYVY: return {1, 2, 0, 0, 0, true};
  case B200_FOURCC_NV12: return {2, 0, 0, 0, 0, true};
  case B200_FOURCC_NV21: return {3, 0, 0, 0, 0, true};
  case B200_FOURCC_I422: return {4, 0, 0, 0, 0, true};
  case B200_FOURCC_I420: return {6, 0, 0, 0, 0, true};
  // FOURCC names are little-endian word order; memory order is reversed.
  case B200_FOURCC_ARGB: return {5, 4, 2, 1, 0, true};   // mem B,G,R,A
  case B200_FOURCC_BGRA: return {5, 4, 1, 2, 3, true};   // mem A,R,G,B
  case B200_FOURCC_ABGR: return {5, 4, 0, 1, 2, true};   // mem R,G,B,A
  case B200_FOURCC_RGBA: return {5, 4, 3, 2, 1, true};   // mem A,B,G,R
  case B200_FOURCC_24BG: return {5, 3, 2, 1, 0, true};   // mem B,G,R
  case B200_FOURCC_RAW:  return {5, 3, 0, 1, 2, true};   // mem R,G,B
  default: return {0, 0, 0, 0, 0, false};
  }
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

size_t b200_frame_bytes(uint32_t fourcc, int w, int h)
{
  FmtInfo fi = fmt_info(fourcc);
  if (!fi.ok || w <= 0 || h <= 0) return 0;
  size_t ysz = (size_t)w * h;
  switch (fi.fmt) {
  case 0: case 1: case 4: return ysz * 2;
  case 2: case 3: case 6: return ysz + (ysz >> 1);
  default: return ysz * fi.bpp;
  }
}

int b200_i420_to_rgb32_dev(const uint8_t *d_in, uint8_t *d_out, int w, int h, int n, void *stream)
{
  if (!d_in || !d_out || w <= 0 || h <= 0 || (w & 1) || (h & 1) || n <= 0) {
    set_error("b200_i420_to_rgb32_dev: bad arguments (w=%d h=%d n=%d; w,h must be even)", w, h, n);
    return B200_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if ((w & 3) == 0 && aligned16(d_out) && ((uintptr_t)d_in & 3) == 0) {
    size_t items = (size_t)(w >> 2) * (h >> 1) * n;
    k_i420_to_rgb32_v4<<<grid_for(items, 8), kThreads, 0, s>>>(d_in, d_out, w, h, n);
  } else {
    size_t items = (size_t)(w >> 1) * (h >> 1) * n;
    k_i420_to_rgb32_generic<<<grid_for(items, 8), kThreads, 0, s>>>(d_in, d_out, w, h, n);
  }
  count_launch();
  B200_CHECK(cudaGetLastError(), "i420_to_rgb32 launch");
  return B200_OK;
}

int b200_half_rgb_dev(const uint8_t *d_in, uint8_t *d_out, int w, int h, int n, void *stream)
{
  if (!d_in || !d_out || w <= 0 || h <= 0 || (w & 1) || n <= 0) { set_error("b200_half_rgb_dev: bad arguments (w must be even)"); return B200_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  if ((w & 3) == 0 && aligned16(d_in) && ((uintptr_t)d_out & 7) == 0) {
    size_t items = (size_t)(w >> 2) * ((h + 1) >> 1) * n;
    k_half_rgb<<<grid_for(items, 8), kThreads, 0, s>>>((const uint32_t *)d_in, (uint32_t *)d_out, w, h, n);
  } else {
    size_t items = (size_t)(w >> 1) * ((h + 1) >> 1) * n;
    k_half_rgb_generic<<<grid_for(items, 8), kThreads, 0, s>>>((const uint32_t *)d_in, (uint32_t *)d_out, w, h, n);
  }
  count_launch();
  B200_CHECK(cudaGetLastError(), "half_rgb launch");
  return B200_OK;
}

int b200_flip_rgb_dev(const uint8_t *d_in, uint8_t *d_out, int w, int h, int hor, int ver, int n, void *stream)
{
  if (!d_in || !d_out || w <= 0 || h <= 0 || n <= 0) { set_error("b200_flip_rgb_dev: bad arguments"); return B200_ERR_ARG; }
  if (!hor && !ver) return B200_OK;   // reference leaves the output untouched
  cudaStream_t s = (cudaStream_t)stream;
  if ((w & 3) == 0 && aligned16(d_in) && aligned16(d_out)) {
    size_t items = (size_t)(w >> 2) * h * n;
    k_flip_rgb_v4<<<grid_for(items, 8), kThreads, 0, s>>>((const uint32_t *)d_in, (uint32_t *)d_out, w, h, hor, ver, n);
  } else {
    size_t items = (size_t)w * h * n;
    k_flip_rgb_generic<<<grid_for(items, 8), kThreads, 0, s>>>((const uint32_t *)d_in, (uint32_t *)d_out, w, h, hor, ver, n);
  }
  count_launch();
  B200_CHECK(cudaGetLastError(), "flip_rgb launch");
  return B200_OK;
}

int b200_selfview_dev(const uint8_t *d_i420, uint8_t *d_out, int w, int h, int half, int hor, int ver, int n, void *stream)
{
  const int ow = half ? w >> 1 : w;
  if (!d_i420 || !d_out || w <= 0 || h <= 0 || (w & 1) || (h & 1) || (ow & 3) || n <= 0 || !aligned16(d_out)) {
    set_error("b200_selfview_dev: bad arguments (w=%d h=%d; output width must be a multiple of 4)", w, h);
    return B200_ERR_ARG;
  }
  const int oh = half ? (h + 1) >> 1 : h;
  size_t items = (size_t)(ow >> 2) * oh * n;
  k_selfview<<<grid_for(items, 8), kThreads, 0, (cudaStream_t)stream>>>(d_i420, d_out, w, h, half, hor, ver, n);
  count_launch();
  B200_CHECK(cudaGetLastError(), "selfview launch");
  return B200_OK;
}

int b200_convert_to_i420_dev(const uint8_t *d_src, uint8_t *d_dst, int w, int h, uint32_t fourcc, int n, void *stream)
{
  FmtInfo fi = fmt_info(fourcc);
  if (!fi.ok) { set_error("b200_convert_to_i420_dev: unsupported fourcc 0x%08x", fourcc); return B200_ERR_ARG; }
  if (!d_src || !d_dst || w <= 0 || h <= 0 || (w & 1) || (h & 1) || n <= 0) {
    set_error("b200_convert_to_i420_dev: bad arguments (w=%d h=%d n=%d; w,h must be even)", w, h, n);
    return B200_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const bool al = aligned16(d_src) && aligned16(d_dst);
  const size_t fin = b200_frame_bytes(fourcc, w, h);
  if (fi.fmt == 6) {
    B200_CHECK(cudaMemcpyAsync(d_dst, d_src, fin * n, cudaMemcpyDeviceToDevice, s), "i420 copy");
    return B200_OK;
  }
  if (fi.fmt <= 1 && (w & 7) == 0 && al) {
    size_t items = (size_t)(w >> 3) * (h >> 1) * n;
    if (fi.fmt == 0) k_packed422_to_i420_v8<true><<<grid_for(items), kThreads, 0, s>>>(d_src, d_dst, w, h, n);
    else             k_packed422_to_i420_v8<false><<<grid_for(items), kThreads, 0, s>>>(d_src, d_dst, w, h, n);
  } else if (fi.fmt >= 2 && fi.fmt <= 4 && (w & 15) == 0 && al && (((size_t)w * h) & 31) == 0) {
    size_t items = ((size_t)w * h / 16 + (size_t)(w >> 4) * (h >> 1)) * n;
    if (fi.fmt == 2)      k_planar_to_i420_v16<0><<<grid_for(items), kThreads, 0, s>>>(d_src, d_dst, w, h, n);
    else if (fi.fmt == 3) k_planar_to_i420_v16<1><<<grid_for(items), kThreads, 0, s>>>(d_src, d_dst, w, h, n);
    else                  k_planar_to_i420_v16<2><<<grid_for(items), kThreads, 0, s>>>(d_src, d_dst, w, h, n);
  } else if (fi.fmt == 5 && (w & 3) == 0 && al) {
    size_t items = (size_t)(w >> 2) * (h >> 1) * n;
    const size_t tiles24 = (size_t)((w + kRgb24TilePx - 1) / kRgb24TilePx) * (h >> 1) * n;
    if (fi.bpp == 4) k_rgb_to_i420_v4<4><<<grid_for(items), kThreads, 0, s>>>(d_src, d_dst, w, h, n, fi.ro, fi.go, fi.bo);
    else if ((w & 15) == 0 && tiles24 < 0x7fffffffull)
      k_rgb24_to_i420_staged<<<grid_for(tiles24 * 32), kThreads, 0, s>>>(d_src, d_dst, w, h, n, fi.ro, fi.go, fi.bo);
    else             k_rgb_to_i420_v4<3><<<grid_for(items), kThreads, 0, s>>>(d_src, d_dst, w, h, n, fi.ro, fi.go, fi.bo);
  } else {
    size_t items = (size_t)(w >> 1) * (h >> 1) * n;
    k_to_i420_generic<<<grid_for(items), kThreads, 0, s>>>(d_src, d_dst, w, h, n, fi.fmt, fi.bpp, fi.ro, fi.go, fi.bo, fin);
  }
  count_launch();
  B200_CHECK(cudaGetLastError(), "convert_to_i420 launch");
  return B200_OK;
}

// ---- host-buffer entry points ---------------------------------------------

static int run_host(const uint8_t *in, size_t in_bytes, uint8_t *out, size_t out_bytes,
                    int (*launch)(const uint8_t *, uint8_t *, void *, void *), void *ctx)
{
  if (b200_device_count() <= 0) { set_error("no CUDA device: libb200media has no CPU fallback"); return B200_ERR_CUDA; }
  Scratch &sc = scratch();
  if (!sc.ensure(in_bytes, out_bytes)) return B200_ERR_CUDA;
  memcpy(sc.h_in, in, in_bytes);
  B200_CHECK(cudaMemcpyAsync(sc.d_in, sc.h_in, in_bytes, cudaMemcpyHostToDevice, sc.stream), "H2D");
  int rc = launch(sc.d_in, sc.d_out, sc.stream, ctx);
  if (rc != B200_OK) return rc;
  B200_CHECK(cudaMemcpyAsync(sc.h_out, sc.d_out, out_bytes, cudaMemcpyDeviceToHost, sc.stream), "D2H");
  B200_CHECK(cudaStreamSynchronize(sc.stream), "sync");
  memcpy(out, sc.h_out, out_bytes);
  return B200_OK;
}

struct WH { int w, h, a, b; uint32_t fourcc; };

int b200_yuv420_to_rgb32(const uint8_t *input, uint8_t *output, uint16_t width, uint16_t height)
{
  if (!input || !output || !width || !height || (width & 1) || (height & 1)) {
    set_error("b200_yuv420_to_rgb32: bad arguments");
    return B200_ERR_ARG;
  }
  WH wh{width, height, 0, 0, 0};
  size_t px = (size_t)width * height;
  int rc = run_host(input, px + px / 2, output, px * 4,
                    [](const uint8_t *di, uint8_t *dout, void *s, void *c) {
                      WH *p = (WH *)c;
                      return b200_i420_to_rgb32_dev(di, dout, p->w, p->h, 1, s);
                    }, &wh);
  return rc == B200_OK ? 1 : rc;   // the reference returns 1 (yuvconversions.cpp:167)
}

int b200_half_rgb(const uint8_t *input, uint8_t *output, uint16_t width, uint16_t height)
{
  if (!input || !output || !width || !height) { set_error("b200_half_rgb: bad arguments"); return B200_ERR_ARG; }
  WH wh{width, height, 0, 0, 0};
  size_t px = (size_t)width * height;
  size_t opx = (size_t)(width / 2) * ((height + 1) / 2);
  return run_host(input, px * 4, output, opx * 4,
                  [](const uint8_t *di, uint8_t *dout, void *s, void *c) {
                    WH *p = (WH *)c;
                    return b200_half_rgb_dev(di, dout, p->w, p->h, 1, s);
                  }, &wh);
}

int b200_flip_rgb(const uint8_t *input, uint8_t *output, uint16_t width, uint16_t height, int hor, int ver)
{
  if (!input || !output || !width || !height) { set_error("b200_flip_rgb: bad arguments"); return B200_ERR_ARG; }
  if (!hor && !ver) return B200_OK;
  WH wh{width, height, hor, ver, 0};
  size_t px = (size_t)width * height;
  return run_host(input, px * 4, output, px * 4,
                  [](const uint8_t *di, uint8_t *dout, void *s, void *c) {
                    WH *p = (WH *)c;
                    return b200_flip_rgb_dev(di, dout, p->w, p->h, p->a, p->b, 1, s);
                  }, &wh);
}

int b200_selfview(const uint8_t *input, uint8_t *output, uint16_t width, uint16_t height, int half, int hor, int ver)
{
  if (!input || !output || !width || !height) { set_error("b200_selfview: bad arguments"); return B200_ERR_ARG; }
  WH wh{width, height, half | (hor << 1) | (ver << 2), 0, 0};
  size_t px = (size_t)width * height;
  size_t opx = half ? (size_t)(width / 2) * ((height + 1) / 2) : px;
  return run_host(input, px + px / 2, output, opx * 4,
                  [](const uint8_t *di, uint8_t *dout, void *s, void *c) {
                    WH *p = (WH *)c;
                    return b200_selfview_dev(di, dout, p->w, p->h, p->a & 1, (p->a >> 1) & 1, (p->a >> 2) & 1, 1, s);
                  }, &wh);
}

int b200_ConvertToI420(const uint8_t *sample, size_t sample_size,
                       uint8_t *dst_y, int sy, uint8_t *dst_u, int su, uint8_t *dst_v, int sv,
                       int crop_x, int crop_y, int src_w, int src_h, int crop_w, int crop_h,
                       int rotation, uint32_t fourcc)
{
  FmtInfo fi = fmt_info(fourcc);
  const bool mjpg = fourcc == B200_FOURCC_MJPG;
  if (!fi.ok && !mjpg) { set_error("b200_ConvertToI420: unsupported fourcc 0x%08x", fourcc); return -1; }
  if (!sample || !dst_y || !dst_u || !dst_v || src_w <= 0 || src_h <= 0) { set_error("b200_ConvertToI420: bad arguments"); return -1; }
  if (crop_x || crop_y || crop_w != src_w || crop_h != src_h || rotation != 0) {
    set_error("b200_ConvertToI420: crop/rotation not supported (the reference never uses them)");
    return -1;
  }
  if ((src_w & 1) || (src_h & 1)) { set_error("b200_ConvertToI420: odd dimensions not supported"); return -1; }
  size_t need = mjpg ? sample_size : b200_frame_bytes(fourcc, src_w, src_h);
  if (mjpg && !sample_size) { set_error("b200_ConvertToI420: an MJPG sample needs its size"); return -1; }
  if (sample_size && sample_size < need) { set_error("b200_ConvertToI420: sample_size %zu < %zu", sample_size, need); return -1; }
  if (b200_device_count() <= 0) { set_error("no CUDA device: libb200media has no CPU fallback"); return B200_ERR_CUDA; }
  size_t ysz = (size_t)src_w * src_h, out_bytes = ysz + ysz / 2;
  Scratch &sc = scratch();
  if (!sc.ensure(mjpg ? 16 : need, out_bytes)) return B200_ERR_CUDA;
  int rc;
  if (mjpg) {
    // the scan is read on the host straight from the caller's buffer; the GPU gets coefficients (mjpg.cu)
    rc = b200_mjpg_to_i420_dev(sample, sample_size, sc.d_out, src_w, src_h, sc.stream);
    if (rc != B200_OK) return -1;                       // libyuv::MJPGToI420 fails likewise: the output stays untouched
  } else {
    memcpy(sc.h_in, sample, need);
    B200_CHECK(cudaMemcpyAsync(sc.d_in, sc.h_in, need, cudaMemcpyHostToDevice, sc.stream), "H2D");
    rc = b200_convert_to_i420_dev(sc.d_in, sc.d_out, src_w, src_h, fourcc, 1, sc.stream);
  }
  if (rc != B200_OK) return rc;
  B200_CHECK(cudaMemcpyAsync(sc.h_out, sc.d_out, out_bytes, cudaMemcpyDeviceToHost, sc.stream), "D2H");
  B200_CHECK(cudaStreamSynchronize(sc.stream), "sync");
  const int cw = src_w / 2, ch = src_h / 2;
  for (int j = 0; j < src_h; j++) memcpy(dst_y + (size_t)j * sy, sc.h_out + (size_t)j * src_w, src_w);
  for (int j = 0; j < ch; j++) {
    memcpy(dst_u + (size_t)j * su, sc.h_out + ysz + (size_t)j * cw, cw);
    memcpy(dst_v + (size_t)j * sv, sc.h_out + ysz + ysz / 4 + (size_t)j * cw, cw);
  }
  return 0;
}

}  // extern "C"
