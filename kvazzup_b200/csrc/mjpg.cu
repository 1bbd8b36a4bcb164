// MJPG camera frames -> I420 (the thirteenth format of LibYUVConverter:
// reference src/media/processing/libyuvconverter.cpp:94,120-127 -> libyuv::MJPGToI420).
//
// Split the way hybrid JPEG decoders are: the Huffman-coded scan of a camera frame is one serial bit
// stream (UVC cameras emit no restart markers), so the host reads it -- a table-driven decoder over a
// 64-bit window, one pass, straight into a page-locked coefficient buffer -- and the GPU does the
// arithmetic: dequantisation, libjpeg's accurate integer inverse DCT (the one libyuv's decoder runs,
// JDCT_ISLOW), level shift and clamp (k_mjpg_idct: eight lanes per 8x8 block, rows loaded as 16-byte
// vectors, the column pass and the row pass meet in shared memory), then the conversion of the frame's
// subsampling to 4:2:0 exactly as libyuv does it (k_mjpg_to_i420: 4:2:2 averages chroma row pairs with
// round-half-up, 4:4:4 takes the rounded 2x2 box, 4:0:0 writes 128).  The result stays in HBM for the
// encoder (b200_mjpg_to_i420_dev) or returns to the host (b200_ConvertToI420 with FOURCC_MJPG).
#include <string.h>

#include <vector>

#include "../../include/b200media.h"
#include "runtime.h"

namespace b200 {

namespace {

// ---- host: JPEG headers and the entropy-coded scan ------------------------------------------------

struct HuffLut {
  // codes of up to kFast bits resolve in one lookup: entry = length << 8 | symbol (0 = longer code)
  static constexpr int kFast = 10;
  uint16_t fast[1 << kFast];
  // longer codes: canonical decoding, first code / first value index of every length
  int32_t maxcode[18];
  int32_t delta[17];
  uint8_t vals[256];
  bool ok = false;

  void build(const uint8_t *counts, const uint8_t *symbols, int n)
  {
    memcpy(vals, symbols, (size_t)n);
    memset(fast, 0, sizeof(fast));
    int code = 0, k = 0;
    ok = false;
    for (int len = 1; len <= 16; len++) {
      delta[len] = k - code;
      if (code + counts[len - 1] > (1 << len) || k + counts[len - 1] > n) return;      // over-subscribed (corrupt DHT): the table stays unusable
      for (int i = 0; i < counts[len - 1]; i++, code++, k++)
        if (len <= kFast) {
          const int lo = code << (kFast - len);
          for (int f = 0; f < (1 << (kFast - len)); f++) fast[lo + f] = (uint16_t)((len << 8) | symbols[k]);
        }
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    ok = true;
  }
};

// standard tables of T.81 Annex K.3 (frames without DHT)
const uint8_t kStdDcLumCounts[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kStdDcChrCounts[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kStdDcSymbols[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kStdAcLumCounts[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125};
const uint8_t kStdAcChrCounts[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119};
const uint8_t kStdAcLumSymbols[162] = {
  1, 2, 3, 0, 4, 17, 5, 18, 33, 49, 65, 6, 19, 81, 97, 7, 34, 113, 20, 50, 129, 145, 161, 8, 35, 66, 177, 193, 21, 82, 209, 240, 36, 51, 98,
  114, 130, 9, 10, 22, 23, 24, 25, 26, 37, 38, 39, 40, 41, 42, 52, 53, 54, 55, 56, 57, 58, 67, 68, 69, 70, 71, 72, 73, 74, 83, 84, 85, 86, 87,
  88, 89, 90, 99, 100, 101, 102, 103, 104, 105, 106, 115, 116, 117, 118, 119, 120, 121, 122, 131, 132, 133, 134, 135, 136, 137, 138, 146,
  147, 148, 149, 150, 151, 152, 153, 154, 162, 163, 164, 165, 166, 167, 168, 169, 170, 178, 179, 180, 181, 182, 183, 184, 185, 186, 194,
  195, 196, 197, 198, 199, 200, 201, 202, 210, 211, 212, 213, 214, 215, 216, 217, 218, 225, 226, 227, 228, 229, 230, 231, 232, 233, 234,
  241, 242, 243, 244, 245, 246, 247, 248, 249, 250};
const uint8_t kStdAcChrSymbols[162] = {
  0, 1, 2, 3, 17, 4, 5, 33, 49, 6, 18, 65, 81, 7, 97, 113, 19, 34, 50, 129, 8, 20, 66, 145, 161, 177, 193, 9, 35, 51, 82, 240, 21, 98, 114,
  209, 10, 22, 36, 52, 225, 37, 241, 23, 24, 25, 26, 38, 39, 40, 41, 42, 53, 54, 55, 56, 57, 58, 67, 68, 69, 70, 71, 72, 73, 74, 83, 84, 85,
  86, 87, 88, 89, 90, 99, 100, 101, 102, 103, 104, 105, 106, 115, 116, 117, 118, 119, 120, 121, 122, 130, 131, 132, 133, 134, 135, 136,
  137, 138, 146, 147, 148, 149, 150, 151, 152, 153, 154, 162, 163, 164, 165, 166, 167, 168, 169, 170, 178, 179, 180, 181, 182, 183, 184,
  185, 186, 194, 195, 196, 197, 198, 199, 200, 201, 202, 210, 211, 212, 213, 214, 215, 216, 217, 218, 226, 227, 228, 229, 230, 231, 232,
  233, 234, 242, 243, 244, 245, 246, 247, 248, 249, 250};

// position in the row-major block of the k-th coefficient in zig-zag order
const uint8_t kNatural[64 + 16] = {
  0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};          // a corrupt run may step past 63: stay inside the block

struct Component { int id, hs, vs, tq, td, ta, bw, bh; size_t first_block; };

struct FrameHeader {
  int w = 0, h = 0, nc = 0, restart = 0, mcu_cols = 0, mcu_rows = 0;
  Component c[3];
  uint16_t q[4][64];                  // row-major
  bool have_q[4] = {false, false, false, false};
  size_t scan = 0, blocks = 0;
};

// The window holds the next bits of the scan with byte stuffing already removed; a marker ends the
// segment and the window is fed zeros (a truncated frame decodes to grey instead of reading past the end).
struct BitWindow {
  const uint8_t *p, *end;
  uint64_t acc = 0;
  int n = 0;
  bool at_marker = false;
  void fill()
  {
    while (n <= 56) {
      unsigned b = 0;
      if (!at_marker && p < end) {
        b = *p++;
        if (b == 0xff) {
          if (p < end && *p == 0) p++;
          else { at_marker = true; p--; b = 0; }
        }
      }
      acc |= (uint64_t)b << (56 - n);
      n += 8;
    }
  }
  unsigned peek(int k) const { return (unsigned)(acc >> (64 - k)); }
  void drop(int k) { acc <<= k; n -= k; }
  int symbol(const HuffLut &t)
  {
    if (n < 16) fill();
    const unsigned e = t.fast[peek(HuffLut::kFast)];
    if (e) { drop((int)(e >> 8)); return (int)(e & 0xff); }
    int code = (int)peek(HuffLut::kFast);
    for (int len = HuffLut::kFast + 1; len <= 16; len++) {
      code = (int)peek(len);
      if (code <= t.maxcode[len]) { drop(len); return t.vals[(code + t.delta[len]) & 0xff]; }
    }
    return -1;
  }
  int value(int size)                 // `size` raw bits, sign-extended the T.81 way (F.2.2.1)
  {
    if (!size) return 0;
    if (n < size) fill();
    const int v = (int)peek(size);
    drop(size);
    return v < (1 << (size - 1)) ? v - (1 << size) + 1 : v;
  }
  void restart_marker()               // byte-align and step over RSTn
  {
    acc = 0; n = 0;
    if (at_marker) { if (p + 1 < end && p[1] >= 0xd0 && p[1] <= 0xd7) p += 2; at_marker = false; }
    else if (p + 1 < end && p[0] == 0xff && p[1] >= 0xd0 && p[1] <= 0xd7) p += 2;
  }
};

bool parse_headers(const uint8_t *d, size_t n, FrameHeader &f, HuffLut (&dc)[4], HuffLut (&ac)[4])
{
  if (n < 4 || d[0] != 0xff || d[1] != 0xd8) { set_error("mjpg: no SOI marker"); return false; }
  size_t p = 2;
  int hmax = 1, vmax = 1;
  while (true) {
    while (p < n && d[p] != 0xff) p++;
    while (p < n && d[p] == 0xff) p++;
    if (p >= n) { set_error("mjpg: no scan in the frame"); return false; }
    const int m = d[p++];
    if (m == 0xd9) { set_error("mjpg: no scan in the frame"); return false; }
    if ((m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;
    if (p + 2 > n) { set_error("mjpg: truncated segment"); return false; }
    const size_t len = ((size_t)d[p] << 8) | d[p + 1];
    if (len < 2 || p + len > n) { set_error("mjpg: truncated segment"); return false; }
    const uint8_t *s = d + p + 2;
    const size_t sl = len - 2;
    switch (m) {
    case 0xdb:
      for (size_t o = 0; o < sl;) {
        const int wide = s[o] >> 4, id = s[o] & 15;
        o++;
        if (id > 3 || o + (wide ? 128u : 64u) > sl) { set_error("mjpg: bad DQT"); return false; }
        for (int i = 0; i < 64; i++, o += wide ? 2 : 1) f.q[id][kNatural[i]] = wide ? (uint16_t)((s[o] << 8) | s[o + 1]) : s[o];
        f.have_q[id] = true;
      }
      break;
    case 0xc4:
      for (size_t o = 0; o + 17 <= sl;) {
        const int cls = s[o] >> 4, id = s[o] & 15;
        int cnt = 0;
        for (int i = 0; i < 16; i++) cnt += s[o + 1 + i];
        if (cls > 1 || id > 3 || cnt > 256 || o + 17 + (size_t)cnt > sl) { set_error("mjpg: bad DHT"); return false; }
        (cls ? ac[id] : dc[id]).build(s + o + 1, s + o + 17, cnt);
        o += 17 + (size_t)cnt;
      }
      break;
    case 0xdd:
      if (sl < 2) { set_error("mjpg: bad DRI"); return false; }
      f.restart = (s[0] << 8) | s[1];
      break;
    case 0xc0: case 0xc1:
      if (sl < 6 || s[0] != 8) { set_error("mjpg: only 8-bit baseline frames are supported"); return false; }
      f.h = (s[1] << 8) | s[2]; f.w = (s[3] << 8) | s[4]; f.nc = s[5];
      if ((f.nc != 1 && f.nc != 3) || sl < 6 + 3 * (size_t)f.nc || f.w <= 0 || f.h <= 0 || f.w > 8192 || f.h > 8192) {
        set_error("mjpg: unsupported frame header (%d components, %dx%d)", f.nc, f.w, f.h);
        return false;
      }
      for (int i = 0; i < f.nc; i++) {
        Component &c = f.c[i];
        c.id = s[6 + 3 * i]; c.hs = s[7 + 3 * i] >> 4; c.vs = s[7 + 3 * i] & 15; c.tq = s[8 + 3 * i];
        if (c.hs < 1 || c.hs > 2 || c.vs < 1 || c.vs > 2 || c.tq > 3) { set_error("mjpg: unsupported sampling factors"); return false; }
        hmax = c.hs > hmax ? c.hs : hmax; vmax = c.vs > vmax ? c.vs : vmax;
      }
      break;
    case 0xda: {
      if (!f.nc || sl < 1 || s[0] != f.nc || sl < 1 + 2 * (size_t)f.nc + 3) { set_error("mjpg: scans that do not hold every component are not supported"); return false; }
      for (int i = 0; i < f.nc; i++) {
        if (s[1 + 2 * i] != f.c[i].id) { set_error("mjpg: scan component order differs from the frame header"); return false; }
        f.c[i].td = s[2 + 2 * i] >> 4; f.c[i].ta = s[2 + 2 * i] & 15;
        if (f.c[i].td > 3 || f.c[i].ta > 3) { set_error("mjpg: bad table selector"); return false; }
      }
      f.scan = p + len;
      if (f.nc == 1) { f.c[0].hs = f.c[0].vs = 1; hmax = vmax = 1; }       // a lone component is not interleaved
      f.mcu_cols = (f.w + 8 * hmax - 1) / (8 * hmax); f.mcu_rows = (f.h + 8 * vmax - 1) / (8 * vmax);
      size_t first = 0;
      for (int i = 0; i < f.nc; i++) {
        Component &c = f.c[i];
        c.bw = f.mcu_cols * c.hs; c.bh = f.mcu_rows * c.vs; c.first_block = first;
        first += (size_t)c.bw * c.bh;
      }
      f.blocks = first;
      return true;
    }
    default:
      if (m == 0xc2 || (m >= 0xc3 && m <= 0xcf && m != 0xc8 && m != 0xcc)) { set_error("mjpg: only baseline (SOF0) frames are supported"); return false; }
      break;
    }
    p += len;
  }
}

// Every block's 64 coefficients, row-major, component after component, blocks in raster order.
bool read_scan(const uint8_t *d, size_t n, const FrameHeader &f, const HuffLut (&dc)[4], const HuffLut (&ac)[4], int16_t *coef)
{
  memset(coef, 0, f.blocks * 64 * sizeof(int16_t));
  BitWindow b{d + f.scan, d + n};
  int pred[3] = {0, 0, 0}, left = f.restart;
  for (int my = 0; my < f.mcu_rows; my++)
    for (int mx = 0; mx < f.mcu_cols; mx++) {
      if (f.restart && left == 0) { b.restart_marker(); pred[0] = pred[1] = pred[2] = 0; left = f.restart; }
      left--;
      for (int i = 0; i < f.nc; i++) {
        const Component &c = f.c[i];
        const HuffLut &tdc = dc[c.td], &tac = ac[c.ta];
        for (int by = 0; by < c.vs; by++)
          for (int bx = 0; bx < c.hs; bx++) {
            int16_t *blk = coef + (c.first_block + (size_t)(my * c.vs + by) * c.bw + mx * c.hs + bx) * 64;
            const int t = b.symbol(tdc);
            if (t < 0 || t > 11) { set_error("mjpg: corrupt scan (DC)"); return false; }
            pred[i] += b.value(t);
            blk[0] = (int16_t)pred[i];
            for (int k = 1; k < 64; k++) {
              const int rs = b.symbol(tac);
              if (rs < 0) { set_error("mjpg: corrupt scan (AC)"); return false; }
              const int size = rs & 15;
              if (!size) { if (rs != 0xf0) break; k += 15; continue; }
              k += rs >> 4;
              blk[kNatural[k]] = (int16_t)b.value(size);
            }
          }
      }
    }
  return true;
}

// ---- device ------------------------------------------------------------------------------------------

struct IdctGeom {
  uint32_t first_block[4];            // first block of every component, [nc] = total
  int bw[3];                          // blocks per row
  uint32_t plane_off[3];              // byte offset of the component's padded plane
  int tq[3];
};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 8-point pass of libjpeg's accurate integer IDCT (Loeffler-Ligtenberg-Moschytz, 13-bit constants)
__device__ __forceinline__ void islow_pass(const int (&in)[8], int (&out)[8], int shift)
{
  const int z1 = (in[2] + in[6]) * 4433;
  const int t2 = z1 - in[6] * 15137, t3 = z1 + in[2] * 6270;
  const int t0 = (in[0] + in[4]) << 13, t1 = (in[0] - in[4]) << 13;
  const int e0 = t0 + t3, e3 = t0 - t3, e1 = t1 + t2, e2 = t1 - t2;
  int o0 = in[7], o1 = in[5], o2 = in[3], o3 = in[1];
  int y1 = o0 + o3, y2 = o1 + o2, y3 = o0 + o2, y4 = o1 + o3;
  const int y5 = (y3 + y4) * 9633;
  o0 *= 2446; o1 *= 16819; o2 *= 25172; o3 *= 12299;
  y1 *= -7373; y2 *= -20995; y3 = y3 * -16069 + y5; y4 = y4 * -3196 + y5;
  o0 += y1 + y3; o1 += y2 + y4; o2 += y2 + y3; o3 += y1 + y4;
  out[0] = descale(e0 + o3, shift); out[7] = descale(e0 - o3, shift);
  out[1] = descale(e1 + o2, shift); out[6] = descale(e1 - o2, shift);
  out[2] = descale(e2 + o1, shift); out[5] = descale(e2 - o1, shift);
  out[3] = descale(e3 + o0, shift); out[4] = descale(e3 - o0, shift);
}

constexpr int kIdctThreads = 256, kBlocksPerCta = kIdctThreads / 8;

// Eight lanes per block.  Lane j loads row j of the coefficients (one 16-byte vector) and the matching
// row of the quantisation table, the dequantised block goes to shared memory (pitch 9: the column reads
// that follow hit eight different banks), lane j transforms column j, writes it back, then transforms
// row j and stores its eight samples as one 8-byte vector.
__global__ void __launch_bounds__(kIdctThreads)
k_mjpg_idct(const int16_t *__restrict__ coef, const uint16_t *__restrict__ qtab, IdctGeom g, int nc, uint8_t *__restrict__ planes)
{
  __shared__ int s_blk[kBlocksPerCta][8 * 9];
  __shared__ uint16_t s_q[4][64];
  for (int i = threadIdx.x; i < 4 * 64; i += kIdctThreads) s_q[i >> 6][i & 63] = qtab[i];
  __syncthreads();
  const int j = threadIdx.x & 7, lb = threadIdx.x >> 3;
  const uint32_t total = g.first_block[nc];
  for (uint32_t blk = blockIdx.x * kBlocksPerCta + lb; blk < total + lb; blk += gridDim.x * kBlocksPerCta) {
    // (the loop bound keeps the eight lanes of a block together and every warp's __syncwarp matched)
    const bool live = blk < total;
    const int c = !live ? 0 : (blk >= g.first_block[2] && nc > 2 ? 2 : (blk >= g.first_block[1] && nc > 1 ? 1 : 0));
    int *sb = s_blk[lb];
    if (live) {
      const uint4 v = __ldg((const uint4 *)(coef + (size_t)blk * 64) + j);
      const uint16_t *q = s_q[g.tq[c]] + 8 * j;
      const int16_t *cv = (const int16_t *)&v;
#pragma unroll
      for (int i = 0; i < 8; i++) sb[j * 9 + i] = (int)cv[i] * (int)q[i];
    }
    __syncwarp();
    int in[8], out[8];
    if (live) {
#pragma unroll
      for (int r = 0; r < 8; r++) in[r] = sb[r * 9 + j];
      islow_pass(in, out, 11);
    }
    __syncwarp();
    if (live) {
#pragma unroll
      for (int r = 0; r < 8; r++) sb[r * 9 + j] = out[r];
    }
    __syncwarp();
    if (live) {
#pragma unroll
      for (int i = 0; i < 8; i++) in[i] = sb[j * 9 + i];
      islow_pass(in, out, 18);
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        lo |= (uint32_t)min(max(out[i] + 128, 0), 255) << (8 * i);
        hi |= (uint32_t)min(max(out[i + 4] + 128, 0), 255) << (8 * i);
      }
      const uint32_t rel = blk - g.first_block[c];
      const int by = (int)(rel / (uint32_t)g.bw[c]), bx = (int)(rel - (uint32_t)by * g.bw[c]);
      *(uint2 *)(planes + g.plane_off[c] + (size_t)(by * 8 + j) * (g.bw[c] * 8) + bx * 8) = make_uint2(lo, hi);
    }
    __syncwarp();
  }
}

// mode 0: 4:2:0 (copy), 1: 4:2:2 (row pairs averaged), 2: 4:4:4 (2x2 box), 3: 4:0:0 (chroma 128).
// One thread per four output samples of a row.
__global__ void __launch_bounds__(256)
k_mjpg_to_i420(const uint8_t *__restrict__ planes, IdctGeom g, int mode, int w, int h, uint8_t *__restrict__ out)
{
  const int cw = w >> 1, ch = h >> 1;
  const int yq = (w + 3) >> 2, cq = (cw + 3) >> 2;
  const int n_y = yq * h, n_c = cq * ch;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_y + 2 * n_c) return;
  if (i < n_y) {
    const int y = i / yq, x = 4 * (i - y * yq);
    const uint8_t *s = planes + g.plane_off[0] + (size_t)y * (g.bw[0] * 8) + x;
    uint8_t *d = out + (size_t)y * w + x;
    for (int k = 0; k < 4 && x + k < w; k++) d[k] = s[k];
    return;
  }
  const int k = i - n_y, c = k >= n_c ? 2 : 1, r = k - (c - 1) * n_c;
  const int y = r / cq, x = 4 * (r - y * cq);
  uint8_t *d = out + (size_t)w * h + (size_t)(c - 1) * cw * ch + (size_t)y * cw + x;
  if (mode == 3) { for (int t = 0; t < 4 && x + t < cw; t++) d[t] = 128; return; }
  const int st = g.bw[c] * 8;
  const uint8_t *s = planes + g.plane_off[c];
  for (int t = 0; t < 4 && x + t < cw; t++) {
    const int xx = x + t;
    int v;
    if (mode == 0) v = s[(size_t)y * st + xx];
    else if (mode == 1) v = (s[(size_t)(2 * y) * st + xx] + s[(size_t)(2 * y + 1) * st + xx] + 1) >> 1;
    else v = (s[(size_t)(2 * y) * st + 2 * xx] + s[(size_t)(2 * y) * st + 2 * xx + 1] + s[(size_t)(2 * y + 1) * st + 2 * xx] +
              s[(size_t)(2 * y + 1) * st + 2 * xx + 1] + 2) >> 2;
    d[t] = (uint8_t)v;
  }
}

// per-thread staging: page-locked coefficients, device coefficients / tables / padded planes
struct MjpgScratch {
  int16_t *h_coef = nullptr, *d_coef = nullptr;
  uint16_t *h_q = nullptr, *d_q = nullptr;
  uint8_t *d_planes = nullptr;
  size_t coef_cap = 0, plane_cap = 0;
  cudaEvent_t done = nullptr;
  bool ensure(size_t coef_bytes, size_t plane_bytes)
  {
    if (!h_q) {
      if (!cuda_ok(cudaMallocHost(&h_q, 4 * 64 * sizeof(uint16_t)), "mjpg tables") || !cuda_ok(cudaMalloc(&d_q, 4 * 64 * sizeof(uint16_t)), "mjpg tables") ||
          !cuda_ok(cudaEventCreateWithFlags(&done, cudaEventDisableTiming), "mjpg event")) return false;
    }
    if (coef_bytes > coef_cap) {
      if (h_coef) cudaFreeHost(h_coef);
      if (d_coef) cudaFree(d_coef);
      h_coef = nullptr; d_coef = nullptr; coef_cap = 0;
      if (!cuda_ok(cudaMallocHost(&h_coef, coef_bytes), "mjpg coefficients") || !cuda_ok(cudaMalloc(&d_coef, coef_bytes), "mjpg coefficients")) return false;
      coef_cap = coef_bytes;
    }
    if (plane_bytes > plane_cap) {
      if (d_planes) cudaFree(d_planes);
      d_planes = nullptr; plane_cap = 0;
      if (!cuda_ok(cudaMalloc(&d_planes, plane_bytes), "mjpg planes")) return false;
      plane_cap = plane_bytes;
    }
    return true;
  }
  ~MjpgScratch()
  {
    if (h_coef) cudaFreeHost(h_coef);
    if (d_coef) cudaFree(d_coef);
    if (h_q) cudaFreeHost(h_q);
    if (d_q) cudaFree(d_q);
    if (d_planes) cudaFree(d_planes);
    if (done) cudaEventDestroy(done);
  }
};

}  // namespace

}  // namespace b200

using namespace b200;

// Host only (no CUDA call): is this a frame b200_mjpg_to_i420_dev can convert?  Parses the headers and reads
// the whole entropy-coded scan, as the conversion does.
extern "C" int b200_mjpg_probe(const uint8_t *jpeg, size_t jpeg_bytes, int *width, int *height, int *subsampling)
{
  if (!jpeg) { set_error("b200_mjpg_probe: bad arguments"); return B200_ERR_ARG; }
  static thread_local HuffLut dc[4], ac[4];
  for (int i = 0; i < 4; i++) dc[i].ok = ac[i].ok = false;
  FrameHeader f;
  if (!parse_headers(jpeg, jpeg_bytes, f, dc, ac)) return B200_ERR_ARG;
  if (!dc[0].ok && !ac[0].ok) {
    dc[0].build(kStdDcLumCounts, kStdDcSymbols, 12); dc[1].build(kStdDcChrCounts, kStdDcSymbols, 12);
    ac[0].build(kStdAcLumCounts, kStdAcLumSymbols, 162); ac[1].build(kStdAcChrCounts, kStdAcChrSymbols, 162);
  }
  int mode;
  if (f.nc == 1) mode = 400;
  else if (f.c[1].hs != 1 || f.c[1].vs != 1 || f.c[2].hs != 1 || f.c[2].vs != 1) { set_error("mjpg: unsupported chroma sampling"); return B200_ERR_ARG; }
  else if (f.c[0].hs == 2 && f.c[0].vs == 2) mode = 420;
  else if (f.c[0].hs == 2 && f.c[0].vs == 1) mode = 422;
  else if (f.c[0].hs == 1 && f.c[0].vs == 1) mode = 444;
  else { set_error("mjpg: unsupported luma sampling %dx%d", f.c[0].hs, f.c[0].vs); return B200_ERR_ARG; }
  for (int i = 0; i < f.nc; i++)
    if (!f.have_q[f.c[i].tq] || !dc[f.c[i].td].ok || !ac[f.c[i].ta].ok) { set_error("mjpg: a table the scan names is missing"); return B200_ERR_ARG; }
  std::vector<int16_t> coef(f.blocks * 64);
  if (!read_scan(jpeg, jpeg_bytes, f, dc, ac, coef.data())) return B200_ERR_ARG;
  if (width) *width = f.w;
  if (height) *height = f.h;
  if (subsampling) *subsampling = mode;
  return B200_OK;
}

extern "C" int b200_mjpg_to_i420_dev(const uint8_t *jpeg, size_t jpeg_bytes, uint8_t *d_i420, int w, int h, void *stream)
{
  if (!jpeg || !d_i420 || w <= 0 || h <= 0 || (w & 1) || (h & 1)) { set_error("b200_mjpg_to_i420_dev: bad arguments (w=%d h=%d; w,h must be even)", w, h); return B200_ERR_ARG; }
  if (b200_device_count() <= 0) { set_error("no CUDA device: libb200media has no CPU fallback"); return B200_ERR_CUDA; }
  static thread_local HuffLut dc[4], ac[4];
  static thread_local MjpgScratch sc;
  for (int i = 0; i < 4; i++) dc[i].ok = ac[i].ok = false;
  FrameHeader f;
  if (!parse_headers(jpeg, jpeg_bytes, f, dc, ac)) return B200_ERR_ARG;
  if (f.w != w || f.h != h) { set_error("mjpg: the frame is %dx%d, not %dx%d", f.w, f.h, w, h); return B200_ERR_ARG; }
  if (!dc[0].ok && !ac[0].ok) {        // no DHT in the frame: the tables of T.81 Annex K
    dc[0].build(kStdDcLumCounts, kStdDcSymbols, 12); dc[1].build(kStdDcChrCounts, kStdDcSymbols, 12);
    ac[0].build(kStdAcLumCounts, kStdAcLumSymbols, 162); ac[1].build(kStdAcChrCounts, kStdAcChrSymbols, 162);
  }
  int mode;
  if (f.nc == 1) mode = 3;
  else if (f.c[1].hs != 1 || f.c[1].vs != 1 || f.c[2].hs != 1 || f.c[2].vs != 1) { set_error("mjpg: unsupported chroma sampling"); return B200_ERR_ARG; }
  else if (f.c[0].hs == 2 && f.c[0].vs == 2) mode = 0;
  else if (f.c[0].hs == 2 && f.c[0].vs == 1) mode = 1;
  else if (f.c[0].hs == 1 && f.c[0].vs == 1) mode = 2;
  else { set_error("mjpg: unsupported luma sampling %dx%d", f.c[0].hs, f.c[0].vs); return B200_ERR_ARG; }
  for (int i = 0; i < f.nc; i++)
    if (!f.have_q[f.c[i].tq] || !dc[f.c[i].td].ok || !ac[f.c[i].ta].ok) { set_error("mjpg: a table the scan names is missing"); return B200_ERR_ARG; }
  const size_t coef_bytes = f.blocks * 64 * sizeof(int16_t);
  if (!sc.ensure(coef_bytes, f.blocks * 64)) return B200_ERR_CUDA;
  // the previous frame's upload from the page-locked buffers must be over before they are rewritten
  B200_CHECK(cudaEventSynchronize(sc.done), "mjpg sync");
  if (!read_scan(jpeg, jpeg_bytes, f, dc, ac, sc.h_coef)) return B200_ERR_ARG;
  memcpy(sc.h_q, f.q, sizeof(f.q));
  IdctGeom g{};
  for (int i = 0; i < f.nc; i++) {
    g.first_block[i] = (uint32_t)f.c[i].first_block; g.bw[i] = f.c[i].bw; g.tq[i] = f.c[i].tq;
    g.plane_off[i] = (uint32_t)(f.c[i].first_block * 64);
  }
  g.first_block[f.nc] = (uint32_t)f.blocks;
  cudaStream_t s = (cudaStream_t)stream;
  B200_CHECK(cudaMemcpyAsync(sc.d_coef, sc.h_coef, coef_bytes, cudaMemcpyHostToDevice, s), "mjpg H2D coefficients");
  B200_CHECK(cudaMemcpyAsync(sc.d_q, sc.h_q, sizeof(f.q), cudaMemcpyHostToDevice, s), "mjpg H2D tables");
  B200_CHECK(cudaEventRecord(sc.done, s), "mjpg record");
  const int grid = (int)((f.blocks + kBlocksPerCta - 1) / kBlocksPerCta);
  k_mjpg_idct<<<grid, kIdctThreads, 0, s>>>(sc.d_coef, sc.d_q, g, f.nc, sc.d_planes);
  const int items = ((w + 3) >> 2) * h + 2 * (((w >> 1) + 3) >> 2) * (h >> 1);
  k_mjpg_to_i420<<<(items + 255) / 256, 256, 0, s>>>(sc.d_planes, g, mode, w, h, d_i420);
  count_launch(2);
  B200_CHECK(cudaGetLastError(), "mjpg launch");
  return B200_OK;
}
