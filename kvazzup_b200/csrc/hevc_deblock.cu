// In-loop deblocking filter (H.265 8.7.2) of the B200 HEVC codec (sm_100a); SURVEY.md 8a-K row K7.
//
// Two launches per picture: every vertical edge of the picture, then every horizontal edge on
// the vertically filtered samples -- the order the standard prescribes.  One thread per 4-line
// edge segment: edges lie on the 8x8 grid, a segment reads 8 samples across the edge for 4
// lines and rewrites at most 3 on each side, so segments of one pass never overlap and the
// filter runs in place.  Edges are the transform unit edges on the 8x8 grid (every CU is one 2Nx2N PU, so
// prediction edges are CU edges, which transform edges include).
#include "hevc_device.cuh"
#include "hevc_kernels.h"

namespace b200 {

namespace {

constexpr int kThreads = 256;

// luma coded block flag of the transform block of unit u that touches segment `seg` of the edge; side 0:
// u lies before the edge (P), 1: after it (Q).  A unit split into four 4x4 luma blocks (tu_log2 == 2, a
// foreign stream) keeps one flag per block in cbf bits 4..7.
__device__ __forceinline__ int edge_cbf(const CuInfo &u, int dir, int side, int seg)
{
  if (u.tu_log2 != 2) return u.cbf & 1;
  const int b = dir == 0 ? (side ? 0 : 1) + 2 * seg : (side ? 0 : 2) + seg;
  return (u.cbf >> (4 + b)) & 1;
}

__device__ __forceinline__ int edge_bs(const FrameParams &fp, const CuInfo &p, const CuInfo &q, int dir, int seg)
{
  if (p.pred_mode == 1 || q.pred_mode == 1) return 2;
  if (edge_cbf(p, dir, 0, seg) || edge_cbf(q, dir, 1, seg)) return 1;
  // different reference PICTURES (8.7.2.4): compared by POC distance, two list entries may name one picture
  if (fp.n_refs > 1 && fp.ref_dist[p.ref_idx & 15] != fp.ref_dist[q.ref_idx & 15]) return 1;
  return (abs(p.mvx - q.mvx) >= 4 || abs(p.mvy - q.mvy) >= 4) ? 1 : 0;
}

// pix -> q0 of line 0; xs steps across the edge, ys along it
__device__ __forceinline__ void luma_segment(uint8_t *pix, int xs, int ys, int bs, int qp, int beta_off, int tc_off)
{
  const int beta = c_beta[clip3(0, 51, qp + 2 * beta_off)];
  const int tc = c_tc[clip3(0, 53, qp + 2 * (bs - 1) + 2 * tc_off)];
  int p[4][4], q[4][4];
#pragma unroll
  for (int l = 0; l < 4; l++)
#pragma unroll
    for (int i = 0; i < 4; i++) { p[l][i] = pix[-(i + 1) * xs + l * ys]; q[l][i] = pix[i * xs + l * ys]; }
  int dp0 = abs(p[0][2] - 2 * p[0][1] + p[0][0]), dp3 = abs(p[3][2] - 2 * p[3][1] + p[3][0]);
  int dq0 = abs(q[0][2] - 2 * q[0][1] + q[0][0]), dq3 = abs(q[3][2] - 2 * q[3][1] + q[3][0]);
  int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
  if (dpq0 + dpq3 >= beta) return;
  bool s0 = 2 * dpq0 < (beta >> 2) && abs(p[0][3] - p[0][0]) + abs(q[0][0] - q[0][3]) < (beta >> 3) &&
            abs(p[0][0] - q[0][0]) < ((5 * tc + 1) >> 1);
  bool s3 = 2 * dpq3 < (beta >> 2) && abs(p[3][3] - p[3][0]) + abs(q[3][0] - q[3][3]) < (beta >> 3) &&
            abs(p[3][0] - q[3][0]) < ((5 * tc + 1) >> 1);
  bool dEp = dp < ((beta + (beta >> 1)) >> 3), dEq = dq < ((beta + (beta >> 1)) >> 3);
#pragma unroll
  for (int l = 0; l < 4; l++) {
    int p0 = p[l][0], p1 = p[l][1], p2 = p[l][2], p3 = p[l][3];
    int q0 = q[l][0], q1 = q[l][1], q2 = q[l][2], q3 = q[l][3];
    uint8_t *c = pix + l * ys;
    if (s0 && s3) {
      c[-xs] = (uint8_t)clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      c[-2 * xs] = (uint8_t)clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
      c[-3 * xs] = (uint8_t)clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
      c[0] = (uint8_t)clip3(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      c[xs] = (uint8_t)clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
      c[2 * xs] = (uint8_t)clip3(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
    } else {
      int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
      if (abs(delta) < tc * 10) {
        delta = clip3(-tc, tc, delta);
        c[-xs] = (uint8_t)clip8(p0 + delta);
        c[0] = (uint8_t)clip8(q0 - delta);
        if (dEp) c[-2 * xs] = (uint8_t)clip8(p1 + clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
        if (dEq) c[xs] = (uint8_t)clip8(q1 + clip3(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
      }
    }
  }
}

__device__ __forceinline__ void chroma_segment(uint8_t *pix, int xs, int ys, int qp_c, int tc_off)
{
  const int tc = c_tc[clip3(0, 53, qp_c + 2 + 2 * tc_off)];
#pragma unroll
  for (int l = 0; l < 4; l++) {
    uint8_t *c = pix + l * ys;
    int p0 = c[-xs], p1 = c[-2 * xs], q0 = c[0], q1 = c[xs];
    int delta = clip3(-tc, tc, ((((q0 - p0) << 2) + p1 - q1 + 4) >> 3));
    c[-xs] = (uint8_t)clip8(p0 + delta);
    c[0] = (uint8_t)clip8(q0 - delta);
  }
}

// dir 0: vertical edges, dir 1: horizontal edges.  Thread <-> (unit, segment).
__global__ void __launch_bounds__(kThreads)
k_deblock(FrameParams fp, uint8_t *rec, const CuInfo *__restrict__ cu, int dir)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int u = i >> 1, seg = i & 1;
  if (u >= fp.w8 * fp.h8) return;
  int x8 = u % fp.w8, y8 = u / fp.w8;
  CuInfo q = cu[u];
  int n8 = q.tu_log2 > 3 ? 1 << (q.tu_log2 - 3) : 1;          // transform unit edges (they include the CU edges)
  if (dir == 0 ? (x8 == 0 || (x8 & (n8 - 1))) : (y8 == 0 || (y8 & (n8 - 1)))) return;
  CuInfo p = cu[dir == 0 ? u - 1 : u - fp.w8];
  int bs = edge_bs(fp, p, q, dir, seg);
  if (!bs) return;
  int x = x8 * 8, y = y8 * 8;
  // QpL = (QpQ + QpP + 1) >> 1 (8.7.2.5.3); the two differ only across CTUs with different ROI offsets
  const int qp = fp.ctu_qp ? (p.qp + q.qp + 1) >> 1 : fp.qp;
  const int c_off = seg ? fp.cr_qp_offset_pps : fp.cb_qp_offset_pps;         // cQpPicOffset (8.7.2.5.5)
  const int qp_c = (fp.ctu_qp || c_off) ? c_chroma_qp_tab[clip3(0, 57, qp + c_off)] : fp.qp_c;
  if (dir == 0) luma_segment(rec + (size_t)(y + 4 * seg) * fp.w + x, 1, fp.w, bs, qp, fp.beta_offset_div2, fp.tc_offset_div2);
  else luma_segment(rec + (size_t)y * fp.w + x + 4 * seg, fp.w, 1, bs, qp, fp.beta_offset_div2, fp.tc_offset_div2);
  if (bs == 2 && ((dir == 0 ? x : y) & 15) == 0) {
    // chroma edges lie on the 8-sample chroma grid and are filtered for bS 2 only; seg 0 -> Cb, seg 1 -> Cr
    const size_t ysz = (size_t)fp.w * fp.h;
    const int cw = fp.w >> 1;
    uint8_t *plane = rec + ysz + (seg ? ysz / 4 : 0);
    uint8_t *pc = plane + (size_t)(y / 2) * cw + x / 2;
    if (dir == 0) chroma_segment(pc, 1, cw, qp_c, fp.tc_offset_div2);
    else chroma_segment(pc, cw, 1, qp_c, fp.tc_offset_div2);
  }
}

// Luma QP of every CU with one quantisation group per CTU (8.6.1, diff_cu_qp_delta_depth = 0); the
// encoder's counterpart of what the parser derives while decoding.  The left / above groups lie in
// other CTBs, so qPY_PRED is the slice QP at the start of a CTU row (WPP) and otherwise the QP of
// the last CU of the previous CTU.  Inside a CTU the CUs before the first coded residual keep the
// predicted QP; that CU codes CuQpDeltaVal and it and all later CUs have the CTU's target QP.
// One CTA per CTU row: warps find the first coded CU of their CTUs, thread 0 runs the (short)
// prediction chain along the row, then the warps write the per-CU QPs.
constexpr int kMaxCtbCols = 128;
__global__ void __launch_bounds__(256)
k_cu_qps(FrameParams fp, CuInfo *cu)
{
  __shared__ uint8_t s_first[kMaxCtbCols], s_pred[kMaxCtbCols];
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int col = warp; col < fp.ctb_cols; col += 8) {
    unsigned first = 64;
    for (int h = 0; h < 2; h++) {
      const int z = 32 * h + lane;
      const int x8 = col * 8 + z_to_x(z), y8 = row * 8 + z_to_y(z);
      const bool coded = x8 < fp.w8 && y8 < fp.h8 && cu[(size_t)y8 * fp.w8 + x8].cbf != 0;
      const unsigned m = __ballot_sync(0xffffffffu, coded);
      if (m && first == 64) first = 32 * h + __ffs(m) - 1;
    }
    if (lane == 0) s_first[col] = (uint8_t)first;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int pred = fp.qp;
    for (int col = 0; col < fp.ctb_cols; col++) {
      const int ctu = row * fp.ctb_cols + col, target = fp.ctu_qp[ctu];
      const bool coded = s_first[col] < 64;
      s_pred[col] = (uint8_t)pred;
      fp.ctu_first[ctu] = s_first[col];
      fp.ctu_delta[ctu] = (int8_t)(coded ? ((target - pred + 26 + 52) % 52) - 26 : 0);
      if (coded) pred = target;
    }
  }
  __syncthreads();
  for (int col = warp; col < fp.ctb_cols; col += 8) {
    const int target = fp.ctu_qp[row * fp.ctb_cols + col];
    for (int h = 0; h < 2; h++) {
      const int z = 32 * h + lane;
      const int x8 = col * 8 + z_to_x(z), y8 = row * 8 + z_to_y(z);
      if (x8 < fp.w8 && y8 < fp.h8) cu[(size_t)y8 * fp.w8 + x8].qp = (uint8_t)(z < s_first[col] ? s_pred[col] : target);
    }
  }
}

}  // namespace

cudaError_t launch_cu_qps(const FrameParams &fp, CuInfo *cu, cudaStream_t s)
{
  if (!fp.ctu_qp || fp.ctb_cols > kMaxCtbCols) return cudaErrorInvalidValue;
  k_cu_qps<<<fp.ctb_rows, 256, 0, s>>>(fp, cu);
  return cudaGetLastError();
}

cudaError_t launch_deblock(const FrameParams &fp, uint8_t *rec, const CuInfo *cu, cudaStream_t s)
{
  int threads = fp.w8 * fp.h8 * 2;
  int grid = (threads + kThreads - 1) / kThreads;
  k_deblock<<<grid, kThreads, 0, s>>>(fp, rec, cu, 0);
  k_deblock<<<grid, kThreads, 0, s>>>(fp, rec, cu, 1);
  return cudaGetLastError();
}

}  // namespace b200
