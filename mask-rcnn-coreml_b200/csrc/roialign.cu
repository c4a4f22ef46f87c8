// roialign.cu -- PyramidROIAlignLayer.evaluate (PyramidROIAlignLayer.swift:79-181).
//
//   roi_level_kernel        roisToInputItems (:351-396): FPN level per roi in fp64, round half away from zero, clamp
//                           2..5; NaN / inf -> padding.  Also zeroes the ticket counter of the staged kernel.
//   roi_order_kernel        processing order of the staged kernel for the boundary layout: per image by (level, y, x)
//   roialign_staged_kernel  both layouts -- boundary (the reference's): maps CHW fp32, output (R,C,P,P) fp32 (copyOutput
//                           :245-274 order); internal (fused pipeline): maps NHWC fp16, output (R,P,P,C) fp16 (the
//                           K-major operand of the head GEMMs).  Persistent, warp-specialised: a planner warp computes
//                           each roi's tap plan and publishes it as a descriptor in shared memory, an issuer warp
//                           stages every distinct feature ROW the roi touches through a chunk-allocated mbarrier ring
//                           (one cp.async.bulk.tensor per row), seven consumer warps compute the P x P samples from
//                           shared memory with packed fp32 arithmetic; rois are handed out by an atomic ticket.
//                           Rois the ring cannot serve (inverted boxes, footprints wider than 32 pixels) are gathered
//                           from global memory inside the same kernel.  DESIGN.md section 4 has the measurements.
//   roialign_chw_kernel, roialign_nhwc_kernel   pure gather kernels: what TMA boxes cannot cover (more than 256
//                           channels or not a multiple of 8, pool > 16, unaligned maps) and MRCNN_ROIALIGN=gather.
// Sampling = MPSNNCropAndResizeBilinear (:212-223) restated as TensorFlow crop_and_resize (bilinear, extrapolation 0) in
// fp32, every operation rounded individually, so the result is bit-identical to oracle/oracle.c on every path.
// Every output block is written (fixes Q5: the reference drops the last group).
#include "common.cuh"
#include "exact_math.cuh"
#include "tma_lite.cuh"
#include <limits.h>
#include <stdlib.h>
#include <string.h>

__global__ void roi_level_kernel(const float* __restrict__ rois, int roi_stride, int64_t total,
                                 double ratio, int32_t* __restrict__ level, int32_t* __restrict__ ticket) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *ticket = 0;               // work counter of the staged kernel that follows on the stream
  if (i >= total) return;
  const float* r = rois + i * roi_stride;
  double y1 = r[0], x1 = r[1], y2 = r[2], x2 = r[3];
  double w = __dsub_rn(x2, x1), h = __dsub_rn(y2, y1);
  double lf = __dadd_rn(log2(__ddiv_rn(sqrt(__dmul_rn(w, h)), ratio)), 4.0);   // :373
  int lv;
  if (isnan(lf) || isinf(lf)) lv = -1;                                          // :374 -> padding
  else {
    double rl = round(lf);                      // half away from zero (Q13)
    rl = rl < 2.0 ? 2.0 : (rl > 5.0 ? 5.0 : rl);
    lv = (int)rl;                                                               // :376
  }
  level[i] = lv;
}

// Processing order of the staged kernel: per image, rois sorted by (level, y centre, x centre), so that the ~300 rois
// in flight at any time read one band of one map and overlapping rois hit in L2 (the reference groups its rois by
// level for the same reason, PyramidROIAlignLayer.swift:432-466; the OUTPUT stays in roi order).  One CTA per image,
// bitonic sort of (key << 32 | roi) in shared memory; n = power of two >= R, at most 4096.
__global__ void __launch_bounds__(1024)
roi_order_kernel(const float* __restrict__ rois, int roi_stride, int R, int n, const int32_t* __restrict__ level,
                 int32_t* __restrict__ order) {
  extern __shared__ unsigned long long so_keys[];
  const int img = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < R) {
      const int64_t gi = (int64_t)img * R + i;
      const float* r = rois + gi * roi_stride;
      const int lv = level[gi];
      const float yc = 0.5f * (r[0] + r[2]), xc = 0.5f * (r[1] + r[3]);
      const unsigned yq = min(4095u, (unsigned)(fminf(fmaxf(yc, 0.0f), 1.0f) * 4096.0f));     // NaN -> 0
      const unsigned xq = min(4095u, (unsigned)(fminf(fmaxf(xc, 0.0f), 1.0f) * 4096.0f));
      const unsigned key = ((unsigned)(lv < 0 ? 7 : lv) << 24) | ((yq & 4095u) << 12) | (xq & 4095u);
      k = ((unsigned long long)key << 32) | (unsigned)i;
    }
    so_keys[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= n; size <<= 1)
    for (int st = size >> 1; st > 0; st >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int j = i ^ st;
        if (j > i) {
          const unsigned long long a = so_keys[i], b = so_keys[j];
          const bool up = (i & size) == 0;
          if ((a > b) == up) { so_keys[i] = b; so_keys[j] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < R; i += blockDim.x) order[(int64_t)img * R + i] = img * R + (int)(so_keys[i] & 0xffffffffu);
}

struct PyramidF32 { const float* p[4]; int h[4]; int w[4]; };
struct PyramidF16 { const __half* p[4]; int h[4]; int w[4]; };

struct SampleAxis { float lerp; int lo, hi; bool ok; };

// in = a1*(D-1) + idx * ((a2-a1)*(D-1)/(P-1))   (TF crop_and_resize, fp32)
__device__ __forceinline__ SampleAxis sample_axis(float a1, float a2, int D, int P, int idx) {
  const float dm1 = (float)(D - 1);
  float in;
  if (P > 1) {
    float scale = __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), dm1), (float)(P - 1));
    in = __fadd_rn(__fmul_rn(a1, dm1), __fmul_rn((float)idx, scale));
  } else {
    in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a1, a2)), dm1);
  }
  SampleAxis s;
  s.ok = !(in < 0.0f || in > dm1);
  float f = floorf(in), c = ceilf(in);
  s.lo = (int)f; s.hi = (int)c;
  s.lerp = __fsub_rn(in, f);
  return s;
}

__device__ __forceinline__ float bilerp(float tl, float tr, float bl, float br, float lx, float ly) {
  float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
  float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
  return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
}

// v0 (gather): grid (chunks of C*P*P outputs, R, batch). One thread per output
// element; consecutive threads walk (c, py, px) so writes are fully coalesced.
__global__ void __launch_bounds__(256)
roialign_chw_kernel(const float* __restrict__ rois, int roi_stride, int R, PyramidF32 pyr, int C, int P,
                    const int32_t* __restrict__ level, float* __restrict__ out) {
  const int img = blockIdx.z, r = blockIdx.y;
  const int64_t ri = (int64_t)img * R + r;
  const int PP = P * P;
  const int blk = C * PP;
  float* o = out + ri * blk;
  const int lv = level[ri];
  const int e0 = blockIdx.x * (blockDim.x * 4);
  if (lv < 0) {
    for (int k = 0; k < 4; ++k) { int e = e0 + k * blockDim.x + threadIdx.x; if (e < blk) o[e] = 0.0f; }
    return;
  }
  const int m = lv - 2;
  const int H = pyr.h[m], W = pyr.w[m];
  const float* fm = pyr.p[m] + (size_t)img * C * H * W;
  const float* rr = rois + ri * roi_stride;
  const float y1 = rr[0], x1 = rr[1], y2 = rr[2], x2 = rr[3];
  #pragma unroll
  for (int k = 0; k < 4; ++k) {
    int e = e0 + k * blockDim.x + threadIdx.x;
    if (e >= blk) break;
    int c = e / PP;
    int rem = e - c * PP;
    int py = rem / P, px = rem - py * P;
    SampleAxis sy = sample_axis(y1, y2, H, P, py);
    SampleAxis sx = sample_axis(x1, x2, W, P, px);
    float v = 0.0f;
    if (sy.ok && sx.ok) {
      const float* pl = fm + (size_t)c * H * W;
      float tl = __ldg(pl + (size_t)sy.lo * W + sx.lo), tr = __ldg(pl + (size_t)sy.lo * W + sx.hi);
      float bl = __ldg(pl + (size_t)sy.hi * W + sx.lo), br = __ldg(pl + (size_t)sy.hi * W + sx.hi);
      v = bilerp(tl, tr, bl, br, sx.lerp, sy.lerp);
    }
    o[e] = v;
  }
}

// NHWC fp16: one CTA per roi.  The 2P axis samples (tap rows / columns and lerp weights) are computed once per
// CTA into shared memory; then one warp per output sample, 8 channels (16 B) per lane per step, so every tap is a
// contiguous C*2-byte run.  Two samples are in flight per warp iteration (8 independent 16-byte loads per lane).
__device__ __forceinline__ uint32_t bilerp2(uint32_t a, uint32_t b, uint32_t c, uint32_t d, float lx, float ly) {
  const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a)), fb = __half22float2(*reinterpret_cast<const __half2*>(&b));
  const float2 fc = __half22float2(*reinterpret_cast<const __half2*>(&c)), fd = __half22float2(*reinterpret_cast<const __half2*>(&d));
  const __half2 r = __floats2half2_rn(bilerp(fa.x, fb.x, fc.x, fd.x, lx, ly), bilerp(fa.y, fb.y, fc.y, fd.y, lx, ly));
  return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ uint4 bilerp8(uint4 a, uint4 b, uint4 c, uint4 d, float lx, float ly) {
  return make_uint4(bilerp2(a.x, b.x, c.x, d.x, lx, ly), bilerp2(a.y, b.y, c.y, d.y, lx, ly),
                    bilerp2(a.z, b.z, c.z, d.z, lx, ly), bilerp2(a.w, b.w, c.w, d.w, lx, ly));
}

__global__ void __launch_bounds__(256, 3)
roialign_nhwc_kernel(const float* __restrict__ rois, int roi_stride, int R, PyramidF16 pyr, int C, int P,
                     const int32_t* __restrict__ level, __half* __restrict__ out) {
  __shared__ int s_lo[2][64], s_hi[2][64];        // [0] = y axis, [1] = x axis; lo = -1 marks an out-of-range sample
  __shared__ float s_lerp[2][64];
  const int img = blockIdx.y, r = blockIdx.x;
  const int64_t ri = (int64_t)img * R + r;
  const int PP = P * P;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __half* o = out + ri * (int64_t)PP * C;
  const int lv = level[ri];
  const int cvec = C >> 3;   // uint4 (8 halves) per pixel
  if (lv < 0) {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int e = threadIdx.x; e < PP * cvec; e += blockDim.x) reinterpret_cast<uint4*>(o)[e] = z;
    return;
  }
  const int m = lv - 2;
  const int H = pyr.h[m], W = pyr.w[m];
  const __half* fm = pyr.p[m] + (size_t)img * H * W * C;
  if (threadIdx.x < 128) {
    const int axis = threadIdx.x >> 6, i = threadIdx.x & 63;
    if (i < P) {
      const float* rr = rois + ri * roi_stride;
      const SampleAxis sa = axis == 0 ? sample_axis(rr[0], rr[2], H, P, i) : sample_axis(rr[1], rr[3], W, P, i);
      s_lo[axis][i] = sa.ok ? sa.lo : -1; s_hi[axis][i] = sa.hi; s_lerp[axis][i] = sa.lerp;
    }
  }
  __syncthreads();
  for (int s0 = wid; s0 < PP; s0 += 2 * nw) {
    const int s1 = s0 + nw;
    const bool has1 = s1 < PP;
    const int py0 = s0 / P, px0 = s0 - py0 * P;
    const int py1 = has1 ? s1 / P : py0, px1 = has1 ? s1 - py1 * P : px0;
    const int yl0 = s_lo[0][py0], yh0 = s_hi[0][py0], xl0 = s_lo[1][px0], xh0 = s_hi[1][px0];
    const int yl1 = s_lo[0][py1], yh1 = s_hi[0][py1], xl1 = s_lo[1][px1], xh1 = s_hi[1][px1];
    const bool ok0 = (yl0 >= 0) && (xl0 >= 0), ok1 = has1 && (yl1 >= 0) && (xl1 >= 0);
    for (int v = lane; v < cvec; v += 32) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      uint4 a0 = z, b0 = z, c0 = z, d0 = z, a1 = z, b1 = z, c1 = z, d1 = z;
      if (ok0) {
        a0 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl0 * W + xl0) * C) + v);
        b0 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl0 * W + xh0) * C) + v);
        c0 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh0 * W + xl0) * C) + v);
        d0 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh0 * W + xh0) * C) + v);
      }
      if (ok1) {
        a1 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl1 * W + xl1) * C) + v);
        b1 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl1 * W + xh1) * C) + v);
        c1 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh1 * W + xl1) * C) + v);
        d1 = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh1 * W + xh1) * C) + v);
      }
      // out-of-range samples are exactly zero (extrapolation value), in-range ones follow the oracle's op order
      reinterpret_cast<uint4*>(o + (size_t)s0 * C)[v] = ok0 ? bilerp8(a0, b0, c0, d0, s_lerp[1][px0], s_lerp[0][py0]) : z;
      if (has1)
        reinterpret_cast<uint4*>(o + (size_t)s1 * C)[v] = ok1 ? bilerp8(a1, b1, c1, d1, s_lerp[1][px1], s_lerp[0][py1]) : z;
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// TMA-staged NHWC kernel
// ----------------------------------------------------------------------------------------------------------------
// CTA = planner warp (warp 0) + CW consumer warps + row-issuing warp (the last one)
#define RA_NEG (-(1 << 30))

struct RoiTmaMaps { CUtensorMap m[4][8]; };       // [pyramid level][box width class k - 1: k chunks of cpx (4 or 8) pixels]

struct RoiTmaArgs {
  const float* rois; int roi_stride; int R; int total;      // total = batch * R
  int C, P;
  const int32_t* level;
  void* out;                       // NHWC: __half (R, P, P, C);  CHW: float (R, C, P, P)
  int pix;                         // bytes from one pixel of a staged row to the next: NHWC C * 2, CHW 4
  int slot_px;                     // widest footprint (pixels) the ring takes (wider rois go the gather way)
  int chunk_bytes;                 // allocation unit of the ring: 8 pixels = 8 * C * 2 bytes (multiple of 128)
  int nch;                         // chunks in the ring
  int* ticket;                     // work counter, zeroed by the level kernel: rois are handed out in order, one at a time
  const int32_t* order;            // processing order (roi_order_kernel) or nullptr = roi order
  int box_px[4][8];                // pixels a box of [level][class] really holds (min(cpx * (class + 1), W of the level))
  int cpx;                         // pixels per ring chunk: 4 (C % 16 == 0, so that a chunk is a multiple of 128 B) or 8
  float negzero;                   // -0.0f as a runtime value (see tl::mul2)
  PyramidF16 pyr;                  // NHWC maps (sizes are read from here in both layouts)
  PyramidF32 pyr32;                // CHW maps
};

// What one lane knows about a roi.  Lanes 0-15 hold y sample `lane`, lanes 16-31 x sample `lane - 16`.
struct LanePlan {
  int ok, lo, hi; float lerp;      // this lane's sample: in range?, floor / ceil tap index, lerp weight
  int pos_lo, pos_hi;              // y lanes: position of the lo / hi tap row in the roi's list of distinct rows
  int base, new_lo, new_hi;        // y lanes: rows this lane introduces, at list positions base, base + new_lo
  int nrows;                       // distinct tap rows of the roi (0: nothing to sample)
  int x0, nx;                      // leftmost tap column, columns spanned (0: no x sample in range)
  bool regular;                    // the ring can serve this roi
};

// Sample coordinates are non-decreasing in the sample index when a2 >= a1 (fp32 rounding is monotonic), so the taps of
// sample i satisfy lo(i) >= lo(i-1), hi(i) >= hi(i-1), lo(i) >= hi(i-1) - 1, and out-of-range samples (TF's extrapolation
// value, 0) can only sit at the two ends.  The distinct rows in ascending order are therefore found with one scan.
__device__ __forceinline__ LanePlan roi_lane_plan(float y1, float x1, float y2, float x2, int H, int W, int P, int lane, int slot_px, int xalign) {
  const unsigned full = 0xffffffffu;
  LanePlan p;
  const int axis = lane >> 4, i = lane & 15;
  p.ok = 0; p.lo = 0; p.hi = 0; p.lerp = 0.0f;
  if (i < P) {
    const SampleAxis sa = axis == 0 ? sample_axis(y1, y2, H, P, i) : sample_axis(x1, x2, W, P, i);
    p.ok = sa.ok ? 1 : 0; p.lo = sa.lo; p.hi = sa.hi; p.lerp = sa.lerp;
  }
  int prev_hi = __shfl_up_sync(full, p.ok ? p.hi : RA_NEG, 1, 16);
  if (i == 0) prev_hi = RA_NEG;
  p.new_lo = (p.ok && p.lo > prev_hi) ? 1 : 0;
  p.new_hi = (p.ok && p.hi > p.lo && p.hi > prev_hi) ? 1 : 0;
  const int cnt = p.new_lo + p.new_hi;
  int incl = cnt;
  #pragma unroll
  for (int d = 1; d < 16; d <<= 1) { const int t = __shfl_up_sync(full, incl, d, 16); if (i >= d) incl += t; }
  p.base = incl - cnt;
  p.pos_lo = (!p.ok || p.new_lo) ? p.base : p.base - 1 - (prev_hi - p.lo);
  p.pos_hi = !p.ok ? p.base : (p.hi == p.lo ? p.pos_lo : (p.hi > prev_hi ? p.base + p.new_lo : p.base - 1));
  p.nrows = __shfl_sync(full, incl, 15);                     // total of the y half
  int mn = p.ok ? p.lo : INT_MAX, mx = p.ok ? p.hi : RA_NEG;
  #pragma unroll
  for (int d = 8; d; d >>= 1) {
    mn = min(mn, __shfl_xor_sync(full, mn, d, 16));
    mx = max(mx, __shfl_xor_sync(full, mx, d, 16));
  }
  // CHW rows start at a multiple of `xalign` pixels: a TMA box whose innermost coordinate is not 16-byte aligned faults
  // (observed: "illegal instruction"); NHWC rows start at any pixel (the innermost dimension is the channel)
  p.x0 = __shfl_sync(full, mn, 16) & ~(xalign - 1);
  const int xe = __shfl_sync(full, mx, 16);
  p.nx = xe >= p.x0 ? xe - p.x0 + 1 : 0;
  p.regular = (y2 >= y1) && (x2 >= x1) && P <= 16 && p.nx <= slot_px;
  return p;
}

// packed version of bilerp2: same operations, same order, each individually rounded (see tl::mul2)
__device__ __forceinline__ uint32_t bilerp2p(uint32_t a, uint32_t b, uint32_t c, uint32_t d, float lx, float ly, float nz) {
  // only the left taps are converted; the right ones enter through the mixed-precision subtraction (r - l in one FHADD)
  const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a));
  const float2 fc = __half22float2(*reinterpret_cast<const __half2*>(&c));
  const float2 dt = make_float2(tl::sub_h_f((unsigned short)(b & 0xffffu), fa.x), tl::sub_h_f((unsigned short)(b >> 16), fa.y));
  const float2 db = make_float2(tl::sub_h_f((unsigned short)(d & 0xffffu), fc.x), tl::sub_h_f((unsigned short)(d >> 16), fc.y));
  const float2 top = tl::add2(fa, tl::mul2(dt, lx, nz));
  const float2 bot = tl::add2(fc, tl::mul2(db, lx, nz));
  const float2 o = tl::add2(top, tl::mul2(tl::sub2(bot, top), ly, nz));
  const __half2 r = __floats2half2_rn(o.x, o.y);
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint4 bilerp8p(uint4 a, uint4 b, uint4 c, uint4 d, float lx, float ly, float nz) {
  return make_uint4(bilerp2p(a.x, b.x, c.x, d.x, lx, ly, nz), bilerp2p(a.y, b.y, c.y, d.y, lx, ly, nz),
                    bilerp2p(a.z, b.z, c.z, d.z, lx, ly, nz), bilerp2p(a.w, b.w, c.w, d.w, lx, ly, nz));
}

// ---- roi descriptors: written by the planner warp (the only warp that computes the roi's plan), read by the consumer
// warps and the row issuer.  A ring of RA_DESCS descriptors lets the planner run up to RA_DESCS - 1 rois ahead.
#define RA_DESCS 4
#define RA_ROWS 32            // row entries (full / empty barrier pairs) of the ring
#define RA_D_HDR 0            // 32 B: kind, level index, image, distinct rows | y1, x1, y2, x2
#define RA_D_HDR2 32          // 16 B: row entry of the roi's first row, leading out-of-range sample rows, in-range sample rows
#define RA_D_XTAB 48          // 16 x 16 B per sample column: lo tap byte offset, hi tap byte offset, lerp, in range
#define RA_D_RTAB 304         // row-wise loop: 32 x 16 B per distinct row (ascending): ring address, full barrier, parity to wait for
#define RA_D_YTAB 816         //               16 x 8 B per sample row: lerp, lo row == hi row
#define RA_D_STAB 304         // sample-row loop (same bytes as RTAB / YTAB): 16 x 48 B per sample row:
                              //   lo row address, hi row address, lo full barrier, hi full barrier |
                              //   lo parity, hi parity, lerp, in range | empty barrier to arrive on (0: none) x 2
#define RA_D_ISSUE 1072       // 16 B for the issuing warp: rows to stage, leftmost column, chunks per row, row bytes
#define RA_D_ROWS 1088        // 32 x 4 B: the distinct tap rows (feature-map row indices)
#define RA_D_ETAB 1216       // 32 x 4 B per distinct row: sample rows COMPLETED by this row (their hi tap): first | count << 8
#define RA_D_EXT 1344         // 16 B: CHW layout: bytes from one channel of a staged row to the next (box pixels * 4)
#define RA_D_BYTES 1360
#define RA_KIND_ZERO 0        // padding roi (or nothing in range): the block is zeros
#define RA_KIND_RING 1        // rows staged through the ring
#define RA_KIND_GATHER 2      // the ring cannot hold this roi: taps straight from global memory
#define RA_KIND_RING_FAST 3   // RA_KIND_RING with every sample in range (and C == 256): the loop without predicates
#define RA_KIND_END 4         // no more rois: every role leaves its loop

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// x-lerp of one staged feature row at one sample column, for this lane's 8 channels:  H = l + (r - l) * lx -- the `top`
// (or `bot`) term of crop_and_resize.  It is evaluated once per (row, column): a row that is the bottom tap of one sample
// row and the top tap of the next is not interpolated twice, and the row's shared memory is free again right after.
struct HRow { float2 v[4]; };

__device__ __forceinline__ HRow hrow(uint32_t row_addr, uint32_t xl_off, uint32_t xh_off, float lx, float nz) {
  const uint4 l = tl::lds128(row_addr + xl_off), r = tl::lds128(row_addr + xh_off);
  const uint32_t lw[4] = {l.x, l.y, l.z, l.w}, rw[4] = {r.x, r.y, r.z, r.w};
  HRow h;
  #pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
    const float2 dt = make_float2(tl::sub_h_f((unsigned short)(rw[i] & 0xffffu), fl.x), tl::sub_h_f((unsigned short)(rw[i] >> 16), fl.y));
    h.v[i] = tl::add2(fl, tl::mul2(dt, lx, nz));
  }
  return h;
}

// out = top + (bot - top) * ly, rounded to fp16 (top == bot when the sample row sits exactly on a feature row)
__device__ __forceinline__ uint4 ylerp_pack(const HRow& top, const HRow& bot, float ly, float nz) {
  uint32_t o[4];
  #pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 r = tl::add2(top.v[i], tl::mul2(tl::sub2(bot.v[i], top.v[i]), ly, nz));
    const __half2 hh = __floats2half2_rn(r.x, r.y);
    o[i] = *reinterpret_cast<const uint32_t*>(&hh);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// PT = pool size known at compile time (7, 14) or 0 (any P <= 16); CTAS = CTAs per SM the launch bounds ask for
// (registers); CW = consumer warps.  Warp 0 = planner, warps 1..CW = consumers (consumer warp w owns sample columns
// w, w + CW, ...), warp CW + 1 = row issuer.  Planning (global reads of the roi, the tap plan, the descriptor) and issuing (waiting for ring space,
// the TMA loads) are separate warps because one warp doing both was the bottleneck: with the planner a few rois ahead
// the issuer is only ever blocked on ring space, i.e. the ring stays full.
// The ring is allocated in chunks of 8 pixels (a row takes 1-4 consecutive chunks, never wrapping: a row that would
// cross the end starts at chunk 0 and the tail is skipped) with one full / empty barrier pair per ROW (RA_ROWS row
// entries), so narrow rows do not occupy the space of wide ones and about twice as many rows are in flight as with
// fixed slots -- the kernel's speed is set by the bytes it keeps in flight.
// The loop over sample rows stays rolled on purpose: an unrolled variant (4096 instructions) was instruction-fetch
// bound (ncu: stall_no_inst on top), see DESIGN.md.
template <int PT, int CW, int CTAS, bool ROWWISE, bool CHW>
__global__ void __launch_bounds__((CW + 2) * 32, CTAS)
roialign_staged_kernel(const __grid_constant__ RoiTmaMaps maps, const RoiTmaArgs a) {
  constexpr int SLOTS = RA_ROWS;
  extern __shared__ uint8_t ra_smem[];
  __shared__ __align__(8) uint64_t s_full[SLOTS], s_empty[SLOTS], s_dfull[RA_DESCS], s_dempty[RA_DESCS];
  __shared__ uint8_t s_chunks_of[SLOTS];                      // producer: chunks held by the row in each row entry
  __shared__ __align__(16) uint8_t s_desc[RA_DESCS * RA_D_BYTES];
  __shared__ int g_lo[2][64], g_hi[2][64];                    // gather path: per-axis taps of the current roi
  __shared__ float g_lerp[2][64];
  constexpr int NPX = PT ? (PT + CW - 1) / CW : 3;     // sample columns per consumer warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t slots = (tl::smem_u32(ra_smem) + 127u) & ~127u;
  const uint32_t full0 = tl::smem_u32(s_full), empty0 = tl::smem_u32(s_empty);
  const uint32_t dfull0 = tl::smem_u32(s_dfull), dempty0 = tl::smem_u32(s_dempty), desc0 = tl::smem_u32(s_desc);
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) { tl::mbar_init(full0 + 8 * s, 1); tl::mbar_init(empty0 + 8 * s, CW); }
    for (int s = 0; s < RA_DESCS; ++s) { tl::mbar_init(dfull0 + 8 * s, 1); tl::mbar_init(dempty0 + 8 * s, CW + 1); }
    tl::fence_barrier_init();
  }
  __syncthreads();
  const int C = a.C, P = PT ? PT : a.P, PP = P * P, cvec = C >> 3;
  const uint32_t pix = (uint32_t)a.pix;

  if (warp == 0) {
    // ---------------- planner: roi -> plan -> descriptor (runs up to RA_DESCS - 1 rois ahead) ----------------
    uint32_t seq = 0;                                         // rows planned so far (row n uses row entry n % RA_ROWS)
    uint32_t n = 0;                                           // rois described so far
    uint32_t cursor = 0;                                      // ring chunk the next row starts at (same rule as the issuer)
    const uint32_t NCH = (uint32_t)a.nch;
    #pragma unroll 1
    for (;; ++n) {
      // next roi: a ticket (rois in order; CTAs that drew cheap rois simply draw more of them -- no tail of unlucky CTAs)
      // (the first roi of a CTA is its block index, tickets start after the grid: one atomic round trip less at start-up)
      int item = 0;
      if (lane == 0) item = n == 0 ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(a.ticket, 1);
      if (lane == 0 && item < a.total && a.order) item = __ldg(a.order + item);
      item = __shfl_sync(0xffffffffu, item, 0);
      const uint32_t d = desc0 + (n % RA_DESCS) * RA_D_BYTES;
      if (item >= a.total) {
        tl::mbar_wait(dempty0 + 8 * (n % RA_DESCS), ((n / RA_DESCS) & 1) ^ 1);
        if (lane == 0) { sts128(d + RA_D_HDR, RA_KIND_END, 0u, 0u, 0u); sts128(d + RA_D_ISSUE, 0u, 0u, 1u, 0u); tl::mbar_arrive(dfull0 + 8 * (n % RA_DESCS)); }
        break;
      }
      const float* rr = a.rois + (int64_t)item * a.roi_stride;
      const int lv = __ldg(a.level + item);
      const float y1 = __ldg(rr), x1 = __ldg(rr + 1), y2 = __ldg(rr + 2), x2 = __ldg(rr + 3);
      const int m = lv >= 0 ? lv - 2 : 0;
      LanePlan p = roi_lane_plan(y1, x1, y2, x2, a.pyr.h[m], a.pyr.w[m], P, lane, a.slot_px, CHW ? 4 : 1);
      const unsigned allok = __ballot_sync(0xffffffffu, (lane & 15) >= P || p.ok);
      int kind = lv < 0 ? RA_KIND_ZERO : (!p.regular ? RA_KIND_GATHER : ((p.nrows == 0 || p.nx == 0) ? RA_KIND_ZERO : RA_KIND_RING));
      const bool ring = kind == RA_KIND_RING;
      if (!ROWWISE && !CHW && ring && allok == 0xffffffffu && cvec == 32) kind = RA_KIND_RING_FAST;
      // chunks per row of this roi, and where row j of the roi lands: rows are placed one after the other from the
      // cursor; a row that would cross the end of the ring starts at chunk 0 instead
      const uint32_t k = ring ? (uint32_t)((p.nx + a.cpx - 1) / a.cpx) : 1u;
      const uint32_t r0 = (NCH - cursor) / k, per_lap = NCH / k;
      auto row_chunk = [&](uint32_t j) -> uint32_t { return j < r0 ? cursor + j * k : ((j - r0) % per_lap) * k; };
      const int img = item / a.R;
      tl::mbar_wait(dempty0 + 8 * (n % RA_DESCS), ((n / RA_DESCS) & 1) ^ 1);     // consumers + issuer are done with this descriptor
      const unsigned yvalid = __ballot_sync(0xffffffffu, lane < 16 && p.ok);
      if (lane == 0) {
        sts128(d + RA_D_HDR, (uint32_t)kind, (uint32_t)m, (uint32_t)img, ring ? (uint32_t)p.nrows : 0u);
        sts128(d + RA_D_HDR + 16, __float_as_uint(y1), __float_as_uint(x1), __float_as_uint(y2), __float_as_uint(x2));
        sts128(d + RA_D_HDR2, seq % SLOTS, yvalid ? (uint32_t)(__ffs(yvalid) - 1) : (uint32_t)P, (uint32_t)__popc(yvalid), (uint32_t)item);
        sts128(d + RA_D_ISSUE, ring ? (uint32_t)p.nrows : 0u, (uint32_t)p.x0, k, (uint32_t)a.box_px[m][k - 1] * (uint32_t)(a.chunk_bytes / a.cpx));
        if (CHW) tl::sts32(d + RA_D_EXT, (uint32_t)a.box_px[m][k - 1] * 4u);
      }
      if (ring) {
        const int i = lane & 15;
        if (ROWWISE) tl::sts32(d + RA_D_ETAB + 4 * lane, 0u);
        const int next_pos_lo = __shfl_down_sync(0xffffffffu, p.pos_lo, 1, 16);
        __syncwarp();
        if (lane >= 16) {
          if (i < P) sts128(d + RA_D_XTAB + 16 * i, (uint32_t)(p.lo - p.x0) * pix, (uint32_t)(p.hi - p.x0) * pix, __float_as_uint(p.lerp), (uint32_t)p.ok);
        } else {
          if (p.ok) {
            if (ROWWISE) {
              // this sample row is completed by (emitted after) row pos_hi; the sample rows completed by one row are consecutive
              const unsigned grp = __match_any_sync(yvalid, p.pos_hi);
              if (__ffs(grp) - 1 == lane) tl::sts32(d + RA_D_ETAB + 4 * p.pos_hi, (uint32_t)lane | ((uint32_t)__popc(grp) << 8));
              tl::sts64(d + RA_D_YTAB + 8 * i, __float_as_uint(p.lerp), p.pos_lo == p.pos_hi ? 1u : 0u);
            }
            #pragma unroll
            for (int w = 0; w < 2; ++w) {
              if (w == 0 ? p.new_lo : p.new_hi) {
                const uint32_t j = (uint32_t)(w == 0 ? p.base : p.base + p.new_lo), q = seq + j;
                if (ROWWISE) sts128(d + RA_D_RTAB + 16 * j, slots + row_chunk(j) * a.chunk_bytes, full0 + 8 * (q % SLOTS), (q / SLOTS) & 1u, 0u);
                tl::sts32(d + RA_D_ROWS + 4 * j, (uint32_t)(w == 0 ? p.lo : p.hi));
              }
            }
          }
          if (!ROWWISE && i < P) {
            // sample-row loop: after sample row i the rows before the next sample row's lo tap are done (at most two:
            // older ones went with earlier sample rows); out-of-range sample rows hold pos = rows so far
            const uint32_t qlo = seq + p.pos_lo, qhi = seq + p.pos_hi;
            const int rel0 = i == 0 ? 0 : p.pos_lo, rel1 = i + 1 < P ? next_pos_lo : p.nrows;
            const uint32_t ra = rel1 > rel0 ? empty0 + 8 * ((seq + rel0) % SLOTS) : 0u;
            const uint32_t rb = rel1 > rel0 + 1 ? empty0 + 8 * ((seq + rel0 + 1) % SLOTS) : 0u;
            sts128(d + RA_D_STAB + 48 * i, slots + row_chunk((uint32_t)p.pos_lo) * a.chunk_bytes,
                   slots + row_chunk((uint32_t)p.pos_hi) * a.chunk_bytes, full0 + 8 * (qlo % SLOTS), full0 + 8 * (qhi % SLOTS));
            sts128(d + RA_D_STAB + 48 * i + 16, (qlo / SLOTS) & 1u, (qhi / SLOTS) & 1u, __float_as_uint(p.lerp), (uint32_t)p.ok);
            tl::sts64(d + RA_D_STAB + 48 * i + 32, ra, rb);
          }
        }
      }
      __syncwarp();
      if (lane == 0) tl::mbar_arrive(dfull0 + 8 * (n % RA_DESCS));
      if (ring) {
        seq += p.nrows;
        const uint32_t last = row_chunk((uint32_t)p.nrows - 1) + k;
        cursor = last == NCH ? 0 : last;
      }
    }
    return;
  }

  if (warp == CW + 1) {
    // ---------------- row issuer (one lane): descriptor -> ring space -> TMA row loads ----------------
    if (lane == 0) {
      uint32_t seq = 0, n = 0, cursor = 0, free_chunks = (uint32_t)a.nch, head_row = 0;
      const uint32_t NCH = (uint32_t)a.nch;
      #pragma unroll 1
      for (;; ++n) {
        const uint32_t d = desc0 + (n % RA_DESCS) * RA_D_BYTES;
        tl::mbar_wait(dfull0 + 8 * (n % RA_DESCS), (n / RA_DESCS) & 1);
        const uint4 hdr = tl::lds128(d + RA_D_HDR), is = tl::lds128(d + RA_D_ISSUE);
        if (hdr.x == RA_KIND_END) break;
        const uint32_t nrows = is.x, k = is.z, bytes = is.w;
        const int x0 = (int)is.y, img = (int)hdr.z;
        const CUtensorMap* tm = &maps.m[hdr.y][k - 1];
        #pragma unroll 1
        for (uint32_t j = 0; j < nrows; ++j) {
          uint32_t c = cursor, need = k;
          if (c + k > NCH) { need += NCH - c; c = 0; }                       // skip the tail of the ring
          while (free_chunks < need || seq - head_row >= (uint32_t)SLOTS) {
            const uint32_t e = head_row % SLOTS;                              // oldest row in flight: wait until it is released
            tl::mbar_wait(empty0 + 8 * e, (head_row / SLOTS) & 1);
            free_chunks += s_chunks_of[e];
            ++head_row;
          }
          free_chunks -= need;
          cursor = c + k == NCH ? 0 : c + k;
          const uint32_t e = seq % SLOTS;
          s_chunks_of[e] = (uint8_t)need;
          const int row = (int)tl::lds32u(d + RA_D_ROWS + 4 * j);
          tl::mbar_expect_tx(full0 + 8 * e, bytes);
          if (CHW) tl::tma_load_4d(slots + c * a.chunk_bytes, tm, full0 + 8 * e, x0, row, 0, img);     // (W, H, C, N)
          else tl::tma_load_4d(slots + c * a.chunk_bytes, tm, full0 + 8 * e, 0, x0, row, img);          // (C, W, H, N)
          ++seq;
        }
        tl::mbar_arrive(dempty0 + 8 * (n % RA_DESCS));
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int cw = warp - 1, ct = threadIdx.x - 32, nct = CW * 32;
  const uint4 z = make_uint4(0, 0, 0, 0);
  const float nz = a.negzero;
  const bool lane_on = lane < cvec;                           // C < 256: the upper lanes have no channels
  uint32_t n = 0;
  #pragma unroll 1
  for (;; ++n) {
    const uint32_t d = desc0 + (n % RA_DESCS) * RA_D_BYTES;
    tl::mbar_wait(dfull0 + 8 * (n % RA_DESCS), (n / RA_DESCS) & 1);
    const uint4 hdr = tl::lds128(d + RA_D_HDR);
    const int kind = (int)hdr.x;
    if (kind == RA_KIND_END) break;
    const int item = (int)tl::lds32u(d + RA_D_HDR2 + 12);
    if (CHW) {
      // ---- boundary layout: staged rows are [channel][pixel] fp32, the output block is (C, P, P) fp32.  A lane owns one
      // (channel of a group of four, sample column) pair: lanes 8 cs + px; the warp walks the channel groups.
      float* ob = reinterpret_cast<float*>(a.out) + (int64_t)item * C * PP;
      if (kind == RA_KIND_RING) {
        const uint32_t pitch = tl::lds32u(d + RA_D_EXT);
        const int cs = lane >> 3, pxl = lane & 7;
        constexpr int NQ = PT ? (PT + 7) / 8 : 2;
        uint4 xq[NQ];
        #pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int px = pxl + 8 * q;
          xq[q] = px < P ? tl::lds128(d + RA_D_XTAB + 16 * px) : z;
        }
        uint32_t e = d + RA_D_STAB;
        #pragma unroll 1
        for (int py = 0; py < P; ++py, e += 48) {
          const uint4 e0 = tl::lds128(e), e1 = tl::lds128(e + 16);
          const uint2 e2 = tl::lds64(e + 32);
          if (e1.w) { tl::mbar_wait_nc(e0.z, e1.x); tl::mbar_wait_nc(e0.w, e1.y); }
          const float ly = __uint_as_float(e1.z);
          #pragma unroll 2
          for (int g = cw; 4 * g < C; g += CW) {
            const int c = 4 * g + cs;
            if (c < C) {
              const uint32_t lo = e0.x + (uint32_t)c * pitch, hi = e0.y + (uint32_t)c * pitch;
              #pragma unroll
              for (int q = 0; q < NQ; ++q) {
                const int px = pxl + 8 * q;
                if (px < P) {
                  float v = 0.0f;
                  if (e1.w && xq[q].w)
                    v = bilerp(tl::lds32(lo + xq[q].x), tl::lds32(lo + xq[q].y), tl::lds32(hi + xq[q].x), tl::lds32(hi + xq[q].y),
                               __uint_as_float(xq[q].z), ly);
                  // streaming store: the block is written once and not read here; keeping it out of L2 leaves the image's
                  // map resident (67 MB of fp32 P2 + 50 MB of output per image would otherwise thrash the 126 MB L2:
                  // ncu showed 2.9x the touched map bytes coming from DRAM)
                  __stcs(ob + ((size_t)c * P + py) * P + px, v);
                }
              }
            }
          }
          if (e2.x) {
            __syncwarp();
            if (lane == 0) { tl::mbar_arrive(e2.x); if (e2.y) tl::mbar_arrive(e2.y); }
          }
        }
      } else if (kind == RA_KIND_ZERO) {
        for (int el = ct; el < C * PP; el += nct) ob[el] = 0.0f;
      } else {
        // gather path: every element from global memory, like roialign_chw_kernel
        const uint4 box = tl::lds128(d + RA_D_HDR + 16);
        const int m = (int)hdr.y;
        const int H = a.pyr.h[m], W = a.pyr.w[m];
        const float* fm = a.pyr32.p[m] + (size_t)hdr.z * C * H * W;
        const float y1 = __uint_as_float(box.x), x1 = __uint_as_float(box.y), y2 = __uint_as_float(box.z), x2 = __uint_as_float(box.w);
        for (int el = ct; el < C * PP; el += nct) {
          const int c = el / PP, rem = el - c * PP, py = rem / P, px = rem - py * P;
          const SampleAxis sy = sample_axis(y1, y2, H, P, py), sx = sample_axis(x1, x2, W, P, px);
          float v = 0.0f;
          if (sy.ok && sx.ok) {
            const float* pl = fm + (size_t)c * H * W;
            v = bilerp(__ldg(pl + (size_t)sy.lo * W + sx.lo), __ldg(pl + (size_t)sy.lo * W + sx.hi),
                       __ldg(pl + (size_t)sy.hi * W + sx.lo), __ldg(pl + (size_t)sy.hi * W + sx.hi), sx.lerp, sy.lerp);
          }
          ob[el] = v;
        }
      }
      __syncwarp();
      if (lane == 0) tl::mbar_arrive(dempty0 + 8 * (n % RA_DESCS));
      continue;
    }
    __half* o = reinterpret_cast<__half*>(a.out) + (int64_t)item * PP * C;
    if (!ROWWISE && kind == RA_KIND_RING_FAST) {
      // every sample in range, all 32 lanes carry channels: no predicates in the loop
      uint4 xe[NPX];
      #pragma unroll
      for (int k = 0; k < NPX; ++k) {
        const int px = cw + k * CW;
        xe[k] = tl::lds128(d + RA_D_XTAB + 16 * (px < P ? px : 0));
        xe[k].x += (uint32_t)lane * 16u; xe[k].y += (uint32_t)lane * 16u;
      }
      uint4* od = reinterpret_cast<uint4*>(o) + (size_t)cw * 32 + lane;           // (py = 0, px = cw)
      uint32_t e = d + RA_D_STAB;
      if (NPX == 1) {
        // one sample column per warp: software-pipelined by hand (the waits and shared-memory loads are volatile asm, the
        // compiler keeps them in order) -- the taps of sample row py + 1 are fetched before the arithmetic of sample row
        // py.  Two register sets used alternately (loop unrolled by two): rotating one set costs 20 moves per sample.
        struct Stage { uint4 ta, tb, tc, td; float ly; uint2 rel; };
        auto fetch = [&](Stage& st, int py) {
          const uint32_t ee = e + 48u * (uint32_t)py;
          const uint4 e0 = tl::lds128(ee), e1 = tl::lds128(ee + 16);
          st.rel = tl::lds64(ee + 32);
          st.ly = __uint_as_float(e1.z);
          tl::mbar_wait_nc(e0.z, e1.x);
          tl::mbar_wait_nc(e0.w, e1.y);
          st.ta = tl::lds128(e0.x + xe[0].x); st.tb = tl::lds128(e0.x + xe[0].y);
          st.tc = tl::lds128(e0.y + xe[0].x); st.td = tl::lds128(e0.y + xe[0].y);
        };
        auto emit = [&](const Stage& st, int py) {
          od[(size_t)py * P * 32] = bilerp8p(st.ta, st.tb, st.tc, st.td, __uint_as_float(xe[0].z), st.ly, nz);
          if (st.rel.x) {                                     // hand back the rows no later sample row reads
            __syncwarp();
            if (lane == 0) { tl::mbar_arrive(st.rel.x); if (st.rel.y) tl::mbar_arrive(st.rel.y); }
          }
        };
        Stage sa, sb;
        fetch(sa, 0);
        int py = 0;
        #pragma unroll 1
        for (; py + 1 < P; py += 2) {
          fetch(sb, py + 1);
          emit(sa, py);
          if (py + 2 < P) fetch(sa, py + 2);
          emit(sb, py + 1);
        }
        if (py < P) emit(sa, py);
      } else {
        #pragma unroll 1
        for (int py = 0; py < P; ++py, e += 48, od += P * 32) {
          const uint4 e0 = tl::lds128(e), e1 = tl::lds128(e + 16);
          const uint2 e2 = tl::lds64(e + 32);
          tl::mbar_wait_nc(e0.z, e1.x);
          tl::mbar_wait_nc(e0.w, e1.y);
          #pragma unroll
          for (int k = 0; k < NPX; ++k) {
            if (cw + k * CW < P) {
              const uint4 ta = tl::lds128(e0.x + xe[k].x), tb = tl::lds128(e0.x + xe[k].y);
              const uint4 tc = tl::lds128(e0.y + xe[k].x), td = tl::lds128(e0.y + xe[k].y);
              od[k * CW * 32] = bilerp8p(ta, tb, tc, td, __uint_as_float(xe[k].z), __uint_as_float(e1.z), nz);
            }
          }
          if (e2.x) {                                           // hand back the rows no later sample row reads
            __syncwarp();
            if (lane == 0) { tl::mbar_arrive(e2.x); if (e2.y) tl::mbar_arrive(e2.y); }
          }
        }
      }
    } else if (!ROWWISE && kind == RA_KIND_RING) {
      // same loop with the predicates: samples outside the map (zeros), fewer than 256 channels
      uint4 xe[NPX];
      #pragma unroll
      for (int k = 0; k < NPX; ++k) {
        const int px = cw + k * CW;
        xe[k] = px < P ? tl::lds128(d + RA_D_XTAB + 16 * px) : z;
        xe[k].x += (uint32_t)lane * 16u; xe[k].y += (uint32_t)lane * 16u;
        if (!lane_on) xe[k].w = 0;
      }
      uint4* od = reinterpret_cast<uint4*>(o) + (size_t)cw * cvec + lane;
      uint32_t e = d + RA_D_STAB;
      #pragma unroll 1
      for (int py = 0; py < P; ++py, e += 48, od += (size_t)P * cvec) {
        const uint4 e0 = tl::lds128(e), e1 = tl::lds128(e + 16);
        const uint2 e2 = tl::lds64(e + 32);
        if (e1.w) { tl::mbar_wait_nc(e0.z, e1.x); tl::mbar_wait_nc(e0.w, e1.y); }
        #pragma unroll
        for (int k = 0; k < NPX; ++k) {
          if (cw + k * CW < P && lane_on) {
            uint4 r = z;
            if (e1.w && xe[k].w) {
              const uint4 ta = tl::lds128(e0.x + xe[k].x), tb = tl::lds128(e0.x + xe[k].y);
              const uint4 tc = tl::lds128(e0.y + xe[k].x), td = tl::lds128(e0.y + xe[k].y);
              r = bilerp8p(ta, tb, tc, td, __uint_as_float(xe[k].z), __uint_as_float(e1.z), nz);
            }
            od[(size_t)k * CW * cvec] = r;
          }
        }
        if (e2.x) {
          __syncwarp();
          if (lane == 0) { tl::mbar_arrive(e2.x); if (e2.y) tl::mbar_arrive(e2.y); }
        }
      }
    } else if (kind == RA_KIND_RING) {
      const uint4 h2 = tl::lds128(d + RA_D_HDR2);
      const int nrows = (int)hdr.w, lead = (int)h2.y, nvalid = (int)h2.z;
      uint4 xe[NPX];
      #pragma unroll
      for (int k = 0; k < NPX; ++k) {
        const int px = cw + k * CW;
        xe[k] = px < P ? tl::lds128(d + RA_D_XTAB + 16 * px) : z;
        xe[k].x += (uint32_t)lane * 16u; xe[k].y += (uint32_t)lane * 16u;
        if (!lane_on) xe[k].w = 0;
      }
      uint4* const od = reinterpret_cast<uint4*>(o) + (size_t)cw * cvec + lane;     // (py = 0, px = cw)
      // sample rows outside the map (only possible at the two ends): zeros
      for (int py = 0; py < P; ++py) {
        if (py == lead) { py += nvalid; if (py >= P) break; }
        #pragma unroll
        for (int k = 0; k < NPX; ++k)
          if (cw + k * CW < P && lane_on) od[(size_t)py * P * cvec + (size_t)k * CW * cvec] = z;
      }
      HRow prev[NPX], cur[NPX];
      #pragma unroll
      for (int k = 0; k < NPX; ++k)
        #pragma unroll
        for (int i = 0; i < 4; ++i) { prev[k].v[i] = make_float2(0.f, 0.f); cur[k].v[i] = make_float2(0.f, 0.f); }
      #pragma unroll 1
      for (int j = 0; j < nrows; ++j) {
        const uint4 r = tl::lds128(d + RA_D_RTAB + 16 * j);
        const uint32_t emit = tl::lds32u(d + RA_D_ETAB + 4 * j);
        tl::mbar_wait_nc(r.y, r.z);
        #pragma unroll
        for (int k = 0; k < NPX; ++k)
          if (xe[k].w) cur[k] = hrow(r.x, xe[k].x, xe[k].y, __uint_as_float(xe[k].z), nz);
        // the row is in registers: hand its ring space back at once
        __syncwarp();
        if (lane == 0) tl::mbar_arrive(empty0 + 8 * ((h2.x + (uint32_t)j) % SLOTS));
        const int cnt = (int)(emit >> 8), py0 = (int)(emit & 255u);
        for (int t = 0; t < cnt; ++t) {
          const uint2 ye = tl::lds64(d + RA_D_YTAB + 8 * (py0 + t));
          const float ly = __uint_as_float(ye.x);
          #pragma unroll
          for (int k = 0; k < NPX; ++k) {
            if (cw + k * CW < P && lane_on) {
              uint4 v = z;
              if (xe[k].w) {
                HRow top;
                #pragma unroll
                for (int i = 0; i < 4; ++i) top.v[i] = ye.y ? cur[k].v[i] : prev[k].v[i];
                v = ylerp_pack(top, cur[k], ly, nz);
              }
              od[(size_t)(py0 + t) * P * cvec + (size_t)k * CW * cvec] = v;
            }
          }
        }
        #pragma unroll
        for (int k = 0; k < NPX; ++k) prev[k] = cur[k];
      }
    } else if (kind == RA_KIND_ZERO) {
      for (int e = ct; e < PP * cvec; e += nct) reinterpret_cast<uint4*>(o)[e] = z;
    } else {
      // gather path: taps straight from global memory, like roialign_nhwc_kernel
      const uint4 box = tl::lds128(d + RA_D_HDR + 16);
      const int m = (int)hdr.y;
      const int H = a.pyr.h[m], W = a.pyr.w[m];
      const __half* fm = a.pyr.p[m] + (size_t)hdr.z * H * W * C;
      tl::named_bar_sync(1, nct);
      if (ct < 128) {
        const int axis = ct >> 6, i = ct & 63;
        if (i < P) {
          const SampleAxis sa = axis == 0 ? sample_axis(__uint_as_float(box.x), __uint_as_float(box.z), H, P, i)
                                          : sample_axis(__uint_as_float(box.y), __uint_as_float(box.w), W, P, i);
          g_lo[axis][i] = sa.ok ? sa.lo : -1; g_hi[axis][i] = sa.hi; g_lerp[axis][i] = sa.lerp;
        }
      }
      tl::named_bar_sync(1, nct);
      for (int s = cw; s < PP; s += CW) {
        const int py = s / P, px = s - py * P;
        const int yl = g_lo[0][py], yh = g_hi[0][py], xl = g_lo[1][px], xh = g_hi[1][px];
        const bool ok = yl >= 0 && xl >= 0;
        for (int v = lane; v < cvec; v += 32) {
          uint4 r = z;
          if (ok) {
            const uint4 ta = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl * W + xl) * C) + v);
            const uint4 tb = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yl * W + xh) * C) + v);
            const uint4 tc = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh * W + xl) * C) + v);
            const uint4 td = __ldg(reinterpret_cast<const uint4*>(fm + ((size_t)yh * W + xh) * C) + v);
            r = bilerp8p(ta, tb, tc, td, g_lerp[1][px], g_lerp[0][py], nz);
          }
          reinterpret_cast<uint4*>(o + (size_t)s * C)[v] = r;
        }
      }
    }
    __syncwarp();
    if (lane == 0) tl::mbar_arrive(dempty0 + 8 * (n % RA_DESCS));
  }
}

static int roi_levels(mrcnn_ctx* ctx, int batch, const float* d_rois, int roi_stride, int64_t R,
                      int32_t** d_level) {
  int64_t total = (int64_t)batch * R;
  if (total > ctx->roi_cap) {
    MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_roi_level);
    MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_roi_level, sizeof(int32_t) * (2 * total + 1)));    // levels, ticket counter, order
    ctx->roi_cap = (int)total;
  }
  // PyramidROIAlignLayer.swift:357 ratio = factor / sqrt(W*H)  (Q15: configured size always used)
  double ratio = (double)ctx->cfg.fpn_selection_factor / sqrt((double)ctx->cfg.image_w * (double)ctx->cfg.image_h);
  ProfScope ps(ctx, PROF_GLUE, (double)total * (roi_stride * 4 + 4));
  roi_level_kernel<<<ceil_div(total, 256), 256, 0, ctx->stream>>>(d_rois, roi_stride, total, ratio, ctx->d_roi_level, ctx->d_roi_level + ctx->roi_cap);
  MRCNN_LAUNCH_CHECK(ctx);
  *d_level = ctx->d_roi_level;
  return MRCNN_OK;
}

// ---- host side of the TMA-staged kernel ------------------------------------------------------------------------
struct RoiTmaEntry {                 // tensor maps of one pyramid (they do not depend on the rois or the pool size)
  const void* p[4]; int hw[8]; int C, batch; bool chw;
  RoiTmaMaps maps; int box_px[4][8]; int cpx;
};
struct RoiTmaCache { std::vector<RoiTmaEntry> entries; int ctas = 0; int rowwise = -1; int slot_px = 0; int mode = -1; int sorted = -1; };

void roialign_release(mrcnn_ctx* ctx) {
  delete (RoiTmaCache*)ctx->roi_tma;
  ctx->roi_tma = nullptr;
}

static RoiTmaCache* roi_cache(mrcnn_ctx* ctx) {
  if (!ctx->roi_tma) ctx->roi_tma = new RoiTmaCache();
  RoiTmaCache* cache = (RoiTmaCache*)ctx->roi_tma;
  if (cache->mode < 0) {                                      // knobs for A/B measurements and tests, read once per context
    const char* e = getenv("MRCNN_ROIALIGN");                 // "gather": the pure gather kernels
    cache->mode = (e && !strcmp(e, "gather")) ? 0 : 1;
    const char* es = getenv("MRCNN_ROIALIGN_SLOT_PX");        // widest footprint (pixels) the ring takes: 8, 16, 24 or 32
    cache->slot_px = es ? atoi(es) : 32;
    if (cache->slot_px != 8 && cache->slot_px != 16 && cache->slot_px != 24 && cache->slot_px != 32) cache->slot_px = 32;
    const char* en = getenv("MRCNN_ROIALIGN_CTAS");           // 2 / 3 CTAs per SM (0 = per-kernel default: 3 for pool 7, 2 for pool 14)
    cache->ctas = en ? std::max(2, std::min(3, atoi(en))) : 0;
    const char* ew = getenv("MRCNN_ROIALIGN_ROWWISE");        // 0 / 1: force the sample-row / feature-row consumer loop
    cache->rowwise = ew ? (atoi(ew) != 0) : -1;
    const char* eo = getenv("MRCNN_ROIALIGN_SORT");           // 0 / 1: process rois in roi order / sorted by (level, y, x)
    cache->sorted = eo ? (atoi(eo) != 0) : -1;
  }
  return cache;
}

// roi processing order for the staged kernel (nullptr: roi order).  Needs the levels (roi_levels ran on the stream).
static int roi_order(mrcnn_ctx* ctx, RoiTmaCache* cache, int batch, const float* d_rois, int roi_stride, int64_t R,
                     bool chw, const int32_t** order) {
  *order = nullptr;
  const bool want = cache->sorted >= 0 ? cache->sorted != 0 : chw;    // MRCNN_ROIALIGN_SORT=0/1; default: boundary layout only
  if (!want || R < 2 || R > 4096) return MRCNN_OK;
  int n = 2;
  while (n < R) n <<= 1;
  int32_t* o = ctx->d_roi_level + ctx->roi_cap + 1;
  ProfScope ps(ctx, PROF_GLUE, (double)batch * R * (roi_stride * 4 + 8));
  roi_order_kernel<<<batch, std::min(n, 1024), (size_t)n * 8, ctx->stream>>>(d_rois, roi_stride, (int)R, n, ctx->d_roi_level, o);
  MRCNN_LAUNCH_CHECK(ctx);
  *order = o;
  return MRCNN_OK;
}

static int roi_tma_entry(mrcnn_ctx* ctx, RoiTmaCache* cache, int batch, const void* const d_fmaps[4],
                         const int32_t hw[8], int C, bool chw, const RoiTmaEntry** out) {
  for (const auto& e : cache->entries)
    if (e.C == C && e.batch == batch && e.chw == chw && !memcmp(e.hw, hw, sizeof(e.hw)) && e.p[0] == d_fmaps[0] && e.p[1] == d_fmaps[1] &&
        e.p[2] == d_fmaps[2] && e.p[3] == d_fmaps[3]) { *out = &e; return MRCNN_OK; }
  PFN_tmapEncodeTiled enc = tmap_encode_fn();
  if (!enc) return mrcnn_fail(ctx, MRCNN_ECUDA, "roialign: cuTensorMapEncodeTiled not available from the driver");
  RoiTmaEntry e;
  for (int l = 0; l < 4; ++l) e.p[l] = d_fmaps[l];
  memcpy(e.hw, hw, sizeof(e.hw)); e.C = C; e.batch = batch; e.chw = chw;
  e.cpx = (chw || C % 16 == 0) ? 4 : 8;
  for (int l = 0; l < 4; ++l) {
    const int H = hw[2 * l], W = hw[2 * l + 1];
    // NHWC: 4-D (C, W, H, N) view, box = all channels x box_px pixels x one row.
    // CHW:  4-D (W, H, C, N) view, box = box_px pixels x one row x all channels (lands as [channel][pixel]).
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch};
    cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
    if (chw) {
      dims[0] = (cuuint64_t)W; dims[1] = (cuuint64_t)H; dims[2] = (cuuint64_t)C;
      str[0] = (cuuint64_t)W * 4; str[1] = (cuuint64_t)W * 4 * H; str[2] = (cuuint64_t)W * 4 * H * C;
    }
    cuuint32_t es[4] = {1, 1, 1, 1};
    for (int wc = 0; wc < 8; ++wc) {
      const int px = std::min(e.cpx * (wc + 1), W);
      e.box_px[l][wc] = px;
      cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)px, 1, 1};
      if (chw) { box[0] = (cuuint32_t)px; box[1] = 1; box[2] = (cuuint32_t)C; }
      CUresult r = enc(&e.maps.m[l][wc], chw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)d_fmaps[l], dims, str, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       // CHW: a box is C separate runs of 16-128 bytes, one per channel plane: promoting each to 256 B
                       // tripled the DRAM reads (ncu: 1.5 GB read for 0.47 GB touched); NHWC rows are >= 2 KB contiguous
                       chw ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        char b[160];
        snprintf(b, sizeof(b), "roialign: cuTensorMapEncodeTiled failed with %d (level %d, C %d, W %d, H %d, box %d px)", (int)r, l + 2, C, W, H, px);
        return mrcnn_fail(ctx, MRCNN_ECUDA, b);
      }
    }
  }
  if (cache->entries.size() >= 8) cache->entries.erase(cache->entries.begin());
  cache->entries.push_back(e);
  *out = &cache->entries.back();
  return MRCNN_OK;
}

template <int PT, int CW, int CTAS, bool ROWWISE, bool CHW = false>
static int launch_roialign_tma_t(mrcnn_ctx* ctx, const RoiTmaMaps& maps, RoiTmaArgs a) {
  auto kern = roialign_staged_kernel<PT, CW, CTAS, ROWWISE, CHW>;
  constexpr int threads = (CW + 2) * 32;
  // ring size: what CTAS resident CTAs leave of the SM's shared memory (228 KB per SM; per CTA: the kernel's static
  // part + 1 KB reserved by the system)
  static int static_smem = -1;
  if (static_smem < 0) {
    cudaFuncAttributes fa;
    MRCNN_CUDA_TRY(ctx, cudaFuncGetAttributes(&fa, kern));
    static_smem = (int)fa.sharedSizeBytes;
  }
  const int budget = (228 * 1024) / CTAS - 1024 - static_smem - 256;
  a.nch = std::min(budget / a.chunk_bytes, 255);
  // the software-pipelined loop holds the rows of two sample rows (4) at once (the others 2); with the tail skip of the
  // allocator a roi whose rows take k chunks needs 5k - 1 (3k - 1) chunks of ring to make progress: wider rois go the
  // gather way
  a.slot_px = std::min(a.slot_px, a.cpx * ((a.nch + 1) / (CHW ? 3 : 5)));
  if (a.slot_px < a.cpx) return mrcnn_fail(ctx, MRCNN_EINVAL, "roialign: too many channels for the staged kernel's ring");
  const int smem = a.nch * a.chunk_bytes + 128;
  static int per_sm_cached[64] = {0};
  static int smem_cached[64] = {0};
  const int dv = ctx->device & 63;
  if (smem_cached[dv] != smem) {
    MRCNN_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MRCNN_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_cached[dv], kern, threads, smem));
    smem_cached[dv] = smem;
  }
  const int per_sm = per_sm_cached[dv];
  if (per_sm < 1) return mrcnn_fail(ctx, MRCNN_ECUDA, "roialign: the staged kernel does not fit on an SM");
  const int grid = (int)std::min<int64_t>((int64_t)a.total, (int64_t)per_sm * ctx->sm_count);
  kern<<<grid, threads, smem, ctx->stream>>>(maps, a);
  return MRCNN_OK;
}

static int launch_roialign_tma(mrcnn_ctx* ctx, RoiTmaCache* cache, const RoiTmaMaps& maps, const RoiTmaArgs& a) {
  // Two consumer loops: by sample row (4 taps per sample; pool 7, where a feature row rarely serves two sample rows) and
  // by feature row (x-lerp once per row; pool 14, where it usually does).  MRCNN_ROIALIGN_ROWWISE / _CTAS force the other
  // variant (measurements, tests); the defaults are what measured fastest.
  const bool roww = cache->rowwise >= 0 ? cache->rowwise != 0 : a.P > 8;
  if (a.P == 7) {
    if (roww) return launch_roialign_tma_t<7, 7, 2, true>(ctx, maps, a);
      // default three CTAs per SM (21 consumer warps, 70 registers, rings of ~66 KB): equal at batch 8, +8 % at batch 64
    return cache->ctas == 2 ? launch_roialign_tma_t<7, 7, 2, false>(ctx, maps, a) : launch_roialign_tma_t<7, 7, 3, false>(ctx, maps, a);
  }
  if (a.P == 14) {
    if (!roww) return launch_roialign_tma_t<14, 7, 2, false>(ctx, maps, a);
    return cache->ctas == 3 ? launch_roialign_tma_t<14, 7, 3, true>(ctx, maps, a) : launch_roialign_tma_t<14, 7, 2, true>(ctx, maps, a);
  }
  return roww ? launch_roialign_tma_t<0, 7, 2, true>(ctx, maps, a) : launch_roialign_tma_t<0, 7, 2, false>(ctx, maps, a);
}

int roialign_nhwc_f16_run(mrcnn_ctx* ctx, int batch, const float* d_rois, int roi_stride, int64_t R,
                          const __half* const d_fmaps[4], const int32_t hw[8], int64_t C, int P,
                          __half* d_out, int32_t* d_level_out) {
  MRCNN_REQUIRE(ctx, batch >= 1 && R >= 1, "roialign: bad batch / num_rois");
  MRCNN_REQUIRE(ctx, (int64_t)batch * R <= INT_MAX, "roialign: batch * num_rois too large");
  MRCNN_REQUIRE(ctx, roi_stride >= 4, "roialign: roi_row_stride must be >= 4");
  MRCNN_REQUIRE(ctx, C >= 8 && (C % 8) == 0 && P >= 1 && P <= 64, "roialign(nhwc): channels must be a multiple of 8");
  for (int l = 0; l < 4; ++l)
    MRCNN_REQUIRE(ctx, d_fmaps[l] && hw[2 * l] >= 1 && hw[2 * l + 1] >= 1, "roialign(nhwc): null feature map / bad size");
  int32_t* lv = nullptr;
  int rc = roi_levels(ctx, batch, d_rois, roi_stride, R, &lv);
  if (rc) return rc;
  PyramidF16 pyr;
  for (int l = 0; l < 4; ++l) { pyr.p[l] = d_fmaps[l]; pyr.h[l] = hw[2 * l]; pyr.w[l] = hw[2 * l + 1]; }
  RoiTmaCache* cache = roi_cache(ctx);
  // bytes this launch has to move at least: the output once + the rois; the map bytes the rois really touch are
  // data dependent (bench.py reports them from the roi footprints and, under ncu, from dram__bytes)
  const double prof_bytes = (double)batch * (2.0 * R * C * P * P + 4.0 * R * roi_stride);
  if (cache->mode == 1 && C <= 256 && P <= 16 && (((uintptr_t)d_fmaps[0] | (uintptr_t)d_fmaps[1] | (uintptr_t)d_fmaps[2] | (uintptr_t)d_fmaps[3]) & 15) == 0) {
    const RoiTmaEntry* e = nullptr;
    rc = roi_tma_entry(ctx, cache, batch, (const void* const*)d_fmaps, hw, (int)C, false, &e);
    if (rc) return rc;
    RoiTmaArgs a;
    a.rois = d_rois; a.roi_stride = roi_stride; a.R = (int)R; a.total = (int)(batch * R);
    a.C = (int)C; a.P = P; a.level = lv; a.out = d_out;
    rc = roi_order(ctx, cache, batch, d_rois, roi_stride, R, false, &a.order);
    if (rc) return rc;
    a.pix = (int)C * 2; a.slot_px = cache->slot_px; a.cpx = e->cpx; a.chunk_bytes = e->cpx * (int)C * 2; a.nch = 0; a.ticket = ctx->d_roi_level + ctx->roi_cap;
    memcpy(a.box_px, e->box_px, sizeof(a.box_px));
    a.negzero = -0.0f; a.pyr = pyr; memset(&a.pyr32, 0, sizeof(a.pyr32));
    ProfScope ps(ctx, PROF_ROIALIGN, prof_bytes);
    rc = launch_roialign_tma(ctx, cache, e->maps, a);
    if (rc) return rc;
  } else {
    ProfScope ps(ctx, PROF_ROIALIGN, prof_bytes);
    dim3 grid((unsigned)R, batch);
    roialign_nhwc_kernel<<<grid, 256, 0, ctx->stream>>>(d_rois, roi_stride, (int)R, pyr, (int)C, P, lv, d_out);
  }
  MRCNN_LAUNCH_CHECK(ctx);
  if (d_level_out)
    MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(d_level_out, lv, sizeof(int32_t) * batch * R, cudaMemcpyDeviceToDevice, ctx->stream));
  return MRCNN_OK;
}

int roialign_chw_run(mrcnn_ctx* ctx, int batch, const float* d_rois, int roi_stride, int64_t R,
                     const float* const d_fmaps[4], const int32_t hw[8], int64_t C, int P,
                     float* d_out, int32_t* d_level_out) {
  MRCNN_REQUIRE(ctx, batch >= 1 && R >= 1 && R <= 65535, "roialign: bad batch / num_rois");
  MRCNN_REQUIRE(ctx, (int64_t)batch * R <= INT_MAX, "roialign: batch * num_rois too large");
  MRCNN_REQUIRE(ctx, roi_stride >= 4, "roialign: roi_row_stride must be >= 4");
  MRCNN_REQUIRE(ctx, C >= 1 && P >= 1 && P <= 64, "roialign: bad channels / pool");
  int32_t* lv = nullptr;
  int rc = roi_levels(ctx, batch, d_rois, roi_stride, R, &lv);
  if (rc) return rc;
  PyramidF32 pyr;
  bool tma_ok = C <= 256 && (C % 8) == 0 && P <= 16;     // pool > 16: every roi would take the in-kernel gather path
  for (int l = 0; l < 4; ++l) {
    pyr.p[l] = d_fmaps[l]; pyr.h[l] = hw[2 * l]; pyr.w[l] = hw[2 * l + 1];
    MRCNN_REQUIRE(ctx, pyr.h[l] >= 1 && pyr.w[l] >= 1, "roialign: bad feature map size");
    // TMA needs 16-byte aligned maps and row strides (W * 4 bytes)
    tma_ok = tma_ok && ((uintptr_t)d_fmaps[l] & 15) == 0 && (pyr.w[l] % 4) == 0;
  }
  const int blk = (int)C * P * P;
  RoiTmaCache* cache = roi_cache(ctx);
  // bytes this launch has to move at least: the output once + the rois (the touched map bytes are data dependent)
  const double prof_bytes = (double)batch * (4.0 * R * blk + 4.0 * R * roi_stride);
  if (cache->mode == 1 && tma_ok) {
    const RoiTmaEntry* e = nullptr;
    rc = roi_tma_entry(ctx, cache, batch, (const void* const*)d_fmaps, hw, (int)C, true, &e);
    if (rc) return rc;
    RoiTmaArgs a;
    a.rois = d_rois; a.roi_stride = roi_stride; a.R = (int)R; a.total = (int)(batch * R);
    a.C = (int)C; a.P = P; a.level = lv; a.out = d_out;
    rc = roi_order(ctx, cache, batch, d_rois, roi_stride, R, true, &a.order);
    if (rc) return rc;
    a.pix = 4; a.slot_px = cache->slot_px; a.cpx = e->cpx; a.chunk_bytes = e->cpx * 4 * (int)C; a.nch = 0;
    a.ticket = ctx->d_roi_level + ctx->roi_cap;
    memcpy(a.box_px, e->box_px, sizeof(a.box_px));
    a.negzero = -0.0f; a.pyr32 = pyr;
    for (int l = 0; l < 4; ++l) { a.pyr.p[l] = nullptr; a.pyr.h[l] = pyr.h[l]; a.pyr.w[l] = pyr.w[l]; }
    ProfScope ps(ctx, PROF_ROIALIGN, prof_bytes);
    if (P == 7) rc = launch_roialign_tma_t<7, 7, 2, false, true>(ctx, e->maps, a);
    else if (P == 14) rc = launch_roialign_tma_t<14, 7, 2, false, true>(ctx, e->maps, a);
    else rc = launch_roialign_tma_t<0, 7, 2, false, true>(ctx, e->maps, a);
    if (rc) return rc;
  } else {
    ProfScope ps(ctx, PROF_ROIALIGN, prof_bytes);
    dim3 grid(ceil_div(blk, 1024), (unsigned)R, batch);
    roialign_chw_kernel<<<grid, 256, 0, ctx->stream>>>(d_rois, roi_stride, (int)R, pyr, (int)C, P, lv, d_out);
  }
  MRCNN_LAUNCH_CHECK(ctx);
  if (d_level_out)
    MRCNN_CUDA_TRY(ctx, cudaMemcpyAsync(d_level_out, lv, sizeof(int32_t) * batch * R, cudaMemcpyDeviceToDevice, ctx->stream));
  return MRCNN_OK;
}
