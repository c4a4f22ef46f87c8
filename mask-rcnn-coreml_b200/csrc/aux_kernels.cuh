// aux_kernels.cuh -- the memory-bound glue kernels of the dense pipeline
// (everything that is not a convolution): image pre-processing, max-pool,
// P6 subsampling, RPN softmax/reorder, classifier softmax+argmax, the mask
// head's slot bookkeeping (its class-selected 1x1 + sigmoid is fused into the deconv epilogue, conv_gemm.cuh),
// and boundary-layout conversions.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// ---------------------------------------------------------------------------
// Image pre-processing (replaces Core ML's image bias, Conversion/task.py:73-75,
// and Keras ZeroPadding2D(3) in front of conv1): u8 RGB -> mean-subtracted fp16,
// zero padded by 3 and space-to-depth 2x2 so that the 7x7/2 stem convolution
// becomes a 4-tap, 64-channel implicit GEMM (see pipeline.cu "conv1").
//   out [B, Hs, Ws, 16] with channel (p*2+q)*3 + c = padded[2*ys+p][2*xs+q][c]
// ---------------------------------------------------------------------------
__global__ void preprocess_s2d_kernel(const uint8_t* __restrict__ rgb, int B, int H, int W, int Hs, int Ws,
                                      float m0, float m1, float m2, __half* __restrict__ out) {
  const int64_t total = (int64_t)B * Hs * Ws;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xs = (int)(i % Ws);
  const int ys = (int)((i / Ws) % Hs);
  const int b = (int)(i / ((int64_t)Ws * Hs));
  const float mean[3] = {m0, m1, m2};
  __align__(16) __half v[16];
  #pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = __float2half_rn(0.0f);
  #pragma unroll
  for (int p = 0; p < 2; ++p)
    #pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int Y = 2 * ys + p - 3, X = 2 * xs + q - 3;
      if (Y >= 0 && Y < H && X >= 0 && X < W) {
        const uint8_t* px = rgb + (((int64_t)b * H + Y) * W + X) * 3;
        #pragma unroll
        for (int c = 0; c < 3; ++c) v[(p * 2 + q) * 3 + c] = __float2half_rn((float)px[c] - mean[c]);
      }
    }
  uint4* o = reinterpret_cast<uint4*>(out + i * 16);
  o[0] = reinterpret_cast<const uint4*>(v)[0];
  o[1] = reinterpret_cast<const uint4*>(v)[1];
}

// 3x3 stride-2 max pool with TF "same" padding (pad at the end only): window rows
// 2y..2y+2 clipped to the map.  NHWC fp16, 8 channels per thread.
__global__ void maxpool3x3s2_kernel(const __half* __restrict__ in, int B, int H, int W, int C, int Ho, int Wo,
                                    __half* __restrict__ out) {
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % cv);
  const int x = (int)((i / cv) % Wo);
  const int y = (int)((i / ((int64_t)cv * Wo)) % Ho);
  const int b = (int)(i / ((int64_t)cv * Wo * Ho));
  __half2 m[4];
  bool first = true;
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = 2 * y + dy;
    if (yy >= H) break;
    for (int dx = 0; dx < 3; ++dx) {
      const int xx = 2 * x + dx;
      if (xx >= W) break;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((int64_t)b * H + yy) * W + xx) * C) + c8);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
      if (first) { for (int k = 0; k < 4; ++k) m[k] = h[k]; first = false; }
      else { for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], h[k]); }
    }
  }
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
  for (int k = 0; k < 4; ++k) oh[k] = m[k];
  reinterpret_cast<uint4*>(out + (((int64_t)b * Ho + y) * Wo + x) * C)[c8] = o;
}

// P6 = MaxPooling2D(pool 1, stride 2)(P5): plain subsampling.
__global__ void subsample2_kernel(const __half* __restrict__ in, int B, int H, int W, int C, int Ho, int Wo,
                                  __half* __restrict__ out) {
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % cv);
  const int x = (int)((i / cv) % Wo);
  const int y = (int)((i / ((int64_t)cv * Wo)) % Ho);
  const int b = (int)(i / ((int64_t)cv * Wo * Ho));
  reinterpret_cast<uint4*>(out + (((int64_t)b * Ho + y) * Wo + x) * C)[c8] =
      __ldg(reinterpret_cast<const uint4*>(in + (((int64_t)b * H + 2 * y) * W + 2 * x) * C) + c8);
}

// RPN head output of one level [B,h,w,24] f32 (ch 0..5 = class logits (anchor, bg/fg),
// ch 6..17 = deltas (anchor, 4)) -> probs [B,N,2] (softmax) and deltas [B,N,4] at
// anchor offset `off` (level-major, then y, x, anchor: the order of anchors.bin).
__global__ void rpn_post_kernel(const float* __restrict__ head, int B, int h, int w, int64_t N, int64_t off,
                                float* __restrict__ probs, float* __restrict__ deltas) {
  const int64_t total = (int64_t)B * h * w * 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int a = (int)(i % 3);
  const int64_t pix = i / 3;                       // b*h*w + y*w + x
  const int b = (int)(pix / ((int64_t)h * w));
  const int64_t loc = pix - (int64_t)b * h * w;
  const float* hp = head + pix * 24;
  const float l0 = hp[a * 2], l1 = hp[a * 2 + 1];
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  const float s = e0 + e1;
  const int64_t n = off + loc * 3 + a;
  float2* pp = reinterpret_cast<float2*>(probs + ((int64_t)b * N + n) * 2);
  *pp = make_float2(e0 / s, e1 / s);
  float4* dp = reinterpret_cast<float4*>(deltas + ((int64_t)b * N + n) * 4);
  *dp = make_float4(hp[6 + a * 4], hp[6 + a * 4 + 1], hp[6 + a * 4 + 2], hp[6 + a * 4 + 3]);
}

// Classifier head tail: logits [R, ld] f32 (0..ncls-1 class logits, ncls.. = bbox (ncls,4))
// -> softmax, first-max argmax (TimeDistributedClassifierLayer.swift:177-192), select.
// One warp per roi. Optionally also writes the full probabilities / boxes (layer ABI parity).
__global__ void cls_post_kernel(const float* __restrict__ logits, int64_t total, int ld, int ncls,
                                float* __restrict__ out6, float* __restrict__ probs_out, float* __restrict__ bbox_out) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= total) return;
  const float* l = logits + r * ld;
  float mx = -INFINITY;
  for (int c = lane; c < ncls; c += 32) mx = fmaxf(mx, l[c]);
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.0f;
  for (int c = lane; c < ncls; c += 32) sum += expf(l[c] - mx);
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int c = lane; c < ncls; c += 32) {
    const float p = expf(l[c] - mx) / sum;
    if (probs_out) probs_out[r * ncls + c] = p;
    if (p > bv) { bv = p; bi = c; }
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (bbox_out) for (int c = lane; c < ncls * 4; c += 32) bbox_out[r * ncls * 4 + c] = l[ncls + c];
  if (lane < 4) out6[r * 6 + lane] = l[ncls + bi * 4 + lane];
  if (lane == 4) out6[r * 6 + 4] = (float)bi;
  if (lane == 5) out6[r * 6 + 5] = bv;
}

// Slot bookkeeping of TimeDistributedMaskLayer (:52, :58-60, :71, :87-89):
//   valid blocks are compacted (index j), block a gets class detections[j*6+4] (Q14),
//   then rows [count(valid), D) are zeroed.  One CTA per image, ballot-based ordered compaction.
__global__ void __launch_bounds__(256)
mask_slots_kernel(const int32_t* __restrict__ block_valid, const float* __restrict__ det, int D,
                  int32_t* __restrict__ slot_valid, int32_t* __restrict__ slot_cls) {
  __shared__ int s_warp[8];
  __shared__ int s_base, s_total;
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int32_t* v = block_valid + (int64_t)img * D;
  const float* dd = det + (int64_t)img * D * 6;
  // pass 1: number of valid blocks
  int c = 0;
  for (int a = tid; a < D; a += 256) c += v[a] ? 1 : 0;
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) s_warp[wid] = c;
  if (tid == 0) s_base = 0;
  __syncthreads();
  if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_total = t; }
  __syncthreads();
  const int cnt = s_total;
  // pass 2: ordered compaction index j of every valid block (chunks of 256 slots)
  for (int start = 0; start < D; start += 256) {
    const int a = start + tid;
    const bool ok = a < D && v[a];
    const unsigned int bal = __ballot_sync(0xffffffffu, ok);
    __syncthreads();
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    const int j = before + __popc(bal & ((1u << lane) - 1u));
    if (a < D) {
      slot_cls[(int64_t)img * D + a] = ok ? (int)dd[j * 6 + 4] : 0;       // class of the j-th detection (Q14)
      slot_valid[(int64_t)img * D + a] = (ok && a < cnt) ? 1 : 0;          // rows [count(valid), D) are zeroed (:87-89)
    }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_base += t; }
  }
}

// (R, C, P, P) fp32 CHW blocks (layer-ABI layout) -> (R, P, P, C) fp16 NHWC, plus a
// per-block "has a non-zero element" flag (removeZeros, TimeDistributedClassifierLayer.swift:116-127;
// intended reading, Q9).  One CTA per block.
__global__ void chw_f32_to_nhwc_f16_kernel(const float* __restrict__ in, int C, int PP, __half* __restrict__ out,
                                           int32_t* __restrict__ nonzero) {
  const int64_t r = blockIdx.x;
  const float* src = in + r * (int64_t)C * PP;
  __half* dst = out + r * (int64_t)C * PP;
  int any = 0;
  for (int e = threadIdx.x; e < C * PP; e += blockDim.x) {
    const int c = e % C, p = e / C;           // write order (coalesced): p major, c minor
    const float v = src[(int64_t)c * PP + p];
    any |= (v != 0.0f);
    dst[e] = __float2half_rn(v);
  }
  any = __syncthreads_or(any);
  if (nonzero && threadIdx.x == 0) nonzero[r] = any;
}

__global__ void level_to_valid_kernel(const int32_t* __restrict__ level, int64_t n, int32_t* __restrict__ valid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) valid[i] = level[i] >= 0;
}
