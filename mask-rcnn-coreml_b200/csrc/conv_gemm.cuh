// conv_gemm.cuh -- the dense hot op: NHWC fp16 convolution as an implicit GEMM on
// the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by
// TMA), used for every dense layer of the pipeline (ResNet+FPN+RPN backbone,
// classifier head, mask head).
//
//   D[m, n] = sum_{tap, c}  X[pixel(m) + offset(tap), c] * W[n, tap, c]
//
//   M tile = a TH x TW rectangle of output pixels of one image (TH*TW = 128),
//            loaded per filter tap as ONE 4-D TMA box (64 ch, TW, TH, 1) at the
//            tap-shifted coordinate; out-of-image taps are zero-filled by TMA, so
//            padding costs nothing and no im2col buffer exists.  Stride-2 layers
//            use the tensor map's traversal strides.
//   N tile = BN output channels; weights are a (Cout, taps*Cin) K-major matrix
//            loaded as 2-D TMA boxes (64, BN).
//   K step = 64 channels of one tap (128 B rows, SWIZZLE_128B on both operands).
//
// Two launch shapes of the same kernel (template parameter CTAS):
//   CTAS = 1  one CTA per SM, 128 x BN tiles, tcgen05.mma.cta_group::1
//   CTAS = 2  thread-block clusters of two CTAs (one TPC): the pair computes a 256 x BN tile with
//             tcgen05.mma.cta_group::2 issued by the leader CTA; each CTA stages its own 128 rows of A and HALF of the
//             B tile (the tensor core reads both halves), which cuts the L2->smem operand traffic per flop by a third
//             and makes the stages smaller, so the ring is deeper (6 instead of 4 stages at BN = 256).
//
// Persistent CTAs (one per SM), 7 warps:
//   warp 0    TMA producer (one elected lane), NSTAGE-deep smem ring, mbarriers
//   warp 1    tcgen05.mma issuer (one elected lane) + TMEM allocator
//   warps 2-5 epilogue: tcgen05.ld (TMEM -> regs), + bias (+ residual, optionally
//             nearest-2x-upsampled = the FPN top-down add) (+ ReLU), fp16 or fp32
//             NHWC stores; optional 2x2 pixel-shuffle store (the mask head's
//             stride-2 transposed conv).  Two TMEM accumulator stages, so the
//             epilogue of tile i overlaps the MMAs of tile i+1.
//   warp 6    store thread of the staged epilogue (one elected lane): TMA stores of the
//             staged 64-channel chunks, buffer hand-back, residual prefetch
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CG_BM 128
#define CG_BK 64
#define CG_THREADS 224           // producer, MMA, 4 epilogue warps, store warp
#define CG_MAX_TAPS 49

struct ConvGemmParams {
  int n_img, h_out, w_out;
  int cout, ldc;               // valid output channels, output pixel stride (elements, multiple of 8)
  int cin;                     // input channels per tap (multiple of 64)
  int ntaps;
  int tw, th;                  // tile rectangle, tw*th == 128
  int tiles_x, tiles_y, tiles_n;
  int stride;                  // spatial stride of the conv
  int relu;
  int out_f32;                 // 1: fp32 output
  int res_mode;                // 0 none, 1 same resolution, 2 residual is half resolution (nearest 2x upsample)
  int res_h, res_w, res_ld;
  int deconv;                  // 1: cout = 4*deconv_c, output pixel (2y+dy, 2x+dx), channel c (2x2 stride-2 transposed conv)
  int deconv_c;
  int tma_out;                 // 1: fp16 NHWC output staged in smem and written with TMA stores (cout % 64 == 0)
  int tma_res;                 // 1: residual (res_mode 1) tiles fetched with TMA into the same staging buffers
  unsigned long long* trace;   // debug: per-CTA event trace (clock64 stamps), nullptr in production (tools/trace_conv.py)
  int ctas;                    // 1 or 2 (CTA pair / cta_group::2), must match the kernel instantiation and the cluster launch
  int nstages;                 // depth of the operand ring (A+B tiles), chosen per layer by the host
  int nbuf_log2;               // log2 of the number of 16 KB output / residual staging buffers (1, 2 or 3)
  int split_out;               // 1: outputs written as fp16 (hi, lo) pairs: channels [0,cout) = hi, [cout,2cout) = lo (2-term activations)
  int md_precise;              // maskdot: keep the activation in fp32 (no fp16 rounding before the class dot product)
  int maskdot;                 // 1: mask-head tail fused into the deconv epilogue (see epilogue_maskdot)
  const int32_t* md_valid;     // [n_img] slot is a real detection
  const int32_t* md_cls;       // [n_img] class id of the slot
  const __half* md_w;          // [md_ncls][deconv_c] final 1x1 weights
  const float* md_b;           // [md_ncls]
  int md_ncls;
  const float* bias;           // [cout] or nullptr
  const __half* residual;
  void* out;
  int8_t tap_dx[CG_MAX_TAPS];  // input offset of each tap, padding already subtracted
  int8_t tap_dy[CG_MAX_TAPS];
  // vertical tap groups (vgroup > 1): taps t*vgroup .. t*vgroup + vgroup-1 have the same dx and dy = dy0, dy0+1, ...; their A
  // tiles are `vgroup` overlapping windows of ONE (th + vgroup - 1)-row patch, loaded once (tmA's box has that height) and
  // addressed by the MMA at row offsets dy * tw: 3x3 convolutions load 3 patches of 10 rows instead of 9 tiles of 8 rows
  // per 64 input channels (2.4x less L2 -> SM traffic for A), the stem 1 patch of 11 rows instead of 4 tiles.
  int vgroup;                  // 1: one A tile per tap (the layout above)
  int na_stages;               // depth of the patch ring (vgroup > 1; the ring of `nstages` slots then holds B tiles only)
  int patch_bytes;             // bytes of one patch slot (multiple of 1024)
  int8_t tap_widx[CG_MAX_TAPS];   // vgroup > 1: position of tap t in the weight matrix's K order
};

namespace cg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "CG_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CG_DONE_%=;\n\t"
      "bra CG_WAIT_%=;\n\t"
      "CG_DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// ---- cluster / CTA-pair (cta_group::2) variants -----------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in this CTA's smem, the transaction bytes are signalled on the LEADER's barrier
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once the prior MMAs retire) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// packed fp32 add (FADD2) and the mixed-precision add float(h) + c (FHADD: exact conversion, one rounding)
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float add_h_f(unsigned short h, float c) {
  float d;
  asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(c));
  return d;
}

// debug event trace: per CTA three role sections (0 producer, 1 MMA, 2 epilogue) of CG_TRACE_MAX (code << 32 | arg,
// clock64) pairs, slot 0 of a section = its event count.  Plain stores with a per-thread cursor: nothing on the
// critical path waits for them.
#define CG_TRACE_MAX 340
struct TraceCursor { unsigned long long* base; uint32_t n; };
__device__ __forceinline__ TraceCursor trace_open(const ConvGemmParams& p, int role) {
  TraceCursor c; c.n = 0;
  c.base = p.trace ? p.trace + ((size_t)blockIdx.x * 3 + role) * (2 * CG_TRACE_MAX + 2) : nullptr;
  return c;
}
__device__ __forceinline__ void trace_ev(TraceCursor& c, uint32_t code, uint32_t arg) {
  if (c.base && c.n < CG_TRACE_MAX) {
    c.base[2 + 2 * c.n] = ((unsigned long long)code << 32) | arg;
    c.base[3 + 2 * c.n] = (unsigned long long)clock64();
    c.base[0] = ++c.n;
  }
}

// The work items of one CTA: item w = first, first + step, ... < total.  An item is a 128 x BN tile (CTAS = 1) or the
// 256 x BN tile of a CTA pair (CTAS = 2), of which this CTA owns the M tile 2 * (w / tiles_n) + crank.
struct TileSched {
  int first, step, total, ctas, crank;
  __device__ __forceinline__ int count() const { return first < total ? (total - 1 - first) / step + 1 : 0; }
  __device__ __forceinline__ void coords(const ConvGemmParams& p, int bn, int w, int& n0, int& x0, int& y0, int& img) const {
    const int nt = w % p.tiles_n;
    const int mt = (w / p.tiles_n) * ctas + crank;
    n0 = nt * bn;
    x0 = (mt % p.tiles_x) * p.tw;
    y0 = ((mt / p.tiles_x) % p.tiles_y) * p.th;
    img = mt / (p.tiles_x * p.tiles_y);       // == n_img for the phantom tile of an odd pair: TMA clips / zero-fills it
  }
};
// accumulator-drained signal: local arrive (CTAS = 1) or arrive on the leader CTA's barrier (CTAS = 2)
__device__ __forceinline__ void tempty_arrive(const TileSched& ts, uint32_t bar) {
  if (ts.ctas == 2) mbar_arrive_cluster(mapa_rank(bar, 0)); else mbar_arrive(bar);
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
//  [0,14) start address >> 4, [16,30) LBO >> 4 (ignored for swizzled K-major, 1),
//  [32,46) SBO >> 4 (8 rows * 128 B = 1024 B), [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: D = F32 (bits 4-5 = 1), A = B = F16 (0), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN, int CTAS> struct Cfg {
  static constexpr int kABytes = CG_BM * CG_BK * 2;          // 16 KB
  static constexpr int kBBytes = (BN / CTAS) * CG_BK * 2;    // a CTA of a pair stages half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  // shared memory = operand ring (nstages x kStageBytes) + output/residual staging (nbuf x 16 KB) + barriers + bias
  // tile; the split between ring depth and staging depth is chosen per layer (host: conv_plan_build):
  //   compute-bound layers   deep ring,  2 staging buffers
  //   memory-bound layers    short ring, 4 staging buffers (residual prefetched 3 chunks ahead, 3 stores in flight)
  static constexpr int kStagesDeep = CTAS == 2 ? ((BN >= 256) ? 6 : 8) : ((BN >= 256) ? 4 : ((BN >= 128) ? 6 : 8));
  static constexpr int kStagesShort = CTAS == 2 ? ((BN >= 256) ? 5 : ((BN >= 128) ? 6 : 8)) : ((BN >= 256) ? 3 : ((BN >= 128) ? 5 : 6));
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;   // BN in {32,64,128,256} -> power of two
  static constexpr int kOutStageBytes = CG_BM * 64 * 2;           // one 128-pixel x 64-channel fp16 sub-tile (SWIZZLE_128B)
  static constexpr int kMaxRingPlusOut = (kStagesDeep * kStageBytes + 2 * kOutStageBytes) > (kStagesShort * kStageBytes + 4 * kOutStageBytes)
                                             ? (kStagesDeep * kStageBytes + 2 * kOutStageBytes) : (kStagesShort * kStageBytes + 4 * kOutStageBytes);
  static constexpr int kBarBytes = 416;                           // mbarriers + TMEM base slot (layout in the kernel)
  static constexpr int kSmemBytes = kMaxRingPlusOut + 1024 /*align slack*/ + kBarBytes + BN * 4 /*bias tile*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
};

// ---- staged epilogue (fp16 NHWC outputs): TMEM -> regs -> (+bias, +residual, ReLU) -> fp16 -> swizzled smem ->
// TMA store, one 64-channel chunk at a time through 2 or 4 16-KB staging buffers.  RES = 0: no residual;
// 1: the residual sub-tile (same geometry as the output sub-tile) is TMA-loaded ahead into the very buffer the result
// is then written to; 2: FPN top-down add, residual gathered from the half-resolution map.
// Two roles.  The four epilogue warps (epilogue_staged) only compute: per chunk they wait for their buffer (res_full
// when a residual is fetched into it, store_free otherwise), convert, and hand the buffer over with one mbarrier
// arrival per warp (store_full).  The store thread (store_thread_staged, warp 6) issues the TMA stores, hands buffers
// back once a store has read them, and fetches the residual chunks ahead.  No block barrier and no serial section in
// the chunk loop: ncu on the res4 1x1 expansions showed every unit far from saturation (DRAM 45 %, L2 27 %, tensor
// 22 %) with the old scheme, where one epilogue thread issued store + wait + prefetch while the other 127 waited at
// the next barrier.
// Branch-free inner loop: flags are compile-time (RES) or folded into data (bias tile in smem, ReLU floor).
template <int BN, int RES>
__device__ __forceinline__ void epilogue_staged(const ConvGemmParams& p, uint8_t* out_gen, float* bias_gen,
                                                uint32_t rfull0, uint32_t sfull0, uint32_t sfree0, uint32_t tfull0, uint32_t tempty0,
                                                uint32_t tmem_base, const TileSched ts, int warp, int lane) {
  constexpr int kStageBytes = CG_BM * 64 * 2;
  const uint32_t nb_log2 = (uint32_t)p.nbuf_log2, nb_mask = (1u << nb_log2) - 1u;
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int nchunks = (min(BN, p.cout) + 63) / 64;
  const __half2 lo2 = __float2half2_rn(p.relu ? 0.0f : -INFINITY);
  uint32_t cc = 0;                         // chunk counter of this CTA (staging buffer = cc & nb_mask)
  int bias_n0 = -1;
  int acc = 0; uint32_t acc_phase = 0;
  const uint32_t row_off = (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);
  TraceCursor tc = trace_open(p, 2);
  if (warp != 2 || lane != 0) tc.base = nullptr;                    // one epilogue thread traces (debug only)
  for (int tile = ts.first; tile < ts.total; tile += ts.step) {
    int n0, x0, y0, img;
    ts.coords(p, BN, tile, n0, x0, y0, img);
    const __half* res_row = p.residual;    // RES == 2: always a readable address (rows outside the map are clipped by the store)
    if (RES == 2) {
      int x = x0 + (row % p.tw), y = y0 + (row / p.tw);
      x = min(x, p.w_out - 1); y = min(y, p.h_out - 1);
      const int im = min(img, p.n_img - 1);
      res_row = p.residual + (((size_t)im * p.res_h + (y >> 1)) * (size_t)p.res_w + (x >> 1)) * (size_t)p.res_ld;
    }
    if (n0 != bias_n0) {                   // stage this N tile's bias (zeros when there is none) in smem
      epi_bar_sync();
      for (int i = threadIdx.x - 64; i < BN; i += 128)
        bias_gen[i] = (p.bias && n0 + i < p.cout) ? __ldg(p.bias + n0 + i) : 0.0f;
      bias_n0 = n0;
      epi_bar_sync();
    }
    mbar_wait(tfull0 + 8u * acc, acc_phase);
    trace_ev(tc, 5, (uint32_t)tile);       // epilogue: accumulator ready
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
    #pragma unroll 1
    for (int c = 0; c < nchunks; ++c, ++cc) {
      const uint32_t buf = cc & nb_mask;
      const uint32_t use = cc >> nb_log2;
      uint8_t* srow = out_gen + buf * kStageBytes + row_off;
      const int nc = n0 + c * 64;
      if (RES == 1) mbar_wait(rfull0 + 8u * buf, use & 1u);          // residual chunk has landed (so the buffer is free, too)
      else mbar_wait(sfree0 + 8u * buf, (use & 1u) ^ 1u);            // the store that last used this buffer has read it
      trace_ev(tc, 9, (uint32_t)c);        // epilogue: staging buffer ready (residual landed / store has read it)
      #pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(c * 64 + hh * 32), v);
        // operands that do not depend on the accumulator are fetched while the TMEM load is in flight
        float4 bq[8];
        uint4 rq[4];
        #pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4* bp = reinterpret_cast<const float4*>(bias_gen + (c * 64 + hh * 32 + g * 8));
          bq[2 * g] = bp[0]; bq[2 * g + 1] = bp[1];
          if (RES == 1) rq[g] = *reinterpret_cast<const uint4*>(srow + ((((uint32_t)(hh * 4 + g)) ^ sw) << 4));
          if (RES == 2) rq[g] = __ldg(reinterpret_cast<const uint4*>(res_row + nc + hh * 32 + g * 8));
        }
        tmem_ld_wait();
        // (acc + bias) + residual in fp32, one rounding to fp16, ReLU on the packed halves (max commutes with the
        // monotonic rounding).  Packed fp32 adds for the bias, the mixed-precision add for the fp16 residual: 92
        // instead of 160 instructions per 32 columns -- one warp per scheduler runs this, every issue slot is time.
        #pragma unroll
        for (int g = 0; g < 4; ++g) {
          float2 f[4];
          f[0] = add2(make_float2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1])), make_float2(bq[2 * g].x, bq[2 * g].y));
          f[1] = add2(make_float2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3])), make_float2(bq[2 * g].z, bq[2 * g].w));
          f[2] = add2(make_float2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5])), make_float2(bq[2 * g + 1].x, bq[2 * g + 1].y));
          f[3] = add2(make_float2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])), make_float2(bq[2 * g + 1].z, bq[2 * g + 1].w));
          if (RES != 0) {
            const uint32_t rw[4] = {rq[g].x, rq[g].y, rq[g].z, rq[g].w};
            #pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[e].x = add_h_f((unsigned short)(rw[e] & 0xffffu), f[e].x);
              f[e].y = add_h_f((unsigned short)(rw[e] >> 16), f[e].y);
            }
          }
          uint4 ov;
          __half2* oh = reinterpret_cast<__half2*>(&ov);
          #pragma unroll
          for (int e = 0; e < 4; ++e) oh[e] = __hmax2(__floats2half2_rn(f[e].x, f[e].y), lo2);
          *reinterpret_cast<uint4*>(srow + ((((uint32_t)(hh * 4 + g)) ^ sw) << 4)) = ov;
        }
      }
      if (c == nchunks - 1) {              // accumulator fully read: hand the TMEM stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tempty_arrive(ts, tempty0 + 8u * acc);
      }
      fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the bulk-copy engine
      __syncwarp();
      if (lane == 0) mbar_arrive(sfull0 + 8u * buf);                 // 4 arrivals (one per epilogue warp) = chunk staged
      trace_ev(tc, 6, (uint32_t)c);        // epilogue: chunk handed to the store thread
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
  }
}

// The store thread of the staged epilogue (one lane of warp 6), see above.
template <int BN, int RES>
__device__ __forceinline__ void store_thread_staged(const ConvGemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmR,
                                                    uint32_t out_base, uint32_t rfull0, uint32_t sfull0, uint32_t sfree0,
                                                    const TileSched ts) {
  constexpr int kStageBytes = CG_BM * 64 * 2;
  const uint32_t nb_log2 = (uint32_t)p.nbuf_log2, nb_mask = (1u << nb_log2) - 1u;
  const int nchunks = (min(BN, p.cout) + 63) / 64;
  const uint32_t my_chunks = (uint32_t)(ts.count() * nchunks);
  // residual prefetch (RES == 1): chunk j of this CTA = work item first + (j / nchunks) * step, chunk j % nchunks
  uint32_t lr_tile = 0xffffffffu;
  int lr_n0 = 0, lr_x0 = 0, lr_y0 = 0, lr_img = 0;
  auto load_residual = [&](uint32_t j) {
    if (j >= my_chunks) return;
    const uint32_t ti = j / (uint32_t)nchunks;
    if (ti != lr_tile) {
      lr_tile = ti;
      ts.coords(p, BN, ts.first + (int)ti * ts.step, lr_n0, lr_x0, lr_y0, lr_img);
    }
    const uint32_t b = j & nb_mask;
    mbar_expect_tx(rfull0 + 8u * b, (uint32_t)kStageBytes);
    tma_load_4d(out_base + b * (uint32_t)kStageBytes, tmR, rfull0 + 8u * b, lr_n0 + (int)(j - ti * (uint32_t)nchunks) * 64, lr_x0, lr_y0, lr_img);
  };
  prefetch_tmap(tmC);
  if (RES == 1) { prefetch_tmap(tmR); for (uint32_t j = 0; j < nb_mask; ++j) load_residual(j); }     // prefetch distance = nbuf - 1 chunks
  uint32_t cc = 0, freed = 0;
  for (int tile = ts.first; tile < ts.total; tile += ts.step) {
    int n0, x0, y0, img;
    ts.coords(p, BN, tile, n0, x0, y0, img);
    for (int c = 0; c < nchunks; ++c, ++cc) {
      const uint32_t buf = cc & nb_mask;
      mbar_wait(sfull0 + 8u * buf, (cc >> nb_log2) & 1u);
      tma_store_4d(tmC, out_base + buf * (uint32_t)kStageBytes, n0 + c * 64, x0, y0, img);
      bulk_commit();
      if (RES == 1) {
        bulk_wait_read<1>();               // store cc-1 has finished reading its buffer ...
        load_residual(cc + nb_mask);       // ... which is the buffer of chunk cc + nbuf - 1: fetch that chunk's residual
      } else {
        // hand back every buffer whose store has been read: all but the newest (2 buffers) / the newest three (4 buffers)
        uint32_t done;
        if (nb_mask == 1u) { bulk_wait_read<1>(); done = cc; } else { bulk_wait_read<3>(); done = cc >= 2u ? cc - 2u : 0u; }
        while (freed < done) { mbar_arrive(sfree0 + 8u * (freed & nb_mask)); ++freed; }
      }
    }
  }
  bulk_wait<0>();
}

// ---- 2-term ("split") outputs for the precise mask head: every activation v is stored as hi = fp16(v) and
// lo = fp16(v - hi) in channels [n] and [cout + n] of a 2*cout-channel tensor.  The next layer sees 2*cout input
// channels with its weights duplicated, so hi*w + lo*w accumulates in fp32: activations carry ~22 mantissa bits
// through the tensor cores.  Same staging / TMA-store scheme as epilogue_staged (4 staging buffers: a hi/lo pair
// per chunk, two pairs in rotation); no residual.
template <int BN>
__device__ __forceinline__ void epilogue_split(const ConvGemmParams& p, const CUtensorMap* tmC, uint8_t* out_gen, uint32_t out_base,
                                               float* bias_gen, uint32_t tfull0, uint32_t tempty0, uint32_t tmem_base,
                                               const TileSched ts, int warp, int lane) {
  constexpr int kStageBytes = CG_BM * 64 * 2;
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const bool e0 = (threadIdx.x == 64);
  const int nchunks = (min(BN, p.cout) + 63) / 64;
  const float lo_clamp = p.relu ? 0.0f : -INFINITY;
  uint32_t cc = 0;
  int bias_n0 = -1;
  int acc = 0; uint32_t acc_phase = 0;
  if (e0) prefetch_tmap(tmC);
  const uint32_t row_off = (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);
  for (int tile = ts.first; tile < ts.total; tile += ts.step) {
    int n0, x0, y0, img;
    ts.coords(p, BN, tile, n0, x0, y0, img);
    if (n0 != bias_n0) {
      epi_bar_sync();
      for (int i = threadIdx.x - 64; i < BN; i += 128)
        bias_gen[i] = (p.bias && n0 + i < p.cout) ? __ldg(p.bias + n0 + i) : 0.0f;
      bias_n0 = n0;
      epi_bar_sync();
    }
    mbar_wait(tfull0 + 8u * acc, acc_phase);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
    #pragma unroll 1
    for (int c = 0; c < nchunks; ++c, ++cc) {
      const uint32_t pair = (cc & 1u) * 2u;                    // staging buffers {0,1} / {2,3}
      uint8_t* srow_hi = out_gen + pair * kStageBytes + row_off;
      uint8_t* srow_lo = srow_hi + kStageBytes;
      const int nc = n0 + c * 64;
      epi_bar_sync();                                           // e0 has seen the stores that last read this pair finish
      #pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(c * 64 + hh * 32), v);
        float4 bq[8];
        #pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4* bp = reinterpret_cast<const float4*>(bias_gen + (c * 64 + hh * 32 + g * 8));
          bq[2 * g] = bp[0]; bq[2 * g + 1] = bp[1];
        }
        tmem_ld_wait();
        #pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
          #pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g * 8 + e]);
          f[0] += bq[2 * g].x; f[1] += bq[2 * g].y; f[2] += bq[2 * g].z; f[3] += bq[2 * g].w;
          f[4] += bq[2 * g + 1].x; f[5] += bq[2 * g + 1].y; f[6] += bq[2 * g + 1].z; f[7] += bq[2 * g + 1].w;
          uint4 hv, lv;
          __half2* hh2 = reinterpret_cast<__half2*>(&hv);
          __half2* lh2 = reinterpret_cast<__half2*>(&lv);
          #pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = fmaxf(f[2 * e], lo_clamp), b = fmaxf(f[2 * e + 1], lo_clamp);
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            hh2[e] = h;
            lh2[e] = __floats2half2_rn(a - hf.x, b - hf.y);
          }
          const uint32_t off = (((uint32_t)(hh * 4 + g)) ^ sw) << 4;
          *reinterpret_cast<uint4*>(srow_hi + off) = hv;
          *reinterpret_cast<uint4*>(srow_lo + off) = lv;
        }
      }
      if (c == nchunks - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tempty_arrive(ts, tempty0 + 8u * acc);
      }
      fence_proxy_async_smem();
      epi_bar_sync();
      if (e0) {
        const uint32_t sbuf = out_base + pair * (uint32_t)kStageBytes;
        tma_store_4d(tmC, sbuf, nc, x0, y0, img);
        tma_store_4d(tmC, sbuf + (uint32_t)kStageBytes, p.cout + nc, x0, y0, img);
        bulk_commit();
        bulk_wait_read<1>();                                    // the group that used the other pair has been read
      }
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
  }
  if (e0) bulk_wait<0>();
}

// ---- mask-head tail fused into the 2x2 stride-2 transposed convolution (TimeDistributedMaskLayer.swift:58-89):
// N tile nt = sub-pixel (dy,dx) with all deconv_c channels; per output pixel
//   m = sigmoid(b[cls] + sum_c fp16(relu(acc[c] + bias[c])) * w[cls][c])
// for the class of the detection in that slot only (the reference computes all 81 planes and keeps one).  The
// deconvolution output itself is never written.  Invalid slots produce 0.
template <int BN>
__device__ __forceinline__ void epilogue_maskdot(const ConvGemmParams& p, float* bias_gen, uint32_t tfull0, uint32_t tempty0,
                                                 uint32_t tmem_base, const TileSched ts, int warp, int lane) {
  const int q = warp & 3;
  const int row = q * 32 + lane;
  float* w_gen = bias_gen + BN;
  float* out = reinterpret_cast<float*>(p.out);
  int acc = 0; uint32_t acc_phase = 0;
  for (int tile = ts.first; tile < ts.total; tile += ts.step) {
    int n0, x0, y0, img;
    ts.coords(p, BN, tile, n0, x0, y0, img);
    const int nt = n0 / BN;
    const int x = x0 + (row % p.tw), y = y0 + (row / p.tw);
    const bool pix_ok = (x < p.w_out) && (y < p.h_out) && (img < p.n_img);
    const int im = min(img, p.n_img - 1);
    const int valid = __ldg(p.md_valid + im);
    int cls = __ldg(p.md_cls + im);
    cls = cls < 0 ? 0 : (cls >= p.md_ncls ? p.md_ncls - 1 : cls);
    epi_bar_sync();                        // everybody is done with the previous tile's bias / weight rows
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      bias_gen[i] = __ldg(p.bias + nt * BN + i);
      w_gen[i] = __half2float(__ldg(p.md_w + (size_t)cls * p.deconv_c + i));
    }
    epi_bar_sync();
    mbar_wait(tfull0 + 8u * acc, acc_phase);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
    float dot = 0.0f;
    #pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(t_addr + (uint32_t)(c * 32), v);
      float4 bq[8], wq[8];
      #pragma unroll
      for (int g = 0; g < 8; ++g) {
        bq[g] = reinterpret_cast<const float4*>(bias_gen + c * 32)[g];
        wq[g] = reinterpret_cast<const float4*>(w_gen + c * 32)[g];
      }
      tmem_ld_wait();
      #pragma unroll
      for (int g = 0; g < 8; ++g) {
        float a0 = fmaxf(__uint_as_float(v[4 * g + 0]) + bq[g].x, 0.0f), a1 = fmaxf(__uint_as_float(v[4 * g + 1]) + bq[g].y, 0.0f);
        float a2 = fmaxf(__uint_as_float(v[4 * g + 2]) + bq[g].z, 0.0f), a3 = fmaxf(__uint_as_float(v[4 * g + 3]) + bq[g].w, 0.0f);
        if (!p.md_precise) {               // the unfused graph stores this activation as fp16
          a0 = __half2float(__float2half_rn(a0)); a1 = __half2float(__float2half_rn(a1));
          a2 = __half2float(__float2half_rn(a2)); a3 = __half2float(__float2half_rn(a3));
        }
        dot = fmaf(a0, wq[g].x, dot); dot = fmaf(a1, wq[g].y, dot); dot = fmaf(a2, wq[g].z, dot); dot = fmaf(a3, wq[g].w, dot);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) tempty_arrive(ts, tempty0 + 8u * acc);
    if (pix_ok) {
      const int dy = nt >> 1, dx = nt & 1;
      const float m = valid ? 1.0f / (1.0f + expf(-(dot + __ldg(p.md_b + cls)))) : 0.0f;
      out[((size_t)img * (2 * p.h_out) + (2 * y + dy)) * (size_t)(2 * p.w_out) + (2 * x + dx)] = m;
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
  }
}

}  // namespace cg

template <int BN, int CTAS>
__global__ void __launch_bounds__(CG_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ ConvGemmParams p) {
  using C = cg::Cfg<BN, CTAS>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-B aligned operand ring (SWIZZLE_128B requirement)
  const uint32_t smem_base = (cg::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nst = p.nstages;
  const bool vg = p.vgroup > 1;
  const int nsa = vg ? p.na_stages : 0;
  // vgroup > 1: [patch ring: na_stages x patch_bytes][B ring: nstages x kBBytes]; else [nstages x (A tile + B tile)]
  const uint32_t b_ring = smem_base + (uint32_t)nsa * (uint32_t)p.patch_bytes;
  const uint32_t ring_bytes = vg ? (uint32_t)nsa * (uint32_t)p.patch_bytes + (uint32_t)nst * (uint32_t)C::kBBytes
                                 : (uint32_t)nst * (uint32_t)C::kStageBytes;
  const uint32_t out_bytes = (uint32_t)C::kOutStageBytes << p.nbuf_log2;
  const uint32_t out_base = smem_base + ring_bytes;                      // nbuf x 16 KB output / residual staging
  const uint32_t bar_base = out_base + out_bytes;
  // barriers (fixed layout): full[8] @0, empty[8] @64, tmem_full[2] @128, tmem_empty[2] @144, TMEM base slot @160,
  // patch_full[3] @168, patch_empty[3] @192, res_full[8] @216, store_full[8] @280, store_free[8] @344 (block of Cfg::kBarBytes = 416)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto tfull_bar = [&](int a) { return bar_base + 128u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 144u + 8u * a; };
  const uint32_t tmem_slot = bar_base + 160u;
  auto afull_bar = [&](int s) { return bar_base + 168u + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 192u + 8u * s; };
  auto rfull_bar = [&](int b) { return bar_base + 216u + 8u * b; };
  auto sfull_bar = [&](int b) { return bar_base + 280u + 8u * b; };
  auto sfree_bar = [&](int b) { return bar_base + 344u + 8u * b; };
  uint8_t* smem_gen = smem_raw + (smem_base - cg::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + ring_bytes + out_bytes + 160u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = (CTAS == 2) ? cg::cluster_ctarank() : 0u;      // rank inside the CTA pair; 0 = leader (issues the MMAs)

  if (warp == 0 && lane == 0) {
    cg::prefetch_tmap(&tmA);
    cg::prefetch_tmap(&tmB);
    // full[s]: one arrival (the leader's expect_tx) + the TMA bytes of every CTA of the pair; tmem_empty: 4 epilogue
    // warps per CTA of the pair (the peer's warps arrive remotely on the leader's barrier)
    for (int s = 0; s < nst; ++s) { cg::mbar_init(full_bar(s), 1); cg::mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { cg::mbar_init(tfull_bar(a), 1); cg::mbar_init(tempty_bar(a), 4 * CTAS); }
    for (int b = 0; b < 8; ++b) cg::mbar_init(rfull_bar(b), 1);
    for (int a = 0; a < 3; ++a) { cg::mbar_init(afull_bar(a), 1); cg::mbar_init(aempty_bar(a), 1); }
    for (int b = 0; b < 8; ++b) { cg::mbar_init(sfull_bar(b), 4); cg::mbar_init(sfree_bar(b), 1); }
    cg::fence_barrier_init();
  }
  if (warp == 1) { if (CTAS == 2) cg::tmem_alloc_2cta(tmem_slot, C::kTmemCols); else cg::tmem_alloc(tmem_slot, C::kTmemCols); }
  cg::tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cg::cluster_sync_all();     // the peer's barriers are initialised before anything is signalled on them
  cg::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps
  // the tail of the previous kernel in the stream; from here on this grid reads what that kernel wrote.  The
  // trigger comes after the wait, so a dependent grid can only be scheduled once every CTA of this grid is running.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tiles_m = p.n_img * p.tiles_y * p.tiles_x;
  cg::TileSched ts;
  ts.ctas = CTAS; ts.crank = (int)crank;
  ts.first = (int)blockIdx.x / CTAS; ts.step = (int)gridDim.x / CTAS;
  ts.total = ((tiles_m + CTAS - 1) / CTAS) * p.tiles_n;
  const int cin_chunks = p.cin / CG_BK;
  const int nkb = p.ntaps * cin_chunks;

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its share of B) =====================
    if (lane == 0 && vg) {
      // patch ring + B ring (see ConvGemmParams::vgroup)
      int sb = 0; uint32_t phb = 0;
      int sa = 0; uint32_t pha = 0;
      const int V = p.vgroup, ngroups = p.ntaps / V;
      const uint32_t a_tx = (uint32_t)p.tw * (uint32_t)(p.th + V - 1) * 128u * (uint32_t)CTAS;
      for (int w = ts.first; w < ts.total; w += ts.step) {
        int n0, px0, py0, img;
        ts.coords(p, BN, w, n0, px0, py0, img);
        const int x0 = px0 * p.stride, y0 = py0 * p.stride;
        const int nb0 = n0 + (int)crank * (BN / CTAS);
        for (int g = 0; g < ngroups; ++g) {
          const int xi = x0 + p.tap_dx[g * V], yi = y0 + p.tap_dy[g * V];
          for (int cc = 0; cc < cin_chunks; ++cc) {
            cg::mbar_wait(aempty_bar(sa), pha ^ 1u);
            const uint32_t a_dst = smem_base + (uint32_t)sa * (uint32_t)p.patch_bytes;
            if (CTAS == 2) {
              if (crank == 0) cg::mbar_expect_tx(afull_bar(sa), a_tx);
              cg::tma_load_4d_2cta(a_dst, &tmA, cg::mapa_rank(afull_bar(sa), 0), cc * CG_BK, xi, yi, img);
            } else {
              cg::mbar_expect_tx(afull_bar(sa), a_tx);
              cg::tma_load_4d(a_dst, &tmA, afull_bar(sa), cc * CG_BK, xi, yi, img);
            }
            if (++sa == nsa) { sa = 0; pha ^= 1u; }
            for (int v = 0; v < V; ++v) {
              cg::mbar_wait(empty_bar(sb), phb ^ 1u);
              const uint32_t b_dst = b_ring + (uint32_t)sb * (uint32_t)C::kBBytes;
              const int kcol = (int)p.tap_widx[g * V + v] * p.cin + cc * CG_BK;
              if (CTAS == 2) {
                if (crank == 0) cg::mbar_expect_tx(full_bar(sb), (uint32_t)(2 * C::kBBytes));
                cg::tma_load_2d_2cta(b_dst, &tmB, cg::mapa_rank(full_bar(sb), 0), kcol, nb0);
              } else {
                cg::mbar_expect_tx(full_bar(sb), (uint32_t)C::kBBytes);
                cg::tma_load_2d(b_dst, &tmB, full_bar(sb), kcol, nb0);
              }
              if (++sb == nst) { sb = 0; phb ^= 1u; }
            }
          }
        }
      }
    } else if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      cg::TraceCursor tc = cg::trace_open(p, 0);
      for (int w = ts.first; w < ts.total; w += ts.step) {
        int n0, px0, py0, img;
        ts.coords(p, BN, w, n0, px0, py0, img);
        const int x0 = px0 * p.stride, y0 = py0 * p.stride;
        const int nb0 = n0 + (int)crank * (BN / CTAS);
        for (int t = 0; t < p.ntaps; ++t) {
          const int xi = x0 + p.tap_dx[t], yi = y0 + p.tap_dy[t];
          for (int cc = 0; cc < cin_chunks; ++cc) {
            cg::mbar_wait(empty_bar(stage), phase ^ 1u);
            cg::trace_ev(tc, 1, (uint32_t)stage);             // producer: slot free, issuing loads
            const uint32_t a_dst = smem_base + stage * C::kStageBytes;
            const uint32_t b_dst = a_dst + C::kABytes;
            if (CTAS == 2) {
              if (crank == 0) cg::mbar_expect_tx(full_bar(stage), (uint32_t)(2 * C::kStageBytes));
              const uint32_t lbar = cg::mapa_rank(full_bar(stage), 0);
              cg::tma_load_4d_2cta(a_dst, &tmA, lbar, cc * CG_BK, xi, yi, img);
              cg::tma_load_2d_2cta(b_dst, &tmB, lbar, t * p.cin + cc * CG_BK, nb0);
            } else {
              cg::mbar_expect_tx(full_bar(stage), (uint32_t)C::kStageBytes);
              cg::tma_load_4d(a_dst, &tmA, full_bar(stage), cc * CG_BK, xi, yi, img);
              cg::tma_load_2d(b_dst, &tmB, full_bar(stage), t * p.cin + cc * CG_BK, nb0);
            }
            if (++stage == nst) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread; of the leader CTA in a pair) =====================
    if (lane == 0 && crank == 0 && vg) {
      constexpr uint32_t idesc = cg::make_idesc_f16(CG_BM * CTAS, BN);
      int sb = 0; uint32_t phb = 0;
      int sa = 0; uint32_t pha = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const int V = p.vgroup, npatch = (p.ntaps / V) * cin_chunks;
      const uint32_t row_step = (uint32_t)p.tw * 128u;          // one tile row further down in the patch (a multiple of 1024 B)
      for (int w = ts.first; w < ts.total; w += ts.step) {
        cg::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        cg::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t first = 0u;
        for (int pi = 0; pi < npatch; ++pi) {
          cg::mbar_wait(afull_bar(sa), pha);
          const uint32_t a_addr = smem_base + (uint32_t)sa * (uint32_t)p.patch_bytes;
          for (int v = 0; v < V; ++v) {
            cg::mbar_wait(full_bar(sb), phb);
            cg::tc_fence_after();
            const uint64_t adesc = cg::make_sw128_desc(a_addr + (uint32_t)v * row_step);
            const uint64_t bdesc = cg::make_sw128_desc(b_ring + (uint32_t)sb * (uint32_t)C::kBBytes);
            #pragma unroll
            for (int k = 0; k < CG_BK / 16; ++k) {
              if (CTAS == 2) cg::umma_f16_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first | (uint32_t)k);
              else cg::umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first | (uint32_t)k);
            }
            first = 1u;
            if (CTAS == 2) cg::umma_commit_2cta(empty_bar(sb)); else cg::umma_commit(empty_bar(sb));      // B slot free when these MMAs retire
            if (++sb == nst) { sb = 0; phb ^= 1u; }
          }
          if (CTAS == 2) cg::umma_commit_2cta(aempty_bar(sa)); else cg::umma_commit(aempty_bar(sa));      // patch free after its last window
          if (++sa == nsa) { sa = 0; pha ^= 1u; }
        }
        if (CTAS == 2) cg::umma_commit_2cta(tfull_bar(acc)); else cg::umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    } else if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = cg::make_idesc_f16(CG_BM * CTAS, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      cg::TraceCursor tc = cg::trace_open(p, 1);
      for (int w = ts.first; w < ts.total; w += ts.step) {
        cg::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);     // every epilogue warp (of both CTAs) has drained this accumulator
        cg::trace_ev(tc, 3, (uint32_t)w);                     // MMA: accumulator free, tile start
        cg::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          cg::mbar_wait(full_bar(stage), phase);              // TMA bytes (of both CTAs) have landed
          cg::trace_ev(tc, 2, (uint32_t)kb);                   // MMA: operands landed
          cg::tc_fence_after();
          const uint32_t a_addr = smem_base + stage * C::kStageBytes;
          const uint64_t adesc = cg::make_sw128_desc(a_addr);
          const uint64_t bdesc = cg::make_sw128_desc(a_addr + C::kABytes);
          #pragma unroll
          for (int k = 0; k < CG_BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle atom: +2 in 16-B units
            if (CTAS == 2) cg::umma_f16_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            else cg::umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs) when the MMAs retire
          if (CTAS == 2) cg::umma_commit_2cta(empty_bar(stage)); else cg::umma_commit(empty_bar(stage));
          if (++stage == nst) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CTAS == 2) cg::umma_commit_2cta(tfull_bar(acc)); else cg::umma_commit(tfull_bar(acc));
        cg::trace_ev(tc, 4, (uint32_t)w);                     // MMA: all MMAs of the tile issued
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp == 6) {
    // ===================== store thread of the staged epilogue =====================
    if (lane == 0 && p.tma_out && !p.split_out && !p.maskdot) {
      const uint32_t rf0 = rfull_bar(0), sf0 = sfull_bar(0), sr0 = sfree_bar(0);
      if (p.tma_res) cg::store_thread_staged<BN, 1>(p, &tmC, &tmR, out_base, rf0, sf0, sr0, ts);
      else if (p.res_mode == 2) cg::store_thread_staged<BN, 2>(p, &tmC, &tmR, out_base, rf0, sf0, sr0, ts);
      else cg::store_thread_staged<BN, 0>(p, &tmC, &tmR, out_base, rf0, sf0, sr0, ts);
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // accumulator row = pixel inside the tile
    int acc = 0; uint32_t acc_phase = 0;
    if (p.maskdot) {
      float* bias_gen = reinterpret_cast<float*>(smem_gen + ring_bytes);       // bias + class weights live in the (unused) staging area
      cg::epilogue_maskdot<BN>(p, bias_gen, tfull_bar(0), tempty_bar(0), tmem_base, ts, warp, lane);
    } else if (p.split_out) {
      uint8_t* out_gen = smem_gen + ring_bytes;
      float* bias_gen = reinterpret_cast<float*>(out_gen + out_bytes + C::kBarBytes);
      cg::epilogue_split<BN>(p, &tmC, out_gen, out_base, bias_gen, tfull_bar(0), tempty_bar(0), tmem_base, ts, warp, lane);
    } else if (p.tma_out) {
      const uint32_t rf0 = rfull_bar(0), tf0 = tfull_bar(0), te0 = tempty_bar(0);
      uint8_t* out_gen = smem_gen + ring_bytes;
      float* bias_gen = reinterpret_cast<float*>(out_gen + out_bytes + C::kBarBytes);
      const uint32_t sf0 = sfull_bar(0), sr0 = sfree_bar(0);
      if (p.tma_res) cg::epilogue_staged<BN, 1>(p, out_gen, bias_gen, rf0, sf0, sr0, tf0, te0, tmem_base, ts, warp, lane);
      else if (p.res_mode == 2) cg::epilogue_staged<BN, 2>(p, out_gen, bias_gen, rf0, sf0, sr0, tf0, te0, tmem_base, ts, warp, lane);
      else cg::epilogue_staged<BN, 0>(p, out_gen, bias_gen, rf0, sf0, sr0, tf0, te0, tmem_base, ts, warp, lane);
    } else
    for (int tile = ts.first; tile < ts.total; tile += ts.step) {
      int n0, tx0, ty0, img;
      ts.coords(p, BN, tile, n0, tx0, ty0, img);
      const int x = tx0 + (row % p.tw);
      const int y = ty0 + (row / p.tw);
      const bool pix_ok = (x < p.w_out) && (y < p.h_out) && (img < p.n_img);
      cg::mbar_wait(tfull_bar(acc), acc_phase);
      cg::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      // output / residual row pointers
      size_t out_off;
      int ch_base = n0;
      if (p.deconv) {
        const int sub = n0 / p.deconv_c;               // BN divides deconv_c or vice versa is enforced by the host
        const int dy = sub >> 1, dx = sub & 1;
        out_off = (((size_t)img * (2 * p.h_out) + (2 * y + dy)) * (size_t)(2 * p.w_out) + (2 * x + dx)) * (size_t)p.ldc;
        ch_base = n0 - sub * p.deconv_c;
      } else {
        out_off = (((size_t)img * p.h_out + y) * (size_t)p.w_out + x) * (size_t)p.ldc;
      }
      const __half* res_row = nullptr;
      if (p.res_mode && pix_ok) {
        const int ry = (p.res_mode == 2) ? (y >> 1) : y;
        const int rx = (p.res_mode == 2) ? (x >> 1) : x;
        res_row = p.residual + (((size_t)img * p.res_h + ry) * (size_t)p.res_w + rx) * (size_t)p.res_ld;
      }
      #pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        const int nc = n0 + chunk * 32;                 // first output channel of this chunk
        if (nc >= p.ldc && !p.deconv) break;            // warp-uniform
        uint32_t v[32];
        cg::tmem_ld32(t_addr + (uint32_t)(chunk * 32), v);
        cg::tmem_ld_wait();
        if (pix_ok) {
          #pragma unroll
          for (int g = 0; g < 4; ++g) {                 // 4 groups of 8 channels
            const int n = nc + g * 8;
            const int cdst = ch_base + chunk * 32 + g * 8;
            if (!p.deconv && n >= p.ldc) break;
            float f[8];
            #pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g * 8 + e]);
            if (p.bias) {
              #pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += (n + e < p.cout) ? __ldg(p.bias + n + e) : 0.0f;
            }
            if (res_row) {
              const uint4 rv = __ldg(reinterpret_cast<const uint4*>(res_row + n));
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
              #pragma unroll
              for (int e = 0; e < 4; ++e) { float2 r2 = __half22float2(rh[e]); f[2 * e] += r2.x; f[2 * e + 1] += r2.y; }
            }
            if (p.relu) {
              #pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);
            }
            if (p.out_f32) {
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + cdst);
              o[0] = make_float4(f[0], f[1], f[2], f[3]);
              o[1] = make_float4(f[4], f[5], f[6], f[7]);
            } else {
              uint4 ov;
              __half2* oh = reinterpret_cast<__half2*>(&ov);
              #pragma unroll
              for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + out_off + cdst) = ov;
            }
          }
        }
      }
      cg::tc_fence_before();
      __syncwarp();
      if (lane == 0) cg::tempty_arrive(ts, tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  // teardown: nobody leaves (or frees TMEM) while the partner may still read this CTA's smem / signal its barriers
  cg::tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cg::cluster_sync_all();
  if (warp == 1) {
    cg::tc_fence_after();
    if (CTAS == 2) cg::tmem_dealloc_2cta(tmem_base, C::kTmemCols); else cg::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}
