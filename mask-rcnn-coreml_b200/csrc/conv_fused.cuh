// Two back-to-back 1x1 convolutions of consecutive ResNet bottleneck blocks in ONE kernel (sm_100a, tcgen05 / TMEM / TMA):
//
//   X = relu(A * W1^T + b1 + R)        block k:     1x1 expansion c1 -> n1 channels with the residual R   ("2c")
//   Y = relu(X * W2^T + b2)            block k + 1: 1x1 reduction n1 -> n2 channels                        ("2a")
//
// per 128-pixel tile.  X is written to global memory (it is the next block's residual) AND kept in shared memory as the A
// operand of the second GEMM: the staging buffers the epilogue converts X into (128 pixels x 64 channels fp16,
// SWIZZLE_128B, the TMA store's source) are exactly a K-major UMMA operand block.  The separate launches read X back from
// DRAM (67 MB per block at batch 8: it does not survive in L2) and pay a second launch head and tail; here X is read
// once, by the tensor core, from shared memory.  Results are bit-identical to the two launches (same K order per
// output row, same epilogue arithmetic): tests/test_conv_gpu.py::test_fused_expand_reduce_bit_identical.
//
// X is produced in chunks of 128 channels (N = 128 MMAs into one of two 128-column TMEM accumulators), each chunk is two
// staging buffers = two K blocks of the second GEMM, which accumulates all chunks into a third, n2-column accumulator:
//
//   TMEM (512 columns)   [acc1 #0: 128][acc1 #1: 128][acc2: n2 <= 256]
//   shared memory        A tile (c1/64 K blocks x 16 KB, resident for the tile) | weight ring 96 KB (a slot = two K blocks
//                        of W1 for one chunk, or one K block of W2) | 4 staging buffers x 16 KB | barriers, bias
//   warp 0  producer     A tile, then the weight slots in the order the MMA warp consumes them
//   warp 1  MMA issuer   GEMM1(j) ... GEMM1(j+1), GEMM2(j), ...: the second GEMM lags one chunk so that the epilogue of
//                        chunk j overlaps the MMAs of chunk j + 1  (CTA pairs: the leader's; the peer's is the relay thread)
//   warps 2-5 epilogue   chunk j: TMEM -> +bias +residual, ReLU, fp16 -> staging (in place over the TMA-loaded residual);
//                        after the last chunk: acc2 -> +bias, ReLU -> staging (Y)
//   warp 6  store thread TMA stores of X / Y sub-chunks; frees a buffer when its store has read it AND the second GEMM has
//                        consumed it, then fetches the residual of the buffer's next chunk into it
#pragma once
#include "dense.h"

// struct FusedParams: dense.h (the host plan embeds it; this header defines a kernel and belongs to one translation unit)

namespace cgf {
using namespace cg;

constexpr int kABytes = 16384;                       // one K block of the A tile: 128 pixels x 64 channels
constexpr int kRingBytes = 98304;                    // weight ring: 3 slots of 32 KB (one CTA per tile) or 6 of 16 KB (CTA pairs: a
                                                     // CTA stages half of every weight tile, so the same bytes hold twice the K blocks)
constexpr int kStageBytes = 16384;                   // staging buffer: 128 pixels x 64 channels
constexpr int kAOff = 0, kRingOff = 4 * kABytes, kStageOff = kRingOff + kRingBytes, kBarOff = kStageOff + 4 * kStageBytes;
constexpr int kBarBytes = 384;
constexpr int kBiasOff = kBarOff + kBarBytes;        // bias1 chunk (2 x 128 floats) + bias2 (256 floats)
constexpr int kSmemBytes = kBiasOff + (2 * 128 + 256) * 4;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can have");

// 64 accumulator columns of this thread's row -> (+bias, +residual already in the staging row, ReLU) -> fp16 -> staging row
template <bool RES>
__device__ __forceinline__ void convert_subchunk(uint32_t t_addr, const float* bias64, uint8_t* srow, uint32_t sw, __half2 lo2,
                                                 TraceCursor* tc = nullptr) {
  #pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    uint32_t v[32];
    tmem_ld32(t_addr + (uint32_t)(hh * 32), v);
    float4 bq[8];
    uint4 rq[4];
    #pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4* bp = reinterpret_cast<const float4*>(bias64 + hh * 32 + g * 8);
      bq[2 * g] = bp[0]; bq[2 * g + 1] = bp[1];
      if (RES) rq[g] = *reinterpret_cast<const uint4*>(srow + ((((uint32_t)(hh * 4 + g)) ^ sw) << 4));
    }
    tmem_ld_wait();
    if (tc) trace_ev(*tc, 15, (uint32_t)hh);           // debug: TMEM half landed
    #pragma unroll
    for (int g = 0; g < 4; ++g) {
      float2 f[4];
      f[0] = add2(make_float2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1])), make_float2(bq[2 * g].x, bq[2 * g].y));
      f[1] = add2(make_float2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3])), make_float2(bq[2 * g].z, bq[2 * g].w));
      f[2] = add2(make_float2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5])), make_float2(bq[2 * g + 1].x, bq[2 * g + 1].y));
      f[3] = add2(make_float2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])), make_float2(bq[2 * g + 1].z, bq[2 * g + 1].w));
      if (RES) {
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
}

// debug event trace, same buffer layout as cg::trace_ev (role sections 0 producer, 1 MMA, 2 epilogue; tools/trace_fused.py)
__device__ __forceinline__ TraceCursor ftrace_open(const FusedParams& p, int role) {
  TraceCursor c; c.n = 0;
  c.base = p.trace ? p.trace + ((size_t)blockIdx.x * 3 + role) * (2 * CG_TRACE_MAX + 2) : nullptr;
  return c;
}

__device__ __forceinline__ void tile_coords(const FusedParams& p, int t, int& x0, int& y0, int& img) {
  x0 = (t % p.tiles_x) * p.tw;
  y0 = ((t / p.tiles_x) % p.tiles_y) * p.th;
  img = t / (p.tiles_x * p.tiles_y);
}
}  // namespace cgf

// CTAS = 2: a CTA pair computes two adjacent pixel tiles with tcgen05.mma.cta_group::2 (M = 256, issued by the leader): each CTA
// stages its own A tile, its own residual / X / Y sub-chunks and HALF of every weight tile, so the weight ring holds twice as many
// K blocks -- with one CTA per tile the MMA warp waits for weights most of the time (128 KB per chunk through a 96 KB ring at the
// ~4000 clk the L2 answers in under this load; tools/trace_fused.py).  Measured: shorter chunk period in isolation, no gain in
// the pipeline ("Kept"): the host launches CTAS = 1 unless MRCNN_FUSE_CTAS=2.
template <int CTAS>
__global__ void __launch_bounds__(CG_THREADS, 1)
conv_fused_expand_reduce_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                                const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmR,
                                const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                                const __grid_constant__ FusedParams p) {
  using namespace cgf;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = cg::smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();                   // the layout has no slack for re-alignment: dynamic smem starts 1024-aligned
  const uint32_t a_base = smem_base + kAOff, ring = smem_base + kRingOff, stg = smem_base + kStageOff, bar = smem_base + kBarOff;
  // barriers: a_full @0, a_empty @8, b_full[6] @16, b_empty[6] @64, t1_full[2] @112, t1_empty[2] @128, t2_full @144, t2_empty @152,
  // r_full[4] @160, s_full[4] @192, x_full[4] @224, x_done[4] @256, s_free[4] @288; CTA pairs, on the leader: c_peer[2] @320 (the
  // peer has drained acc1 #p and staged its X chunk), t2e_peer @336 (the peer has drained acc2) -- one remote arrival each, made
  // by the peer's relay thread; TMEM base slot @344 (block of kBarBytes = 384)
  const uint32_t a_full = bar, a_empty = bar + 8;
  auto b_full = [&](int s) { return bar + 16u + 8u * s; };
  auto b_empty = [&](int s) { return bar + 64u + 8u * s; };
  auto t1_full = [&](int a) { return bar + 112u + 8u * a; };
  auto t1_empty = [&](int a) { return bar + 128u + 8u * a; };
  const uint32_t t2_full = bar + 144, t2_empty = bar + 152;
  auto r_full = [&](int b) { return bar + 160u + 8u * b; };
  auto s_full = [&](int b) { return bar + 192u + 8u * b; };
  auto x_full = [&](int b) { return bar + 224u + 8u * b; };
  auto x_done = [&](int b) { return bar + 256u + 8u * b; };
  auto s_free = [&](int b) { return bar + 288u + 8u * b; };
  auto c_peer = [&](int a) { return bar + 320u + 8u * a; };
  const uint32_t t2e_peer = bar + 336u;
  const uint32_t tmem_slot = bar + 344u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + kBarOff + 344);
  float* bias1_s = reinterpret_cast<float*>(smem_raw + kBiasOff);             // [2][128]
  float* bias2_s = bias1_s + 256;                                              // [n2 <= 256]

  constexpr int kSlots = 3 * CTAS;
  constexpr uint32_t kSlotBytes = 32768u / CTAS, kHalfSlot = kSlotBytes / 2;     // a W1 slot = two K blocks, kHalfSlot apart
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB1 = p.c1 / 64, NCH = p.n1 / 128, NY = p.n2 / 64;
  const uint32_t crank = (CTAS == 2) ? cg::cluster_ctarank() : 0u;          // rank in the pair; 0 = leader (issues the MMAs)
  // TMA bytes of both CTAs land on the leader's barriers.  The epilogue warps arrive on LOCAL barriers only: a remote
  // mbarrier.arrive.release.cluster costs the issuing warp ~1900 clk (tools/trace_fused.py), so the peer's completions are
  // forwarded to the leader by a relay thread (the peer's otherwise idle MMA warp), one remote arrival per X chunk.
  auto commit = [&](uint32_t b) { if (CTAS == 2) cg::umma_commit_2cta(b); else cg::umma_commit(b); };
  if (warp == 0 && lane == 0) {
    cg::prefetch_tmap(&tmA); cg::prefetch_tmap(&tmB1); cg::prefetch_tmap(&tmB2);
    cg::mbar_init(a_full, 1); cg::mbar_init(a_empty, 1);
    for (int s = 0; s < kSlots; ++s) { cg::mbar_init(b_full(s), 1); cg::mbar_init(b_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { cg::mbar_init(t1_full(a), 1); cg::mbar_init(t1_empty(a), 4); cg::mbar_init(c_peer(a), 1); }
    cg::mbar_init(t2_full, 1); cg::mbar_init(t2_empty, 4); cg::mbar_init(t2e_peer, 1);
    for (int b = 0; b < 4; ++b) {
      cg::mbar_init(r_full(b), 1); cg::mbar_init(s_full(b), 4); cg::mbar_init(x_full(b), 4); cg::mbar_init(x_done(b), 1);
      cg::mbar_init(s_free(b), 1);
    }
    cg::fence_barrier_init();
  }
  if (warp == 1) { if (CTAS == 2) cg::tmem_alloc_2cta(tmem_slot, 512); else cg::tmem_alloc(tmem_slot, 512); }
  cg::tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cg::cluster_sync_all();             // the peer's barriers are initialised before anything is signalled on them
  cg::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // work item w = first, first + step, ... < items: a tile (CTAS = 1) or a pair of adjacent tiles, of which this CTA owns tile
  // CTAS * w + crank (the phantom tile of an odd count lies outside the tensor: TMA zero-fills its loads and clips its stores)
  const int items = (p.n_img * p.tiles_y * p.tiles_x + CTAS - 1) / CTAS;
  const int first = (int)blockIdx.x / CTAS, step = (int)gridDim.x / CTAS;
  const int tiles = items;                            // loop bound of every role below
  auto my_tile = [&](int w) { return CTAS * w + (int)crank; };

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      int slot = 0; uint32_t sph = 0;
      uint32_t tl = 0;                                 // local tile counter
      const uint32_t rows1 = 128u / CTAS, rows2 = (uint32_t)p.n2 / CTAS;     // rows of a W1 / W2 tile this CTA stages
      auto load_b2 = [&](int c) {
        for (int i = 0; i < 2; ++i) {
          cg::mbar_wait(b_empty(slot), sph ^ 1u);
          const uint32_t dst = ring + (uint32_t)slot * kSlotBytes;
          if (CTAS == 2) {
            if (crank == 0) cg::mbar_expect_tx(b_full(slot), (uint32_t)p.n2 * 128u);
            cg::tma_load_2d_2cta(dst, &tmB2, cg::mapa_rank(b_full(slot), 0), c * 128 + i * 64, (int)(crank * rows2));
          } else {
            cg::mbar_expect_tx(b_full(slot), (uint32_t)p.n2 * 128u);
            cg::tma_load_2d(dst, &tmB2, b_full(slot), c * 128 + i * 64, 0);
          }
          if (++slot == kSlots) { slot = 0; sph ^= 1u; }
        }
      };
      for (int t = first; t < tiles; t += step, ++tl) {
        int x0, y0, img;
        tile_coords(p, my_tile(t), x0, y0, img);
        cg::mbar_wait(a_empty, (tl & 1u) ^ 1u);
        if (CTAS == 2) {
          if (crank == 0) cg::mbar_expect_tx(a_full, (uint32_t)(2 * KB1) * kABytes);
          const uint32_t lbar = cg::mapa_rank(a_full, 0);
          for (int kb = 0; kb < KB1; ++kb) cg::tma_load_4d_2cta(a_base + (uint32_t)kb * kABytes, &tmA, lbar, kb * 64, x0, y0, img);
        } else {
          cg::mbar_expect_tx(a_full, (uint32_t)KB1 * kABytes);
          for (int kb = 0; kb < KB1; ++kb) cg::tma_load_4d(a_base + (uint32_t)kb * kABytes, &tmA, a_full, kb * 64, x0, y0, img);
        }
        for (int j = 0; j < NCH; ++j) {
          for (int kb = 0; kb < KB1; kb += 2) {
            const int nk = min(2, KB1 - kb);
            cg::mbar_wait(b_empty(slot), sph ^ 1u);
            // a slot holds two K blocks of the chunk's W1 rows: [K block kb][K block kb + 1], each rows1 x 128 B
            if (CTAS == 2) {
              if (crank == 0) cg::mbar_expect_tx(b_full(slot), (uint32_t)nk * 16384u);
              const uint32_t lbar = cg::mapa_rank(b_full(slot), 0);
              for (int i = 0; i < nk; ++i)
                cg::tma_load_2d_2cta(ring + (uint32_t)slot * kSlotBytes + (uint32_t)i * kHalfSlot, &tmB1, lbar, (kb + i) * 64, j * 128 + (int)(crank * rows1));
            } else {
              cg::mbar_expect_tx(b_full(slot), (uint32_t)nk * 16384u);
              for (int i = 0; i < nk; ++i)
                cg::tma_load_2d(ring + (uint32_t)slot * kSlotBytes + (uint32_t)i * kHalfSlot, &tmB1, b_full(slot), (kb + i) * 64, j * 128);
            }
            if (++slot == kSlots) { slot = 0; sph ^= 1u; }
          }
          if (j >= 1) load_b2(j - 1);
        }
        load_b2(NCH - 1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (of the leader CTA in a pair) =====================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc1 = cg::make_idesc_f16(128 * CTAS, 128);
      const uint32_t idesc2 = cg::make_idesc_f16(128 * CTAS, p.n2);
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum) {
        if (CTAS == 2) cg::umma_f16_2cta(d, ad, bd, idesc, accum); else cg::umma_f16(d, ad, bd, idesc, accum);
      };
      int slot = 0; uint32_t sph = 0;
      uint32_t tl = 0;
      uint32_t t1_uses[2] = {0, 0};
      uint32_t xpar = 0;                               // bit b: parity of the X uses of staging buffer b consumed so far
      const uint32_t acc2 = tmem_base + 256u;
      uint32_t cp_uses[2] = {0, 0};
      auto mma2 = [&](int c) {
        if (c == 0) {
          cg::mbar_wait(t2_empty, (tl & 1u) ^ 1u);
          if (CTAS == 2) cg::mbar_wait(t2e_peer, (tl & 1u) ^ 1u);
          cg::tc_fence_after();
        }
        if (CTAS == 2) {                                // the peer's half of X chunk c is staged (and its acc1 buffer drained)
          cg::mbar_wait(c_peer(c & 1), cp_uses[c & 1] & 1u);
          ++cp_uses[c & 1];
        }
        for (int i = 0; i < 2; ++i) {
          const int b = 2 * (c & 1) + i;
          cg::mbar_wait(x_full(b), (xpar >> b) & 1u);
          xpar ^= 1u << b;
          cg::mbar_wait(b_full(slot), sph);
          cg::tc_fence_after();
          const uint64_t adesc = cg::make_sw128_desc(stg + (uint32_t)b * kStageBytes);
          const uint64_t bdesc = cg::make_sw128_desc(ring + (uint32_t)slot * kSlotBytes);
          #pragma unroll
          for (int k = 0; k < 4; ++k)
            mma(acc2, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc2, (c | i | k) ? 1u : 0u);
          commit(b_empty(slot));
          commit(x_done(b));
          if (++slot == kSlots) { slot = 0; sph ^= 1u; }
        }
      };
      TraceCursor tc = ftrace_open(p, 1);
      for (int t = first; t < tiles; t += step, ++tl) {
        cg::mbar_wait(a_full, tl & 1u);
        cg::trace_ev(tc, 3, (uint32_t)t);               // MMA: A tile landed
        for (int j = 0; j < NCH; ++j) {
          const int acc = j & 1;
          cg::mbar_wait(t1_empty(acc), (t1_uses[acc] & 1u) ^ 1u);
          ++t1_uses[acc];
          cg::trace_ev(tc, 10, (uint32_t)j);             // MMA: acc1 buffer free, GEMM1(j) starts
          cg::tc_fence_after();
          const uint32_t d1 = tmem_base + (uint32_t)(acc * 128);
          for (int kb = 0; kb < KB1; kb += 2) {
            const int nk = min(2, KB1 - kb);
            cg::mbar_wait(b_full(slot), sph);
            cg::tc_fence_after();
            for (int i = 0; i < nk; ++i) {
              const uint64_t adesc = cg::make_sw128_desc(a_base + (uint32_t)(kb + i) * kABytes);
              const uint64_t bdesc = cg::make_sw128_desc(ring + (uint32_t)slot * kSlotBytes + (uint32_t)i * kHalfSlot);
              #pragma unroll
              for (int k = 0; k < 4; ++k)
                mma(d1, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc1, ((kb + i) | k) ? 1u : 0u);
            }
            commit(b_empty(slot));
            if (++slot == kSlots) { slot = 0; sph ^= 1u; }
          }
          commit(t1_full(acc));
          cg::trace_ev(tc, 11, (uint32_t)j);             // MMA: GEMM1(j) issued (weights landed)
          if (j == NCH - 1) commit(a_empty);                    // the A tile is free once the last chunk's MMAs retire
          if (j >= 1) { mma2(j - 1); cg::trace_ev(tc, 12, (uint32_t)(j - 1)); }     // MMA: GEMM2(j-1) issued (X chunk + weights landed)
        }
        mma2(NCH - 1);
        commit(t2_full);
      }
    }
    if (CTAS == 2 && lane == 0 && crank == 1) {
      // ===================== relay (peer CTA): local completions -> one remote arrival on the leader's barriers =====================
      uint32_t xpar = 0;
      uint32_t tl = 0;
      for (int t = first; t < tiles; t += step, ++tl) {
        for (int j = 0; j < NCH; ++j) {
          const int b = 2 * (j & 1) + 1;                // the chunk's second sub-chunk: every warp arrives there last
          cg::mbar_wait(x_full(b), (xpar >> b) & 1u);
          xpar ^= 1u << b;
          cg::mbar_arrive_cluster(cg::mapa_rank(c_peer(j & 1), 0));
        }
        cg::mbar_wait(t2_empty, tl & 1u);
        cg::mbar_arrive_cluster(cg::mapa_rank(t2e_peer, 0));
      }
    }
  } else if (warp == 6) {
    // ===================== store thread =====================
    if (lane == 0) {
      cg::prefetch_tmap(&tmX); cg::prefetch_tmap(&tmY); cg::prefetch_tmap(&tmR);
      const int my_tiles = first < tiles ? (tiles - 1 - first) / step + 1 : 0;
      auto load_residual = [&](int tl, int j, int s) {        // chunk j, sub-chunk s of local tile tl -> buffer 2 * (j & 1) + s
        if (tl >= my_tiles) return;
        int x0, y0, img;
        tile_coords(p, my_tile(first + tl * step), x0, y0, img);
        const int b = 2 * (j & 1) + s;
        cg::mbar_expect_tx(r_full(b), (uint32_t)kStageBytes);
        cg::tma_load_4d(stg + (uint32_t)b * kStageBytes, &tmR, r_full(b), j * 128 + s * 64, x0, y0, img);
      };
      for (int j = 0; j < 2 && j < NCH; ++j) for (int s = 0; s < 2; ++s) load_residual(0, j, s);
      uint32_t upar = 0, dpar = 0;                     // bit b: parity of the uses of buffer b stored so far / X uses whose MMAs were awaited
      // the item before the current one, freed once its store has been read: kind 0 none, 1 X (j, s), 2 Y (y)
      int pk = 0, ptl = 0, pj = 0, ps = 0;
      auto free_prev = [&]() {
        if (pk == 1) {
          const int b = 2 * (pj & 1) + ps;
          cg::mbar_wait(x_done(b), (dpar >> b) & 1u);          // the second GEMM has read it
          dpar ^= 1u << b;
          if (pj + 2 < NCH) load_residual(ptl, pj + 2, ps);
          else if (b < NY) cg::mbar_arrive(s_free(b));         // next use: a Y sub-chunk of this tile
          else load_residual(ptl + 1, pj & 1, ps);             // next use: the same chunk slot of the next tile
        } else if (pk == 2) {
          load_residual(ptl + 1, ps >> 1, ps & 1);             // Y item in buffer ps: next use is chunk (b >> 1), sub-chunk (b & 1) of the next tile
        }
      };
      for (int tl = 0; tl < my_tiles; ++tl) {
        int x0, y0, img;
        tile_coords(p, my_tile(first + tl * step), x0, y0, img);
        for (int j = 0; j < NCH; ++j)
          for (int s = 0; s < 2; ++s) {
            const int b = 2 * (j & 1) + s;
            cg::mbar_wait(s_full(b), (upar >> b) & 1u);
            upar ^= 1u << b;
            cg::tma_store_4d(&tmX, stg + (uint32_t)b * kStageBytes, j * 128 + s * 64, x0, y0, img);
            cg::bulk_commit();
            cg::bulk_wait_read<1>();                            // every store but this one has read its buffer
            free_prev();
            pk = 1; ptl = tl; pj = j; ps = s;
          }
        for (int y = 0; y < NY; ++y) {
          cg::mbar_wait(s_full(y), (upar >> y) & 1u);
          upar ^= 1u << y;
          cg::tma_store_4d(&tmY, stg + (uint32_t)y * kStageBytes, y * 64, x0, y0, img);
          cg::bulk_commit();
          cg::bulk_wait_read<1>();
          free_prev();
          pk = 2; ptl = tl; ps = y;
        }
        // the tile's last item is freed here, not after the next store: the next tile's first sub-chunk may be waiting for
        // exactly this buffer's residual (n2 = 64: the only Y sub-chunk lives in buffer 0)
        cg::bulk_wait_read<0>();
        free_prev();
        pk = 0;
      }
      cg::bulk_wait<0>();
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tid = (int)threadIdx.x - 64;
    const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    const __half2 lo1 = __float2half2_rn(p.relu1 ? 0.0f : -INFINITY), lo2 = __float2half2_rn(p.relu2 ? 0.0f : -INFINITY);
    for (int i = tid; i < p.n2; i += 128) bias2_s[i] = __ldg(p.bias2 + i);
    uint32_t t1_uses[2] = {0, 0};
    uint32_t rpar = 0;                                 // bit b: parity of the X uses of buffer b
    uint32_t tl = 0;
    uint8_t* stg_gen = smem_raw + kStageOff;
    TraceCursor tc = ftrace_open(p, 2);
    if (warp != 2 || lane != 0) tc.base = nullptr;
    for (int t = first; t < tiles; t += step, ++tl) {
      for (int j = 0; j < NCH; ++j) {
        const int acc = j & 1;
        bias1_s[acc * 128 + tid] = __ldg(p.bias1 + j * 128 + tid);
        cg::epi_bar_sync();                            // bias chunk staged (and everybody is past chunk j - 2's reads of this half)
        cg::mbar_wait(t1_full(acc), t1_uses[acc] & 1u);
        ++t1_uses[acc];
        cg::trace_ev(tc, 5, (uint32_t)j);              // epilogue: acc1 ready
        cg::tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128);
        #pragma unroll 1
        for (int s = 0; s < 2; ++s) {
          const int b = 2 * acc + s;
          cg::mbar_wait(r_full(b), (rpar >> b) & 1u);  // residual sub-chunk landed (so the buffer is free, too)
          rpar ^= 1u << b;
          cg::trace_ev(tc, 9, (uint32_t)(2 * j + s));  // epilogue: residual sub-chunk landed
          convert_subchunk<true>(t_addr + (uint32_t)(s * 64), bias1_s + acc * 128 + s * 64, stg_gen + b * kStageBytes + row_off, sw, lo1, &tc);
          cg::trace_ev(tc, 16, (uint32_t)(2 * j + s)); // debug: converted
          if (s == 1) {
            cg::tc_fence_before();
            __syncwarp();
            if (lane == 0) cg::mbar_arrive(t1_empty(acc));
          }
          cg::fence_proxy_async_smem();                // generic-proxy writes -> visible to the TMA store and to the MMA
          __syncwarp();
          if (lane == 0) { cg::mbar_arrive(s_full(b)); cg::mbar_arrive(x_full(b)); }
          cg::trace_ev(tc, 6, (uint32_t)(2 * j + s));  // epilogue: sub-chunk staged
        }
      }
      cg::mbar_wait(t2_full, tl & 1u);
      cg::trace_ev(tc, 13, (uint32_t)t);               // epilogue: acc2 ready
      cg::tc_fence_after();
      const uint32_t t_addr2 = tmem_base + ((uint32_t)(q * 32) << 16) + 256u;
      #pragma unroll 1
      for (int y = 0; y < NY; ++y) {
        cg::mbar_wait(s_free(y), tl & 1u);             // the store thread has released the buffer (its last X sub-chunk is out)
        convert_subchunk<false>(t_addr2 + (uint32_t)(y * 64), bias2_s + y * 64, stg_gen + y * kStageBytes + row_off, sw, lo2);
        if (y == NY - 1) {
          cg::tc_fence_before();
          __syncwarp();
          if (lane == 0) cg::mbar_arrive(t2_empty);
        }
        cg::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) cg::mbar_arrive(s_full(y));
        cg::trace_ev(tc, 14, (uint32_t)y);             // epilogue: Y sub-chunk staged
      }
    }
  }
  // nobody leaves (or frees TMEM) while the partner may still read this CTA's smem / signal its barriers
  cg::tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cg::cluster_sync_all();
  if (warp == 1) {
    cg::tc_fence_after();
    if (CTAS == 2) cg::tmem_dealloc_2cta(tmem_base, 512); else cg::tmem_dealloc(tmem_base, 512);
  }
}
