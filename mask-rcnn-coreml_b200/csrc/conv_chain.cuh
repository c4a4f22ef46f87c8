// conv_chain.cuh -- a whole ResNet stage (up to CH_MAX_LAYERS convolutions) as ONE persistent launch of the
// tcgen05 implicit-GEMM convolution (conv_gemm.cuh), with tile-granular dataflow between the layers.
// EXPERIMENTAL: off unless MRCNN_CHAIN=1; bit-identical to the layer-by-layer launches, not faster yet.
//
// Why: at batch 8 the res3..res5 layers have only 64..256 M tiles for 148 SMs.  Launched one by one, every layer
// pays a partial last round (256 tiles = 1.73 rounds) and a launch head + tail of ~5 us out of its 25..40 us
// (tools/gpu/quant_probe.sh: the same layers run 1.4-1.7x faster per tile when the tile count is a multiple of 148).
// Here the CTA pairs walk ONE list of work items that spans all layers of the stage (layer-major), so there is no
// per-layer round structure and no per-layer launch; an item only waits for the TILES it reads:
//
//   done[l][m]   global counter of (layer l, 128-pixel tile m), +1 (release) by the epilogue of every N tile of that
//                tile once its TMA stores have completed.  The TMA producer of a dependent item spins (acquire) until
//                the counters it needs have reached the producing layer's N-tile count, then fences the async proxy
//                and loads: the same tile for a 1x1 convolution and for the residual, the <= 9 neighbouring tiles of
//                the same image for a 3x3 convolution.  (Image-granular counters were tried first: every image is then
//                a serial chain of layers with only 16 items per layer, and the 74 pairs starve: 22 % of the producer
//                time was spent spinning, tools/chain_stats.py.)
//
// Roles of a CTA (8 warps): TMA producer, MMA issuer (leader CTA of the pair), four epilogue warps that only compute,
// a store thread (TMA stores of the staged chunks, buffer hand-back, residual prefetch) and a signal thread (the
// gpu-scope release-increment of the counters costs ~1 us: on its own thread it delays nothing else; the store thread
// tells it which store groups have been written, cp.async.bulk.wait_group being per thread).
//
// Deadlock freedom: every role of every CTA processes its items in list order, a wait only ever refers to items
// that come EARLIER in the list, all CTAs are co-resident (grid <= cudaOccupancyMaxActiveClusters), and the only
// thread that blocks on a counter is the producer (the store thread's residual prefetch polls the producer's progress
// word and is skipped, not blocked, while the producer has not yet passed the dependency check of that item).
// Write-after-read safety of the activation buffers follows from the same per-tile chains provided the output of the
// layer in front of a 3x3 convolution is double-buffered across blocks (pipeline.cu); stages are separate
// launches because the buffer geometry changes between them.
//
// Configuration (uniform for the launch): CTA pairs (tcgen05.mma.cta_group::2, 256 x bn tiles, bn <= 256 per layer),
// 5-stage operand ring, four 16 KB output / residual staging buffers, staged TMA-store epilogue with optional
// TMA-fetched residual (the only epilogue modes ResNet bottlenecks need).
#pragma once
#include "conv_gemm.cuh"

#define CH_MAX_LAYERS 96
#define CH_STAGES 5
#define CH_NBUF 4
#define CH_MAX_TAPS 9
#define CH_STAGE_BYTES 32768      // A 128x64 fp16 (16 KB) + this CTA's half of B (<= 128 x 64 fp16 = 16 KB)
#define CH_A_BYTES 16384
#define CH_OUT_BYTES 16384
// debug statistics (clock64 totals per CTA): where each role waits
#define CH_NSTAT 16
#define CH_STAT_TOTAL 0        // kernel body
#define CH_STAT_P_FLAGS 1      // producer: spinning on dependency counters
#define CH_STAT_P_EMPTY 2      // producer: waiting for a free ring slot
#define CH_STAT_M_FULL 3       // MMA: waiting for operands
#define CH_STAT_M_TEMPTY 4     // MMA: waiting for a drained accumulator
#define CH_STAT_E_TFULL 5      // epilogue (first thread): waiting for a finished accumulator
#define CH_STAT_E_RFULL 6      // epilogue (first thread): waiting for a residual chunk
#define CH_STAT_E_BULK 7       // store thread: in cp.async.bulk.wait_group
#define CH_STAT_E_ITEMS 8      // items processed by this CTA
#define CH_STAT_E_CHUNKS 9
#define CH_STAT_E_FLUSH 10     // flushes of pending signals before an idle wait
#define CH_STAT_P_SPINS 11     // producer: items that had to spin at least once
#define CH_STAT_E_SERIAL 12    // store thread: busy storing a chunk (store issue, hand-back, publishing completed groups)
#define CH_STAT_E_BARRIER 13   // epilogue (first thread): waiting for a staging buffer to be handed back
#define CH_STAT_E_SIGNAL 14    // signal thread: inside the release-increments
#define CH_THREADS 256             // producer, MMA, 4 epilogue warps, store warp, signal warp
#define CH_SIGQ 16                 // completion-signal queue entries (store thread -> signal thread)
#define CH_SMEM_BYTES (1024 + CH_STAGES * CH_STAGE_BYTES + CH_NBUF * CH_OUT_BYTES + 512)

struct ChainLayer {               // 64 bytes, lives in the kernel parameter (constant bank)
  uint32_t reserved;
  uint16_t tiles_x, tiles_y, tiles_n, tiles_per_img;
  uint16_t cout, bn, cin_chunks;
  uint8_t ntaps, stride, tw, th, relu, res;
  uint8_t nbhd_a;                 // 1: the input dependency covers the 3x3 tile neighbourhood (3x3 convolution), 0: the same tile
  uint8_t pad1;
  int16_t dep_a, dep_r;           // chain-local index of the layer that wrote this layer's input / residual, -1: before the launch
  uint16_t need_a, need_r;        // completions per tile that layer produces (its number of N tiles)
  int8_t tap_dx[CH_MAX_TAPS], tap_dy[CH_MAX_TAPS];
  uint8_t pad2[4];
  const float* bias;
};
static_assert(sizeof(ChainLayer) == 64, "ChainLayer must stay 64 bytes");

// The work list is a sequence of segments.  A segment is either a run of consecutive items of one layer, or two such
// runs (of different layers, over different images) zipped together: item i of a zipped segment belongs to stream X
// iff floor((i+1)*nx/n) > floor(i*nx/n), n = nx + ny.  The host uses zipped segments to software-pipeline two halves of
// the batch two layers apart, so that a pair alternates between MMA-heavy items (3x3 convolutions) and epilogue-heavy
// ones (the 1x1 expansions with their residual) instead of running each kind back to back.
#define CH_MAX_SEGS 224
struct ChainSeg {                 // 24 bytes
  uint32_t item_end;              // cumulative number of pair items up to and including this segment
  int16_t lx, ly;                 // layers of the two streams (ly < 0: single stream)
  uint32_t x0, y0;                // first item of each stream inside its layer
  uint32_t nx, ny;                // items of each stream in this segment
};

struct ChainParams {
  int n_layers, n_img;
  int n_segs;
  uint32_t total_items;
  int flag_stride;                // counters per layer (>= number of 128-pixel tiles of any layer)
  uint32_t* done;                 // [n_layers][flag_stride], zeroed before the launch
  const CUtensorMap* maps;        // [n_layers][4] = A, B, C, R (global memory, 64-byte aligned)
  unsigned long long* stats;      // debug: [gridDim.x][CH_NSTAT] clock totals per CTA (nullptr in production), see CH_STAT_*
  ChainLayer L[CH_MAX_LAYERS];
  ChainSeg S[CH_MAX_SEGS];
};

#ifdef CONV_CHAIN_KERNEL      // the kernel itself is compiled in dense.cu only
namespace cg {

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr) {
  uint32_t v; asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// position of one CTA in the chain's item list
struct ChainPos {
  uint32_t g;        // global item index
  uint32_t sbase;    // first item of segment s
  int s;             // segment
  int l;             // layer of item g (set by seek)
  uint32_t w;        // index of item g inside its layer (set by seek)
  __device__ __forceinline__ void init(uint32_t g0) { g = g0; sbase = 0; s = 0; l = 0; w = 0; }
  __device__ __forceinline__ void seek(const ChainParams& cp) {
    while (s < cp.n_segs - 1 && g >= cp.S[s].item_end) { sbase = cp.S[s].item_end; ++s; }
    const ChainSeg& G = cp.S[s];
    const uint32_t i = g - sbase;
    if (G.ly < 0) { l = G.lx; w = G.x0 + i; return; }
    const uint32_t n = G.nx + G.ny;
    const uint32_t xa = (uint32_t)(((unsigned long long)i * G.nx) / n);             // X items before item i
    const uint32_t xb = (uint32_t)(((unsigned long long)(i + 1u) * G.nx) / n);
    if (xb > xa) { l = G.lx; w = G.x0 + xa; } else { l = G.ly; w = G.y0 + (i - xa); }
  }
  __device__ __forceinline__ void coords(const ChainParams& cp, int crank, int& n0, int& x0, int& y0, int& img) const {
    int mt; coords(cp, crank, n0, x0, y0, img, mt);
  }
  __device__ __forceinline__ void coords(const ChainParams& cp, int crank, int& n0, int& x0, int& y0, int& img, int& mt) const {
    const ChainLayer& L = cp.L[l];
    const int nt = (int)(w % L.tiles_n);
    mt = (int)(w / L.tiles_n) * 2 + crank;
    n0 = nt * L.bn;
    x0 = (mt % L.tiles_x) * L.tw;
    y0 = ((mt / L.tiles_x) % L.tiles_y) * L.th;
    img = mt / L.tiles_per_img;      // == n_img for the phantom tile of an odd pair: TMA clips / zero-fills it
  }
};

}  // namespace cg

__global__ void __launch_bounds__(CH_THREADS, 1)
conv_chain_kernel(const __grid_constant__ ChainParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (cg::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_base = smem_base + CH_STAGES * CH_STAGE_BYTES;
  const uint32_t bar_base = out_base + CH_NBUF * CH_OUT_BYTES;
  // barrier block (320 B): full[8] @0, empty[8] @64, tmem_full[2] @128, tmem_empty[2] @144, TMEM base slot @160, deps_ok @164,
  // res_full[4] @168, store_full[4] @200 (4 arrivals: one per epilogue warp), store_free[4] @232 (store thread),
  // signal queue @264: written (tail) @264, completed store group @268, finished @272, then CH_SIGQ x u64 counter address
  // @280 and CH_SIGQ x u32 store-group sequence @408
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto tfull_bar = [&](int a) { return bar_base + 128u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 144u + 8u * a; };
  const uint32_t tmem_slot = bar_base + 160u;
  auto rfull_bar = [&](int b) { return bar_base + 168u + 8u * b; };
  auto sfull_bar = [&](int b) { return bar_base + 200u + 8u * b; };
  auto sfree_bar = [&](int b) { return bar_base + 232u + 8u * b; };
  const uint32_t deps_ok_addr = bar_base + 164u;     // number of this CTA's items whose dependencies the producer has seen satisfied
  uint8_t* smem_gen = smem_raw + (smem_base - cg::smem_u32(smem_raw));
  uint8_t* out_gen = smem_gen + CH_STAGES * CH_STAGE_BYTES;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(out_gen + CH_NBUF * CH_OUT_BYTES + 160u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cg::cluster_ctarank();

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < CH_STAGES; ++s) { cg::mbar_init(full_bar(s), 1); cg::mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { cg::mbar_init(tfull_bar(a), 1); cg::mbar_init(tempty_bar(a), 8); }
    for (int b = 0; b < CH_NBUF; ++b) { cg::mbar_init(rfull_bar(b), 1); cg::mbar_init(sfull_bar(b), 4); cg::mbar_init(sfree_bar(b), 1); }
    *reinterpret_cast<volatile uint32_t*>(out_gen + CH_NBUF * CH_OUT_BYTES + 164u) = 0u;
    for (int k = 0; k < 4; ++k) *reinterpret_cast<volatile uint32_t*>(out_gen + CH_NBUF * CH_OUT_BYTES + 264u + 4u * k) = 0u;   // signal queue: tail, completed, finished, head
    cg::fence_barrier_init();
  }
  if (warp == 1) cg::tmem_alloc_2cta(tmem_slot, 512);
  cg::tc_fence_before();
  __syncthreads();
  cg::cluster_sync_all();
  cg::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t pidx = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  unsigned long long* st = cp.stats ? cp.stats + (size_t)blockIdx.x * CH_NSTAT : nullptr;
  const long long t_begin = clock64();
#define CH_TIMED(slot, stmt) do { if (st) { const long long _t = clock64(); stmt; acc_##slot += (unsigned long long)(clock64() - _t); } else { stmt; } } while (0)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      cg::ChainPos pos; pos.init(pidx);
      int cur_l = -1;
      uint32_t ord = 0;
      // Dependency counters are polled one item ahead: the loads for item k+1 are issued before the TMA copies of item
      // k, so their L2 round trip is hidden behind that item's work.  The counters of the two most recently satisfied
      // (layer, image) pairs are remembered, so a run of items of the same image polls once.
      unsigned long long acc_pf = 0, acc_pe = 0, n_spin = 0;
      const uint32_t* nf_a = nullptr; const uint32_t* nf_r = nullptr;     // flags polled for the current item
      uint32_t nv_a = 0, nv_r = 0;                                       // ... and the values that came back
      const uint32_t* ok_a = nullptr; const uint32_t* ok_r = nullptr;     // flags known to be satisfied
      // same-tile dependencies of item q: counter addresses (nullptr: none / neighbourhood dependency, handled at wait time)
      auto dep_flags = [&](const cg::ChainPos& q, const uint32_t*& fa, const uint32_t*& fr) {
        fa = nullptr; fr = nullptr;
        if (q.g >= cp.total_items) return;
        const ChainLayer& Q = cp.L[q.l];
        int n0, x0, y0, img, mt;
        q.coords(cp, (int)crank, n0, x0, y0, img, mt);
        if (img >= cp.n_img) return;
        if (Q.dep_a >= 0 && !Q.nbhd_a) fa = cp.done + (size_t)Q.dep_a * cp.flag_stride + mt;
        if (Q.dep_r >= 0) fr = cp.done + (size_t)Q.dep_r * cp.flag_stride + mt;
      };
      {
        cg::ChainPos q = pos; if (q.g < cp.total_items) q.seek(cp);
        dep_flags(q, nf_a, nf_r);
        if (nf_a) nv_a = cg::ld_acquire_gpu(nf_a);
        if (nf_r) nv_r = cg::ld_acquire_gpu(nf_r);
      }
      for (; pos.g < cp.total_items; pos.g += npairs, ++ord) {
        pos.seek(cp);
        const ChainLayer& L = cp.L[pos.l];
        const CUtensorMap* tmA = cp.maps + 4 * pos.l;
        const CUtensorMap* tmB = tmA + 1;
        if (pos.s != cur_l) {                               // new segment: fetch the descriptors of the next one ahead of time
          if (cur_l < 0) { cg::prefetch_tmap(tmA); cg::prefetch_tmap(tmB); }
          if (pos.s + 1 < cp.n_segs) {
            const ChainSeg& NS = cp.S[pos.s + 1];
            cg::prefetch_tmap(cp.maps + 4 * NS.lx); cg::prefetch_tmap(cp.maps + 4 * NS.lx + 1);
            if (NS.ly >= 0) { cg::prefetch_tmap(cp.maps + 4 * NS.ly); cg::prefetch_tmap(cp.maps + 4 * NS.ly + 1); }
          }
          cur_l = pos.s;
        }
        int n0, px0, py0, img, mt;
        pos.coords(cp, (int)crank, n0, px0, py0, img, mt);
        // this item's dependencies (same-tile counters were polled during the previous item; spin only if they were
        // not yet satisfied then)
        {
          const long long t0 = st ? clock64() : 0;
          bool spun = false;
          if (nf_a && nf_a != ok_a) {
            while (nv_a < (uint32_t)L.need_a) { spun = true; __nanosleep(40); nv_a = cg::ld_acquire_gpu(nf_a); }
            ok_a = nf_a;
          }
          if (nf_r && nf_r != ok_r) {
            while (nv_r < (uint32_t)L.need_r) { spun = true; __nanosleep(40); nv_r = cg::ld_acquire_gpu(nf_r); }
            ok_r = nf_r;
          }
          if (L.dep_a >= 0 && L.nbhd_a && img < cp.n_img) {
            // 3x3 convolution: the (up to) 9 tiles around this one, all polled at once
            const int tx = mt % L.tiles_x, ty = (mt / L.tiles_x) % L.tiles_y;
            const uint32_t* fb = cp.done + (size_t)L.dep_a * cp.flag_stride + (mt - ty * L.tiles_x - tx);   // tile (0,0) of this image
            const int xl = tx > 0 ? tx - 1 : tx, xh = tx + 1 < L.tiles_x ? tx + 1 : tx;
            const int yl = ty > 0 ? ty - 1 : ty, yh = ty + 1 < L.tiles_y ? ty + 1 : ty;
            for (;;) {
              uint32_t v[9];
              #pragma unroll
              for (int k = 0; k < 9; ++k) {
                const int yy = yl + k / 3, xx = xl + k % 3;
                v[k] = (yy <= yh && xx <= xh) ? cg::ld_acquire_gpu(fb + yy * L.tiles_x + xx) : 0xffffffffu;
              }
              uint32_t mn = 0xffffffffu;
              #pragma unroll
              for (int k = 0; k < 9; ++k) mn = v[k] < mn ? v[k] : mn;
              if (mn >= (uint32_t)L.need_a) break;
              spun = true;
              __nanosleep(40);
            }
          }
          if (st) { acc_pf += (unsigned long long)(clock64() - t0); n_spin += spun ? 1 : 0; }
        }
        cg::fence_proxy_async_all();                       // acquired generic-proxy view -> the TMA (async proxy) loads below
        cg::st_release_cta_shared(deps_ok_addr, ord + 1);  // lets the epilogue prefetch this item's residual
        // poll for the next item now; the values are consumed at the top of the next iteration
        {
          cg::ChainPos q = pos; q.g += npairs;
          if (q.g < cp.total_items) q.seek(cp);
          dep_flags(q, nf_a, nf_r);
          if (nf_a && nf_a != ok_a) nv_a = cg::ld_acquire_gpu(nf_a);
          if (nf_r && nf_r != ok_r) nv_r = cg::ld_acquire_gpu(nf_r);
        }
        const int x0 = px0 * L.stride, y0 = py0 * L.stride;
        const int half = L.bn >> 1;
        const int nb0 = n0 + (int)crank * half;
        const uint32_t tx = 2u * (uint32_t)(CH_A_BYTES + half * 128);
        const int ntaps = L.ntaps, cin_chunks = L.cin_chunks, cin = L.cin_chunks * CG_BK;
        for (int t = 0; t < ntaps; ++t) {
          const int xi = x0 + L.tap_dx[t], yi = y0 + L.tap_dy[t];
          for (int cc = 0; cc < cin_chunks; ++cc) {
            CH_TIMED(pe, cg::mbar_wait(empty_bar(stage), phase ^ 1u));
            const uint32_t a_dst = smem_base + stage * CH_STAGE_BYTES;
            if (crank == 0) cg::mbar_expect_tx(full_bar(stage), tx);
            const uint32_t lbar = cg::mapa_rank(full_bar(stage), 0);
            cg::tma_load_4d_2cta(a_dst, tmA, lbar, cc * CG_BK, xi, yi, img);
            cg::tma_load_2d_2cta(a_dst + CH_A_BYTES, tmB, lbar, t * cin + cc * CG_BK, nb0);
            if (++stage == CH_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (st) { st[CH_STAT_P_FLAGS] = acc_pf; st[CH_STAT_P_EMPTY] = acc_pe; st[CH_STAT_P_SPINS] = n_spin; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (lane == 0 && crank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      unsigned long long acc_mf = 0, acc_mt = 0;
      cg::ChainPos pos; pos.init(pidx);
      for (; pos.g < cp.total_items; pos.g += npairs) {
        pos.seek(cp);
        const ChainLayer& L = cp.L[pos.l];
        const int nkb = (int)L.ntaps * (int)L.cin_chunks;
        const uint32_t idesc = cg::make_idesc_f16(256, (int)L.bn);
        CH_TIMED(mt, cg::mbar_wait(tempty_bar(acc), acc_phase ^ 1u));
        cg::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int kb = 0; kb < nkb; ++kb) {
          CH_TIMED(mf, cg::mbar_wait(full_bar(stage), phase));
          cg::tc_fence_after();
          const uint32_t a_addr = smem_base + stage * CH_STAGE_BYTES;
          const uint64_t adesc = cg::make_sw128_desc(a_addr);
          const uint64_t bdesc = cg::make_sw128_desc(a_addr + CH_A_BYTES);
          #pragma unroll
          for (int k = 0; k < CG_BK / 16; ++k)
            cg::umma_f16_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          cg::umma_commit_2cta(empty_bar(stage));
          if (++stage == CH_STAGES) { stage = 0; phase ^= 1u; }
        }
        cg::umma_commit_2cta(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (st) { st[CH_STAT_M_FULL] = acc_mf; st[CH_STAT_M_TEMPTY] = acc_mt; }
    }
  } else if (warp < 6) {
    // ===================== epilogue warps (2..5): TMEM -> (+bias, +residual, ReLU) -> fp16 -> staging buffer =====================
    // No block barrier and no serial section here: buffers are handed to / taken back from the store thread (warp 6)
    // through mbarriers (store_full / store_free), residual chunks arrive on res_full.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool t0 = (threadIdx.x == 64);      // statistics only
    if (!t0) st = nullptr;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    uint32_t cc = 0;                 // chunk counter of this CTA (staging buffer = cc & 3)
    uint32_t rphase = 0;             // bit b = parity to wait for on res_full[b]
    int acc = 0; uint32_t acc_phase = 0;
    unsigned long long acc_et = 0, acc_er = 0, acc_ef = 0, n_items = 0;
    cg::ChainPos pos; pos.init(pidx);
    for (; pos.g < cp.total_items; pos.g += npairs) {
      pos.seek(cp);
      const ChainLayer& L = cp.L[pos.l];
      int n0, x0, y0, img;
      pos.coords(cp, (int)crank, n0, x0, y0, img);
      const int nchunks = L.bn >> 6;
      const bool res = L.res != 0;
      const float lo = L.relu ? 0.0f : -INFINITY;
      // bias straight from global memory (a layer's bias is <= 8 KB: L1 resident after the first tile; the loads are
      // issued under the TMEM load's latency)
      const float* bias_row = L.bias + n0;
      CH_TIMED(et, cg::mbar_wait(tfull_bar(acc), acc_phase));
      ++n_items;
      cg::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      #pragma unroll 1
      for (int c = 0; c < nchunks; ++c, ++cc) {
        const uint32_t buf = cc & 3u;
        uint8_t* srow = out_gen + buf * CH_OUT_BYTES + row_off;
        if (res) {
          // the residual chunk has landed in this buffer (which also means the store that last used it has been read)
          CH_TIMED(er, cg::mbar_wait(rfull_bar((int)buf), (rphase >> buf) & 1u));
          rphase ^= (1u << buf);
        } else {
          // the store that last used this buffer (chunk cc - 4) has been read; passes at once for the first four chunks
          CH_TIMED(ef, cg::mbar_wait(sfree_bar((int)buf), ((cc >> 2) & 1u) ^ 1u));
        }
        #pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          cg::tmem_ld32(t_addr + (uint32_t)(c * 64 + hh * 32), v);
          float4 bq[8];
          uint4 rq[4];
          #pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4* bp = reinterpret_cast<const float4*>(bias_row + (c * 64 + hh * 32 + g * 8));
            bq[2 * g] = __ldg(bp); bq[2 * g + 1] = __ldg(bp + 1);
            rq[g] = make_uint4(0u, 0u, 0u, 0u);
            if (res) rq[g] = *reinterpret_cast<const uint4*>(srow + ((((uint32_t)(hh * 4 + g)) ^ sw) << 4));
          }
          cg::tmem_ld_wait();
          #pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
            #pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g * 8 + e]);
            f[0] += bq[2 * g].x; f[1] += bq[2 * g].y; f[2] += bq[2 * g].z; f[3] += bq[2 * g].w;
            f[4] += bq[2 * g + 1].x; f[5] += bq[2 * g + 1].y; f[6] += bq[2 * g + 1].z; f[7] += bq[2 * g + 1].w;
            const __half2* rh = reinterpret_cast<const __half2*>(&rq[g]);     // zeros when there is no residual
            #pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 r2 = __half22float2(rh[e]); f[2 * e] += r2.x; f[2 * e + 1] += r2.y; }
            uint4 ov;
            __half2* oh = reinterpret_cast<__half2*>(&ov);
            #pragma unroll
            for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(fmaxf(f[2 * e], lo), fmaxf(f[2 * e + 1], lo));
            *reinterpret_cast<uint4*>(srow + ((((uint32_t)(hh * 4 + g)) ^ sw) << 4)) = ov;
          }
        }
        if (c == nchunks - 1) {                // accumulator fully read: hand the TMEM stage back to the MMA warp
          cg::tc_fence_before();
          __syncwarp();
          if (lane == 0) cg::mbar_arrive_cluster(cg::mapa_rank(tempty_bar(acc), 0));
        }
        cg::fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the bulk-copy engine
        __syncwarp();
        if (lane == 0) cg::mbar_arrive(sfull_bar((int)buf));     // 4 arrivals (one per warp) = chunk staged
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (st) {
      st[CH_STAT_E_TFULL] = acc_et; st[CH_STAT_E_RFULL] = acc_er; st[CH_STAT_E_BARRIER] = acc_ef;
      st[CH_STAT_E_ITEMS] = n_items; st[CH_STAT_E_CHUNKS] = cc;
      st[CH_STAT_TOTAL] = (unsigned long long)(clock64() - t_begin);
    }
  } else if (warp == 6) {
    // ===================== store thread (warp 6, one lane): TMA stores, residual prefetch, completion signals =====================
    if (lane == 0) {
      unsigned long long acc_ss = 0, acc_eb = 0, n_flush = 0;
      uint32_t stored = 0;            // chunks stored so far == store groups committed
      uint32_t freed = 0;             // stores whose staging buffer has been handed back (store_free arrived)
      // walk of the chunks to store
      cg::ChainPos pos; pos.init(pidx);
      int sc = 0;                     // next chunk of item pos
      // walk of the residual chunks to fetch (runs ahead)
      cg::ChainPos pf; pf.init(pidx);
      uint32_t pf_ord = 0, pf_j = 0; int pf_c = 0;
      // completion signals: pushed here as (counter address, store-group sequence), fired by the signal thread (warp 7)
      // once this thread has seen that group written (cp.async.bulk.wait_group is per thread) and published `completed`
      const uint32_t q_tail_addr = bar_base + 264u, q_done_addr = bar_base + 268u, q_fin_addr = bar_base + 272u;
      const uint32_t q_addr = bar_base + 280u, q_seq = bar_base + 408u;
      uint32_t q_tail = 0, published = 0;
      auto publish = [&](uint32_t complete_upto) {
        if ((int32_t)(complete_upto - published) > 0) { published = complete_upto; cg::st_release_cta_shared(q_done_addr, published); }
      };
      // residual loads for every chunk whose buffer is free (chunk j needs store j-4 read: j <= freed + 3) and whose item
      // the producer has cleared (deps_ok: its dependencies hold).  This thread never blocks, so a chunk the epilogue
      // warps already wait for is fetched as soon as the producer's progress word allows it.
      auto prefetch_residuals = [&]() -> bool {
        bool any = false;
        while (pf.g < cp.total_items) {
          pf.seek(cp);
          const ChainLayer& PL = cp.L[pf.l];
          const int nch = PL.bn >> 6;
          const uint32_t j = pf_j + (uint32_t)pf_c;
          if (j > freed + 3u) break;
          if (!PL.res) { pf_j += (uint32_t)nch; pf.g += npairs; ++pf_ord; pf_c = 0; continue; }
          if (cg::ld_acquire_cta_shared(deps_ok_addr) <= pf_ord) break;
          int n0, x0, y0, img;
          pf.coords(cp, (int)crank, n0, x0, y0, img);
          const uint32_t b = j & 3u;
          cg::mbar_expect_tx(rfull_bar((int)b), (uint32_t)CH_OUT_BYTES);
          cg::tma_load_4d(out_base + b * (uint32_t)CH_OUT_BYTES, cp.maps + 4 * pf.l + 3, rfull_bar((int)b), n0 + pf_c * 64, x0, y0, img);
          any = true;
          if (++pf_c == nch) { pf_c = 0; pf_j += (uint32_t)nch; pf.g += npairs; ++pf_ord; }
        }
        return any;
      };
      int cur_l = -1;
      int n0 = 0, x0 = 0, y0 = 0, img = 0, mt = 0, nchunks = 1;
      bool have_item = false;
      while (pos.g < cp.total_items) {
        if (!have_item) {
          pos.seek(cp);
          if (pos.s != cur_l) {
            if (cur_l < 0) { cg::prefetch_tmap(cp.maps + 4 * pos.l + 2); cg::prefetch_tmap(cp.maps + 4 * pos.l + 3); }
            if (pos.s + 1 < cp.n_segs) {
              const ChainSeg& NS = cp.S[pos.s + 1];
              cg::prefetch_tmap(cp.maps + 4 * NS.lx + 2); cg::prefetch_tmap(cp.maps + 4 * NS.lx + 3);
              if (NS.ly >= 0) { cg::prefetch_tmap(cp.maps + 4 * NS.ly + 2); cg::prefetch_tmap(cp.maps + 4 * NS.ly + 3); }
            }
            cur_l = pos.s;
          }
          pos.coords(cp, (int)crank, n0, x0, y0, img, mt);
          nchunks = cp.L[pos.l].bn >> 6;
          sc = 0;
          have_item = true;
        }
        bool progress = false;
        // (1) residual fetches
        progress |= prefetch_residuals();
        // (2) store the next chunk if the epilogue warps have staged it
        const uint32_t buf = stored & 3u;
        if (cg::mbar_test(sfull_bar((int)buf), (stored >> 2) & 1u)) {
          const long long tser = st ? clock64() : 0;
          cg::tma_store_4d(cp.maps + 4 * pos.l + 2, out_base + buf * (uint32_t)CH_OUT_BYTES, n0 + sc * 64, x0, y0, img);
          cg::bulk_commit();
          ++stored;
          if (sc == nchunks - 1 && img < cp.n_img) {
            // item complete once this store group has been written: queue its (layer, tile) completion signal
            const uint32_t slot = q_tail & (CH_SIGQ - 1);
            const unsigned long long addr = (unsigned long long)(cp.done + (size_t)pos.l * cp.flag_stride + mt);
            while (q_tail - cg::ld_acquire_cta_shared(bar_base + 276u) >= CH_SIGQ) __nanosleep(20);   // queue full (head @276)
            asm volatile("st.shared::cta.u64 [%0], %1;" ::"r"(q_addr + 8u * slot), "l"(addr) : "memory");
            asm volatile("st.shared::cta.u32 [%0], %1;" ::"r"(q_seq + 4u * slot), "r"(stored) : "memory");
            ++q_tail;
            cg::st_release_cta_shared(q_tail_addr, q_tail);
          }
          CH_TIMED(eb, cg::bulk_wait_read<1>());            // every store but the newest has been read: hand those buffers back
          while (freed + 1u < stored) { cg::mbar_arrive(sfree_bar((int)(freed & 3u))); ++freed; }
          CH_TIMED(eb, cg::bulk_wait<1>());                 // every group but the newest has been written
          publish(stored - 1u);
          if (++sc == nchunks) { have_item = false; pos.g += npairs; }
          if (st) acc_ss += (unsigned long long)(clock64() - tser);
          progress = true;
        }
        if (!progress) {
          if (published != stored || freed < stored) {
            // idle: do not sit on buffers or completion signals
            CH_TIMED(eb, cg::bulk_wait<0>());
            while (freed < stored) { cg::mbar_arrive(sfree_bar((int)(freed & 3u))); ++freed; }
            publish(stored);
            ++n_flush;
          } else {
            __nanosleep(20);
          }
        }
      }
      cg::bulk_wait<0>();
      while (freed < stored) { cg::mbar_arrive(sfree_bar((int)(freed & 3u))); ++freed; }
      publish(stored);
      cg::st_release_cta_shared(q_fin_addr, 1u);
      if (st) { st[CH_STAT_E_SERIAL] = acc_ss; st[CH_STAT_E_BULK] = acc_eb; st[CH_STAT_E_FLUSH] = n_flush; }
    }
  } else {
    // ===================== signal thread (warp 7, one lane): release-increments of the (layer, tile) counters =====================
    // A gpu-scope release costs ~1 us; on its own thread it delays neither the stores nor the residual fetches.  The
    // release is cumulative: this thread acquired `completed` from the store thread, which observed the writes complete.
    if (lane == 0) {
      unsigned long long acc_esig = 0;
      const uint32_t q_tail_addr = bar_base + 264u, q_done_addr = bar_base + 268u, q_fin_addr = bar_base + 272u, q_head_addr = bar_base + 276u;
      const uint32_t q_addr = bar_base + 280u, q_seq = bar_base + 408u;
      uint32_t head = 0;
      for (;;) {
        const uint32_t fin = cg::ld_acquire_cta_shared(q_fin_addr);
        const uint32_t tail = cg::ld_acquire_cta_shared(q_tail_addr);
        const uint32_t done = cg::ld_acquire_cta_shared(q_done_addr);
        bool fired = false;
        while (head != tail) {
          const uint32_t slot = head & (CH_SIGQ - 1);
          uint32_t seq; unsigned long long addr;
          asm volatile("ld.shared::cta.u32 %0, [%1];" : "=r"(seq) : "r"(q_seq + 4u * slot) : "memory");
          if ((int32_t)(done - seq) < 0) break;
          asm volatile("ld.shared::cta.u64 %0, [%1];" : "=l"(addr) : "r"(q_addr + 8u * slot) : "memory");
          const long long tsig = st ? clock64() : 0;
          cg::red_release_gpu_add(reinterpret_cast<uint32_t*>(addr), 1u);
          if (st) acc_esig += (unsigned long long)(clock64() - tsig);
          ++head;
          cg::st_release_cta_shared(q_head_addr, head);
          fired = true;
        }
        if (fin && head == tail) break;       // `fin` was read before `tail`: nothing can have been pushed after that tail
        if (!fired) __nanosleep(20);
      }
      if (st) st[CH_STAT_E_SIGNAL] = acc_esig;
    }
  }
  cg::tc_fence_before();
  __syncthreads();
  cg::cluster_sync_all();
  if (warp == 1) {
    cg::tc_fence_after();
    cg::tmem_dealloc_2cta(tmem_base, 512);
  }
}
#endif  // CONV_CHAIN_KERNEL
