// proposal.cu -- ProposalLayer.evaluate (ProposalLayer.swift:103-195) on sm_100a.
//
// The reference argsorts all N object scores (vDSP_vsorti, Utils.swift:56-66),
// keeps the first n = min(N, preNMS), gathers + decodes + clips them and runs a
// sequential greedy NMS.  Here, per image (all images of the batch in one grid):
//   1. radix select (13-bit digits) of the n largest COMPOSITE keys
//        comp = (order-preserving key of score) << idx_bits | (idx_mask - anchor)
//      which encodes the intended tie-break (equal score -> lower anchor first),
//      so exactly n keys are >= the selected threshold, ties included;
//   2. compaction of those n keys (unordered; they are unique);
//   3. one CTA per image: in-smem bitonic sort (descending) of the n keys, then
//      fused gather of deltas/anchors + x std + decode + clip (BoxUtils.swift);
//   4. tiled fp64 IoU bitmask + chunked sequential resolve (nms.cuh);
//   5. rois written in keep order, zero padded (ProposalLayer.swift:181-192).
// The strided slice of the object probability (Utils.swift:17-26) is folded into
// the key computation (float2 load, .y).
#include "common.cuh"
#include "nms.cuh"

#define RADIX_BITS 13
#define RADIX_BINS (1 << RADIX_BITS)
#define SEL_THREADS 1024
#define SEL_ITEMS 8     // elements per thread per block in the histogram passes

struct SelState {
  unsigned long long prefix;   // selected high bits so far (right aligned)
  unsigned long long thr;      // final threshold: select comp >= thr
  uint32_t need;               // how many to take inside the current prefix bucket
  uint32_t ticket;             // block arrival counter
  uint32_t cand_count;         // compaction cursor
  uint32_t pad;
};

__device__ __forceinline__ unsigned long long make_comp(float s, uint32_t idx, int idx_bits) {
  unsigned long long k = score_key(s);
  unsigned long long m = (1ull << idx_bits) - 1ull;
  return (k << idx_bits) | (m - (unsigned long long)idx);
}

__global__ void sel_init_kernel(SelState* st, uint32_t* hist, int n_take) {
  const int img = blockIdx.x;
  for (int i = threadIdx.x; i < RADIX_BINS; i += blockDim.x) hist[(size_t)img * RADIX_BINS + i] = 0u;
  if (threadIdx.x == 0) {
    st[img].prefix = 0ull; st[img].thr = 0ull; st[img].need = (uint32_t)n_take;
    st[img].ticket = 0u; st[img].cand_count = 0u;
  }
}

// One radix pass. total_bits = 32 + idx_bits; pass p looks at digit bits
// [shift, shift+bits) of comp, among elements whose higher bits equal st.prefix.
__global__ void __launch_bounds__(SEL_THREADS)
sel_hist_kernel(const float2* __restrict__ probs, int64_t N, int idx_bits, int shift, int bits,
                int first_pass, int last_pass, SelState* __restrict__ st, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[RADIX_BINS];
  __shared__ uint32_t s_part[SEL_THREADS / 32];
  __shared__ uint32_t s_last;
  const int img = blockIdx.y;
  const int tid = threadIdx.x;
  for (int i = tid; i < RADIX_BINS; i += SEL_THREADS) sh[i] = 0u;
  __syncthreads();
  const unsigned long long prefix = st[img].prefix;
  const uint32_t need = st[img].need;   // read before any block of this pass can update it
  const float2* p = probs + (size_t)img * N;
  const uint32_t dmask = (1u << bits) - 1u;
  const int64_t base = (int64_t)blockIdx.x * (SEL_THREADS * SEL_ITEMS);
  #pragma unroll
  for (int it = 0; it < SEL_ITEMS; ++it) {
    int64_t i = base + (int64_t)it * SEL_THREADS + tid;
    if (i < N) {
      unsigned long long comp = make_comp(p[i].y, (uint32_t)i, idx_bits);
      bool match = first_pass ? true : ((comp >> (shift + bits)) == prefix);
      if (match) atomicAdd(&sh[(uint32_t)(comp >> shift) & dmask], 1u);
    }
  }
  __syncthreads();
  uint32_t* gh = hist + (size_t)img * RADIX_BINS;
  for (int i = tid; i < RADIX_BINS; i += SEL_THREADS) {
    uint32_t v = sh[i];
    if (v) atomicAdd(&gh[i], v);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    uint32_t t = atomicAdd(&st[img].ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // Last block of this image: find the bin (scanning from the top) where the
  // cumulative count reaches `need`.
  const int per = RADIX_BINS / SEL_THREADS;          // 8 bins per thread
  // thread t owns bins [hi - per*t - per + 1 .. hi - per*t], i.e. descending order
  uint32_t local[per];
  uint32_t sum = 0;
  #pragma unroll
  for (int k = 0; k < per; ++k) {
    int bin = RADIX_BINS - 1 - (tid * per + k);
    local[k] = __ldcg(&gh[bin]);
    sum += local[k];
  }
  // exclusive scan of `sum` over threads (descending bins)
  uint32_t incl = sum;
  const int lane = tid & 31, wid = tid >> 5;
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_part[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t v = s_part[lane];
    uint32_t iv = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += u;
    }
    s_part[lane] = iv - v;   // exclusive
  }
  __syncthreads();
  uint32_t excl = s_part[wid] + incl - sum;
  // the selected bin is the one where excl_before < need <= excl_before + count
  uint32_t run = excl;
  #pragma unroll
  for (int k = 0; k < per; ++k) {
    uint32_t cnt = local[k];
    if (run < need && need <= run + cnt) {
      int bin = RADIX_BINS - 1 - (tid * per + k);
      unsigned long long np = first_pass ? (unsigned long long)bin
                                         : ((prefix << bits) | (unsigned long long)bin);
      st[img].prefix = np;
      st[img].need = need - run;
      if (last_pass) st[img].thr = np;   // shift == 0 on the last pass
    }
    run += cnt;
  }
  __syncthreads();
  for (int i = tid; i < RADIX_BINS; i += SEL_THREADS) gh[i] = 0u;
  if (tid == 0) st[img].ticket = 0u;
}

// Compaction: every comp >= thr goes to cand[img][slot]; exactly n_take of them.
__global__ void __launch_bounds__(SEL_THREADS)
sel_compact_kernel(const float2* __restrict__ probs, int64_t N, int idx_bits,
                   SelState* __restrict__ st, unsigned long long* __restrict__ cand, int cand_stride) {
  const int img = blockIdx.y;
  const unsigned long long thr = st[img].thr;
  const float2* p = probs + (size_t)img * N;
  const int64_t base = (int64_t)blockIdx.x * (SEL_THREADS * SEL_ITEMS);
  const int lane = threadIdx.x & 31;
  #pragma unroll
  for (int it = 0; it < SEL_ITEMS; ++it) {
    int64_t i = base + (int64_t)it * SEL_THREADS + threadIdx.x;
    unsigned long long comp = 0ull;
    bool take = false;
    if (i < N) { comp = make_comp(p[i].y, (uint32_t)i, idx_bits); take = comp >= thr; }
    unsigned int bal = __ballot_sync(0xffffffffu, take);
    if (bal) {
      int leader = __ffs(bal) - 1;
      uint32_t slot0 = 0;
      if (lane == leader) slot0 = atomicAdd(&st[img].cand_count, (uint32_t)__popc(bal));
      slot0 = __shfl_sync(0xffffffffu, slot0, leader);
      if (take) {
        uint32_t slot = slot0 + __popc(bal & ((1u << lane) - 1u));
        if (slot < (uint32_t)cand_stride) cand[(size_t)img * cand_stride + slot] = comp;
      }
    }
  }
}

// One CTA per image: bitonic sort (descending) of SORT_N composite keys, then the fused gather + decode + clip of the n
// best.  Register-blocked: thread t holds the E = SORT_N / 1024 keys t*E .. t*E+E-1.  A compare-exchange step of stride j
// pairs element e with e ^ j: for j < E the partner is in the thread's own registers, for j < 32 E in another lane of
// the warp (shuffle), only for larger strides in another warp (shared memory + block barrier): 15 of the 91 steps of an
// 8192-key sort instead of all of them (88 -> ~25 us; the reference's vDSP_vsorti of 261,888 scores is its "avg of
// 45 ms" bottleneck, ProposalLayer.swift:131).
template <int E>
__device__ __forceinline__ void bitonic_cx(unsigned long long& mine, unsigned long long other, bool keep_max) {
  const bool gt = mine > other;
  mine = (gt == keep_max) ? mine : other;
}

template <int SORT_N>
__global__ void __launch_bounds__(1024)
sort_decode_kernel(const unsigned long long* __restrict__ cand, int cand_stride, int n_take,
                   int idx_bits, int64_t N, const float4* __restrict__ deltas,
                   const float4* __restrict__ anchors, float4 sd,
                   float4* __restrict__ sboxes, int32_t* __restrict__ sorder, int out_stride) {
  constexpr int E = SORT_N / 1024;
  static_assert(E >= 2 && E <= 16, "sort_decode_kernel: 2048 <= SORT_N <= 16384");
  extern __shared__ unsigned long long keys[];
  const int img = blockIdx.x;
  const int tid = threadIdx.x;
  unsigned long long v[E];
  #pragma unroll
  for (int r = 0; r < E; ++r) {
    const int e = tid * E + r;
    v[r] = (e < n_take) ? cand[(size_t)img * cand_stride + e] : 0ull;
  }
  #pragma unroll 1
  for (int k = 2; k <= SORT_N; k <<= 1) {
    // strides handled across threads (j >= E), largest first
    #pragma unroll 1
    for (int j = k >> 1; j >= E; j >>= 1) {
      const int tm = j / E;                              // partner thread = tid ^ tm
      const bool lower = (tid & tm) == 0;
      if (tm >= 32) {
        #pragma unroll
        for (int r = 0; r < E; ++r) keys[r * 1024 + tid] = v[r];          // r-major: conflict-free for a warp
        __syncthreads();
        #pragma unroll
        for (int r = 0; r < E; ++r) {
          const bool desc = (((tid * E + r) & k) == 0);
          bitonic_cx<E>(v[r], keys[r * 1024 + (tid ^ tm)], lower == desc);
        }
        __syncthreads();
      } else {
        #pragma unroll
        for (int r = 0; r < E; ++r) {
          const bool desc = (((tid * E + r) & k) == 0);
          bitonic_cx<E>(v[r], __shfl_xor_sync(0xffffffffu, v[r], tm), lower == desc);
        }
      }
    }
    // strides inside the thread's own registers (j < E)
    #pragma unroll
    for (int j = E >> 1; j > 0; j >>= 1) {
      if (j <= (k >> 1)) {
        #pragma unroll
        for (int r = 0; r < E; ++r) {
          if ((r & j) == 0) {
            const bool desc = (((tid * E + r) & k) == 0);
            const unsigned long long a = v[r], b = v[r | j];
            const bool gt = a > b;
            v[r] = (gt == desc) ? a : b;                 // the lower element keeps the larger key when descending
            v[r | j] = (gt == desc) ? b : a;
          }
        }
      }
    }
  }
  #pragma unroll
  for (int r = 0; r < E; ++r) keys[tid * E + r] = v[r];
  __syncthreads();
  const unsigned long long m = (1ull << idx_bits) - 1ull;
  const float4* d = deltas + (size_t)img * N;
  for (int r = tid; r < n_take; r += 1024) {
    unsigned long long comp = keys[r];
    uint32_t a = (uint32_t)(m - (comp & m));
    float4 box = decode_box(__ldg(&anchors[a]), __ldg(&d[a]), sd);
    sboxes[(size_t)img * out_stride + r] = box;
    sorder[(size_t)img * out_stride + r] = (int32_t)a;
  }
}

// One CTA per image: sequential NMS resolve + output (ProposalLayer.swift:169-192).
__global__ void __launch_bounds__(1024)
proposal_resolve_kernel(const float4* __restrict__ sboxes, const int32_t* __restrict__ sorder,
                        const unsigned long long* __restrict__ mask, int n, int stride, int words,
                        int max_proposals, float4* __restrict__ rois_out,
                        int32_t* __restrict__ keep_anchor, int32_t* __restrict__ count_out) {
  extern __shared__ unsigned long long smem_u64[];
  const int img = blockIdx.x;
  int* kept = nullptr;                  // [max_proposals]
  NmsResolveSmem s = nms_resolve_carve(smem_u64, words, &kept);
  const float4* b = sboxes + (size_t)img * stride;
  const unsigned long long* mk = mask + (size_t)img * stride * words;
  int cnt = nms_resolve(b, nullptr, mk, n, words, max_proposals, 0, 0, s, kept);
  for (int i = threadIdx.x; i < max_proposals; i += blockDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int a = -1;
    if (i < cnt) { int r = kept[i]; v = b[r]; a = sorder[(size_t)img * stride + r]; }
    rois_out[(size_t)img * max_proposals + i] = v;
    if (keep_anchor) keep_anchor[(size_t)img * max_proposals + i] = a;
  }
  if (count_out && threadIdx.x == 0) count_out[img] = cnt;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static int proposal_ensure_ws(mrcnn_ctx* ctx, int batch, int64_t N, int pre) {
  if (batch <= ctx->ws_batch && N <= ctx->ws_anchors && pre <= ctx->ws_pre) return MRCNN_OK;
  MRCNN_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_hist); cudaFree(ctx->d_sel); cudaFree(ctx->d_cand);
  cudaFree(ctx->d_sboxes); cudaFree(ctx->d_sorder); cudaFree(ctx->d_mask);
  int b = batch > ctx->ws_batch ? batch : ctx->ws_batch;
  int64_t n = N > ctx->ws_anchors ? N : ctx->ws_anchors;
  int p = pre > ctx->ws_pre ? pre : ctx->ws_pre;
  int sort_n = next_pow2(p);
  int words = ceil_div(p, 64);
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_hist, sizeof(uint32_t) * RADIX_BINS * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_sel, sizeof(SelState) * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_cand, sizeof(unsigned long long) * (size_t)sort_n * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_sboxes, sizeof(float4) * (size_t)p * b));
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_sorder, sizeof(int32_t) * (size_t)p * b));
  ctx->mask_bytes = sizeof(unsigned long long) * (size_t)p * words * b;
  MRCNN_CUDA_TRY(ctx, cudaMalloc(&ctx->d_mask, ctx->mask_bytes));
  ctx->ws_batch = b; ctx->ws_anchors = n; ctx->ws_pre = p;
  return MRCNN_OK;
}

int proposal_run(mrcnn_ctx* ctx, int batch, int64_t N, const float* d_probs, const float* d_deltas,
                 float* d_rois_out, int32_t* d_keep_anchor, int32_t* d_count) {
  const mrcnn_config& cfg = ctx->cfg;
  MRCNN_REQUIRE(ctx, batch >= 1, "proposal: batch must be >= 1");
  MRCNN_REQUIRE(ctx, N >= 1 && N <= (1ll << 30), "proposal: num_anchors out of range");
  MRCNN_REQUIRE(ctx, ctx->d_anchors && ctx->num_anchors == N,
                "proposal: anchors not loaded or anchor count differs from num_anchors");
  const int pre = (int)(N < cfg.pre_nms_max_proposals ? N : cfg.pre_nms_max_proposals);
  MRCNN_REQUIRE(ctx, pre <= 16384, "proposal: pre_nms_max_proposals > 16384 is not supported");
  MRCNN_REQUIRE(ctx, cfg.max_proposals >= 1 && cfg.max_proposals <= 8192, "proposal: max_proposals out of range");
  int rc = proposal_ensure_ws(ctx, batch, N, pre);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  int idx_bits = 1;
  while ((1ll << idx_bits) < N) ++idx_bits;
  const int total_bits = 32 + idx_bits;
  const int stride = ctx->ws_pre;
  const int sort_n = next_pow2(pre);
  const int cand_stride = next_pow2(ctx->ws_pre);
  const int words = ceil_div(ctx->ws_pre, 64);

  {
  ProfScope ps(ctx, PROF_TOPK_SELECT, (double)batch * N * 8.0);
  sel_init_kernel<<<batch, 256, 0, s>>>(ctx->d_sel, ctx->d_hist, pre);
  MRCNN_LAUNCH_CHECK(ctx);
  dim3 grid(ceil_div(N, SEL_THREADS * SEL_ITEMS), batch);
  int npass = ceil_div(total_bits, RADIX_BITS);
  int hi = total_bits;
  for (int p = 0; p < npass; ++p) {
    int bits = (p == 0) ? (total_bits - RADIX_BITS * (npass - 1)) : RADIX_BITS;
    int shift = hi - bits;
    sel_hist_kernel<<<grid, SEL_THREADS, 0, s>>>((const float2*)d_probs, N, idx_bits, shift, bits,
                                                 p == 0, p == npass - 1, ctx->d_sel, ctx->d_hist);
    MRCNN_LAUNCH_CHECK(ctx);
    hi = shift;
  }
  sel_compact_kernel<<<grid, SEL_THREADS, 0, s>>>((const float2*)d_probs, N, idx_bits, ctx->d_sel,
                                                  ctx->d_cand, cand_stride);
  MRCNN_LAUNCH_CHECK(ctx);
  }
  float4 sd = make_float4(cfg.bbox_std[0], cfg.bbox_std[1], cfg.bbox_std[2], cfg.bbox_std[3]);
#define LAUNCH_SORT(SN)                                                                          \
  do {                                                                                           \
    static bool attr_set_##SN[64] = {false};                                                     \
    if (!attr_set_##SN[ctx->device & 63]) {                                                      \
      MRCNN_CUDA_TRY(ctx, cudaFuncSetAttribute(sort_decode_kernel<SN>,                           \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                               (int)(sizeof(unsigned long long) * SN)));         \
      attr_set_##SN[ctx->device & 63] = true;                                                    \
    }                                                                                            \
    sort_decode_kernel<SN><<<batch, 1024, sizeof(unsigned long long) * SN, s>>>(                                       \
        ctx->d_cand, cand_stride, pre, idx_bits, N, (const float4*)d_deltas,                     \
        (const float4*)ctx->d_anchors, sd, ctx->d_sboxes, ctx->d_sorder, stride);                \
  } while (0)
  {
  ProfScope ps(ctx, PROF_SORT_DECODE, (double)batch * pre * (8.0 + 32.0 + 20.0));
  if (sort_n <= 2048) LAUNCH_SORT(2048);
  else if (sort_n <= 4096) LAUNCH_SORT(4096);
  else if (sort_n <= 8192) LAUNCH_SORT(8192);
  else LAUNCH_SORT(16384);
#undef LAUNCH_SORT
  MRCNN_LAUNCH_CHECK(ctx);
  }

  const int tiles = ceil_div(pre, NMS_TILE);
  dim3 mgrid(tiles, tiles, batch);
  {
  ProfScope ps(ctx, PROF_NMS_MASK, (double)batch * pre * (16.0 + 8.0 * (words + 1) / 2));
  nms_mask_kernel<<<mgrid, NMS_TILE, 0, s>>>(ctx->d_sboxes, nullptr, nullptr, pre, stride, words,
                                             cfg.proposal_nms_iou, ctx->d_mask);
  MRCNN_LAUNCH_CHECK(ctx);
  }
  ProfScope ps2(ctx, PROF_NMS_RESOLVE, (double)batch * (cfg.max_proposals * 8.0 * words + cfg.max_proposals * 20.0));
  size_t rs = nms_resolve_smem_bytes(words) + sizeof(int) * cfg.max_proposals;
  proposal_resolve_kernel<<<batch, 1024, rs, s>>>(ctx->d_sboxes, ctx->d_sorder, ctx->d_mask, pre, stride,
                                                 words, cfg.max_proposals, (float4*)d_rois_out,
                                                 d_keep_anchor, d_count);
  MRCNN_LAUNCH_CHECK(ctx);
  return MRCNN_OK;
}
