// L2 -> shared-memory feed rate of bulk copies (the path the conv kernel's operand ring uses), all SMs at once.
//   mode 0: every CTA streams its own L2-resident region       (A tiles: distinct per CTA)
//   mode 1: every CTA streams the SAME region                  (B tiles: the weights, shared by all CTAs)
//   mode 2: half of the slots from the own region, half shared (a 1x1 layer's operand mix)
//   mode 3: as 2, the shared half fetched once per CLUSTER of `csize` CTAs and multicast (each CTA loads 1/csize of it)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/l2_feed tools/micro/l2_feed.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int SLOT = 32768, NSLOT = 6;

// one thread per CTA issues; slots are recycled as soon as they land (no consumer: pure feed rate)
__global__ void __launch_bounds__(128, 1) feed_kernel(const uint8_t* own, const uint8_t* shared, size_t region, int iters, int mode, int csize) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[NSLOT], ebars[NSLOT];
  const uint32_t s0 = smem_u32(smem), b0 = smem_u32(bars), eb0 = smem_u32(ebars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(b0 + 8 * i, 1); mbar_init(eb0 + 8 * i, csize); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (csize > 1) cluster_sync();
  const uint32_t rank = csize > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    const uint8_t* mine = own + (size_t)blockIdx.x * region;
    const uint32_t nreg = (uint32_t)(region / SLOT);
    for (int it = 0; it < iters + NSLOT; ++it) {
      const int s = it % NSLOT;
      const bool mc = mode == 3 && csize > 1 && (s & 1);
      if (it >= NSLOT) {
        mbar_wait(b0 + 8 * s, ((it / NSLOT) - 1) & 1);
        // a multicast slot is rewritten by every CTA of the cluster: tell all of them that this CTA is done with it
        // (what the consumers' empty barriers do in a real kernel)
        if (mc) for (int r = 0; r < csize; ++r) {
          uint32_t ra; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(eb0 + 8 * s), "r"(r));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
        }
      }
      if (it >= iters) continue;
      if (mc) {
        if (it >= NSLOT) mbar_wait(eb0 + 8 * s, ((it / NSLOT) - 1) & 1);
        mbar_expect_tx(b0 + 8 * s, SLOT);
        const uint32_t part = SLOT / csize;
        bulk_load_mc(s0 + s * SLOT + rank * part, shared + (size_t)((it >> 1) % nreg) * SLOT + rank * part, part, b0 + 8 * s, (uint16_t)((1u << csize) - 1));
      } else {
        mbar_expect_tx(b0 + 8 * s, SLOT);
        const bool sh = mode == 1 || ((mode == 2 || mode == 3) && (it & 1));
        const uint8_t* src = sh ? shared + (size_t)((it >> (mode == 1 ? 0 : 1)) % nreg) * SLOT : mine + (size_t)(it % nreg) * SLOT;
        bulk_load(s0 + s * SLOT, src, SLOT, b0 + 8 * s);
      }
    }
  }
  __syncthreads();
  if (csize > 1) cluster_sync();
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  const size_t region = 512 << 10;           // per-CTA working set: 148 x 512 KB = 74 MB, L2-resident
  uint8_t *own, *shared;
  CK(cudaMalloc(&own, region * sms)); CK(cudaMalloc(&shared, region));
  CK(cudaMemset(own, 1, region * sms)); CK(cudaMemset(shared, 2, region));
  CK(cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SLOT * NSLOT));
  CK(cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 4000;
  struct Case { int mode, csize; const char* name; };
  const Case cases[] = {{0, 1, "own regions (distinct per CTA)"}, {1, 1, "one shared region (all CTAs read the same bytes)"},
                        {2, 1, "half own / half shared, unicast"}, {3, 2, "half own / half shared multicast x2"},
                        {3, 4, "half own / half shared multicast x4"}, {3, 8, "half own / half shared multicast x8"}};
  for (const Case& c : cases) {
    cudaLaunchConfig_t cfg = {};
    int grid = sms;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {(unsigned)c.csize, 1, 1};
    cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = SLOT * NSLOT; cfg.attrs = at; cfg.numAttrs = 1;
    if (c.csize > 1) {
      cfg.gridDim = dim3(sms / c.csize * c.csize);
      int ncl = 0; CK(cudaOccupancyMaxActiveClusters(&ncl, feed_kernel, &cfg));
      grid = ncl * c.csize; if (grid > sms / c.csize * c.csize) grid = sms / c.csize * c.csize;
    }
    cfg.gridDim = dim3(grid);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      CK(cudaLaunchKernelEx(&cfg, feed_kernel, (const uint8_t*)own, (const uint8_t*)shared, region, iters, c.mode, c.csize));
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms;
    }
    const double bytes = (double)grid * iters * SLOT;
    printf("%-52s grid %3d  %8.3f ms  %7.2f TB/s into shared memory  (%5.1f GB/s per SM)\n", c.name, grid, best, bytes / best * 1e-9, bytes / best * 1e-6 / grid);
  }
  return 0;
}
