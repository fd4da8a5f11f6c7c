// Probe: does TMA multicast across a 2-CTA cluster relieve the per-SM L2->SM port?
// Every CTA fills 160 KB of shared memory per iteration from an L2-resident buffer, either
//   mode 0: unicast   — each CTA requests all 160 KB itself
//   mode 1: multicast — each CTA requests 80 KB with .multicast::cluster mask 0b11 (both CTAs receive)
// and the two modes use the same barrier / cluster-sync structure.  Build + run (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mc_probe tools/probes/mc_probe.cu && /tmp/mc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kChunk = 16 * 1024;
constexpr int kChunks = 10;  // 160 KB per iteration

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const char* __restrict__ src, int iters, int mode, int shared_src, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kChunks * kChunk);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cluster_id = blockIdx.x >> 1;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const char* base = src + (shared_src ? 0 : (size_t)cluster_id * kChunks * kChunk);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(kChunks * kChunk) : "memory");
      if (mode == 0) {
        for (int c = 0; c < kChunks; ++c)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                       "r"(s32(smem + c * kChunk)), "l"(base + c * kChunk), "r"(kChunk), "r"(s32(bar)) : "memory");
      } else {
        for (int c = rank; c < kChunks; c += 2)  // my half of the chunks, delivered to both CTAs
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
                       "r"(s32(smem + c * kChunk)), "l"(base + c * kChunk), "r"(kChunk), "r"(s32(bar)), "h"((uint16_t)3) : "memory");
      }
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                     : "=r"(ok) : "r"(s32(bar)), "r"(it & 1) : "memory");
      }
    }
    // nobody overwrites a stage the peer may still be receiving into
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms / 2 * 2;
  char* src;
  long long* out;
  const size_t bytes = (size_t)(grid / 2) * kChunks * kChunk;
  cudaMalloc(&src, bytes);
  cudaMemset(src, 1, bytes);
  cudaMalloc(&out, grid * sizeof(long long));
  const int smem = kChunks * kChunk + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 400;
  for (int shared_src = 1; shared_src >= 0; --shared_src)
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        probe<<<grid, 128, smem>>>(src, iters, mode, shared_src, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      static long long h[512];
      cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < grid; ++i) s += (double)h[i];
      const double cyc = s / grid / iters;
      printf("src=%s mode=%s: %.0f cycles per 160 KB per SM = %.1f B/clk/SM received\n",
             shared_src ? "shared" : "per-cluster", mode ? "multicast" : "unicast", cyc, kChunks * kChunk / cyc);
    }
  return 0;
}
