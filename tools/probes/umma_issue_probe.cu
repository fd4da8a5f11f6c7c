// Probe (to run first thing next round): what bounds the tcgen05.mma issue rate?
// One CTA (or one CTA pair), operands resident in shared memory, no TMA traffic.  For each case the
// cycles per "k-block" (4 k16 steps x `chunks` instructions) are reported for
//   order 0: chunk-outer / k-inner   (4 dependent accumulations back to back — the production order)
//   order 1: k-outer / chunk-inner   (independent accumulators interleaved)
//   sync  0: 2000 k-blocks queued, one commit at the end
//   sync  1: tcgen05.commit + mbarrier wait after every k-block (the production pattern, minus the loads)
//   pair  1: tcgen05.mma.cta_group::2 (M = 256 over two CTAs), issued by the leader
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vila_b200/csrc -o /tmp/umma_issue tools/probes/umma_issue_probe.cu && /tmp/umma_issue
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace vb;

template <bool kPair>
__global__ void __launch_bounds__(128, 1)
probe(int N, int chunks, int order, int sync_each, int iters, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;          // 128 x 64 bf16 (this CTA's A rows)
  uint8_t* b_s = smem + 16384;  // up to 512 x 64 bf16 (pair: this CTA's half of every chunk)
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + 65536);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  for (int i = tid; i < (16384 + 65536) / 16; i += 128)  // any finite bf16 pattern will do
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    if (kPair) tmem_alloc_pair<512>(tptr);
    else tmem_alloc<512>(tptr);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tptr;
  const uint32_t idesc = make_idesc_bf16(kPair ? 256 : 128, N, 0, 0);
  const uint64_t ad = make_smem_desc(smem_u32(a_s), 16, 1024, kLayoutSW128);
  const int b_rows = kPair ? N / 2 : N;  // rows of one chunk staged in THIS CTA
  uint32_t phase = 0;
  if (tid == 0 && rank == 0) {
    auto issue = [&](int ch, int k) {
      const uint64_t bd = make_smem_desc(smem_u32(b_s) + ch * b_rows * 128, 16, 1024, kLayoutSW128);
      if (kPair) umma_f16_pair(tm + ch * N, ad + 2 * k, bd + 2 * k, idesc, 1u);
      else umma_f16(tm + ch * N, ad + 2 * k, bd + 2 * k, idesc, 1u);
    };
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (order == 0) {
        for (int ch = 0; ch < chunks; ++ch)
          for (int k = 0; k < 4; ++k) issue(ch, k);
      } else {
        for (int k = 0; k < 4; ++k)
          for (int ch = 0; ch < chunks; ++ch) issue(ch, k);
      }
      if (sync_each) {
        if (kPair) umma_commit_pair(&bar[0]);
        else umma_commit(&bar[0]);
        mbar_wait(&bar[0], phase);
        phase ^= 1;
      }
    }
    if (!sync_each) {
      if (kPair) umma_commit_pair(&bar[0]);
      else umma_commit(&bar[0]);
      mbar_wait(&bar[0], phase);
    }
    cyc[0] = (clock64() - t0) / iters;
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();
  else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair<512>(tm);
    else tmem_dealloc<512>(tm);
  }
}

int main() {
  long long* dcyc;
  cudaMalloc(&dcyc, 16);
  const int smem = 16384 + 65536 + 64 + 1024;
  cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int cases[][2] = {{256, 1}, {256, 2}, {144, 2}, {128, 2}, {128, 4}, {64, 4}, {32, 8}};
  for (int pair = 0; pair < 2; ++pair)
    for (auto& cs : cases)
      for (int order = 0; order < 2; ++order)
        for (int sync_each = 0; sync_each < 2; ++sync_each) {
          const int N = cs[0], chunks = cs[1];
          if (N * chunks > 512) continue;
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(pair ? 2 : 1);
          cfg.blockDim = dim3(128);
          cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = pair ? 2 : 1;
          at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
          cfg.attrs = at;
          cfg.numAttrs = 1;
          cudaError_t e = pair ? cudaLaunchKernelEx(&cfg, probe<true>, N, chunks, order, sync_each, 1000, dcyc)
                               : cudaLaunchKernelEx(&cfg, probe<false>, N, chunks, order, sync_each, 1000, dcyc);
          if (e == cudaSuccess) e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          long long c;
          cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
          const double ideal = 4.0 * chunks * (N / 2.0);  // 128 rows per SM x N x 16 / 4096 MAC/clk
          printf("%s N=%3d x%d  order=%s sync=%s : %5lld clk per k-block (math %4.0f -> %3.0f %%)\n",
                 pair ? "pair " : "single", N, chunks, order ? "k-outer " : "ch-outer", sync_each ? "each" : "none",
                 c, ideal, 100.0 * ideal / c);
        }
  return 0;
}
