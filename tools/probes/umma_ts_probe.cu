// Probe: tcgen05.mma with the A operand in TMEM (staged by tcgen05.cp 128x256b from a 128B-swizzled
// K-major shared-memory tile) — (1) does it reproduce the shared-memory-A result bit for bit, (2) what
// does one k-block (K = 64) of the swap-AB skinny GEMM cost per flavour:  SS = A re-read from shared
// memory by every instruction, TS = A copied once into TMEM and shared by the N chunks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vila_b200/csrc -o /tmp/umma_ts tools/probes/umma_ts_probe.cu && /tmp/umma_ts
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
using namespace vb;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

// out: [2][128][256] floats (SS result, TS result); cyc: [2] cycles per k-block (SS, TS)
__global__ void __launch_bounds__(128, 1)
probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, int N, int chunks, int iters,
      float* out, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;                 // 128 x 64 bf16, SW128
  uint8_t* b_s = smem + 16384;         // 512 x 64 bf16, SW128 (chunks * N rows used)
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + 65536);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 8; i += 128) {  // 16-byte pieces
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(a_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
  }
  for (int i = tid; i < 512 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(b_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + (r % 256) * 64 + c * 8);
  }
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tptr;
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  const uint64_t ad = make_smem_desc(smem_u32(a_s), 16, 1024, kLayoutSW128);
  const uint32_t a_tm = tm + 480;
  uint32_t phase = 0;
  if (tid == 0) {
    // ---- correctness: D_ss -> cols [0, N), D_ts -> cols [256, 256 + N) (N <= 224 here) ----
    const uint64_t bd = make_smem_desc(smem_u32(b_s), 16, 1024, kLayoutSW128);
    for (int k = 0; k < 4; ++k) umma_f16(tm, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
    for (int k = 0; k < 4; ++k) tmem_cp_128x256b(a_tm + 8 * k, ad + 2 * k);
    for (int k = 0; k < 4; ++k) umma_f16_ts(tm + 256, a_tm + 8 * k, bd + 2 * k, idesc, k ? 1u : 0u);
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], phase);
  phase ^= 1;
  tc_fence_after();
  for (int c = 0; c < 256 / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(0 * 128 + warp * 32 + lane) * 256 + c * 32 + j] = __uint_as_float(r[j]);
    tmem_ld_32x32b_x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + 256 + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c * 32 + j < 224) out[(1 * 128 + warp * 32 + lane) * 256 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- timing: `chunks` N-wide instructions per k16 step, 4 steps per k-block ----
  for (int mode = 0; mode < 2; ++mode) {
    long long t0 = 0;
    if (tid == 0) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (mode == 1)
          for (int k = 0; k < 4; ++k) tmem_cp_128x256b(a_tm + 8 * k, ad + 2 * k);
        for (int ch = 0; ch < chunks; ++ch) {
          const uint64_t bd = make_smem_desc(smem_u32(b_s) + ch * N * 128, 16, 1024, kLayoutSW128);
          for (int k = 0; k < 4; ++k) {
            if (mode == 0) umma_f16(tm + ch * N, ad + 2 * k, bd + 2 * k, idesc, 1u);
            else umma_f16_ts(tm + ch * N, a_tm + 8 * k, bd + 2 * k, idesc, 1u);
          }
        }
      }
      umma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], phase);
    phase ^= 1;
    if (tid == 0) cyc[mode] = (clock64() - t0) / iters;
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tm);
  }
}

int main() {
  std::vector<__nv_bfloat16> hA(128 * 64), hB(256 * 64);
  srand(1);
  for (auto& v : hA) v = __float2bfloat16((float)(rand() % 7 - 3));
  for (auto& v : hB) v = __float2bfloat16((float)(rand() % 5 - 2));
  __nv_bfloat16 *dA, *dB;
  float* dout;
  long long* dcyc;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dout, 2 * 128 * 256 * 4); cudaMalloc(&dcyc, 16);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 16384 + 65536 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int cases[][2] = {{144, 2}, {128, 2}, {224, 1}, {96, 3}, {64, 4}};
  for (auto& cs : cases) {
    const int N = cs[0], chunks = cs[1];
    cudaMemset(dout, 0, 2 * 128 * 256 * 4);
    probe<<<1, 128, smem>>>(dA, dB, N, chunks, 2000, dout, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d error: %s\n", N, cudaGetErrorString(e)); return 1; }
    std::vector<float> o(2 * 128 * 256);
    long long cyc[2];
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost);
    double e_ss = 0, e_ts = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N && n < 224; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += __bfloat162float(hA[m * 64 + k]) * __bfloat162float(hB[n * 64 + k]);
        e_ss = fmax(e_ss, fabs(o[(0 * 128 + m) * 256 + n] - ref));
        e_ts = fmax(e_ts, fabs(o[(1 * 128 + m) * 256 + n] - ref));
      }
    printf("N=%3d x%d chunks: max|err| SS=%g TS=%g ; cycles per k-block (K=64): SS=%lld TS=%lld\n", N, chunks, e_ss,
           e_ts, cyc[0], cyc[1]);
  }
  return 0;
}
