// Skinny-M ("swap-AB") bf16 GEMM for sm_100a:  C[M,N] = epi(X[M,K] · W[N,K]^T),  M <= 512 tokens.
//
// At prefill of a short prompt (M ~ 280) or a small decode batch the weights dominate the traffic and
// a 128-row token tile wastes the tensor core (280 rows -> 3 tiles = 384 rows) while leaving SMs idle
// (N/128 x 3 tiles).  Here the roles are swapped: the UMMA M dimension (128 TMEM lanes) carries 128
// WEIGHT rows and the UMMA N dimension carries ALL tokens (padded only to a multiple of 32, issued as
// 256-wide + remainder instructions), so every weight byte is fetched exactly once, and the K
// dimension is split across CTAs so that (N/128) x splits ~ 148 units fill the machine.  Partial sums
// of the splits are parked in fp32 workspace slots and added in a fixed order by the last-arriving
// CTA (deterministic).
//
//   warp 0: TMA producer (W tile 128x64 + token tile M_pad x 64 per k-block, 128B swizzle)
//   warp 1: tcgen05.mma issuer + TMEM allocator (accumulator: 128 lanes x M_pad fp32 columns)
//   warps 2-5: epilogue — lane == output feature n, column == token m: C[m, n] written as 64-byte
//              coalesced segments per token; bias / GELU / residual / SwiGLU as in gemm_tcgen05.cu
#include "common.cuh"
#include "kernels.h"

namespace vb {
namespace {

constexpr int BW = 128;      // weight rows per tile (UMMA M)
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr size_t kCounterBytes = 64 * 1024;

__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  float t = 1.0f - 2.0f / (__expf(2.0f * u) + 1.0f);
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

struct SkinnyArgs {
  __nv_bfloat16* C;
  int ldc, M, N, K;
  int m_pad;        // tokens padded to a multiple of 32 (<= 512)
  int stages;
  int splits;       // k-splits per weight block
  int kb_per_split;
  float* ws;        // [unit][m_pad][128] fp32 partials
  int* counters;    // [n_blocks]
  GemmEpilogue epi;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                   SkinnyArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int x_bytes = a.m_pad * BK * 2;
  const int stage_bytes = BW * BK * 2 + x_bytes;
  uint8_t* bar_base = smem + a.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* acc_full = empty_bar + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  uint32_t* last_flag = tmem_ptr + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x;
  const int n_blk = unit / a.splits, split = unit - n_blk * a.splits;
  const int nkb = (a.K + BK - 1) / BK;
  const int kb0 = split * a.kb_per_split;
  const int kb1 = min(nkb, kb0 + a.kb_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < a.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // weights are parameters: their first stages are requested before the dependency wait
      const int pre = a.epi.static_w ? min(a.stages, kb1 - kb0) : 0;
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], stage_bytes);
        tma_load_2d(smem + i * stage_bytes, &tmap_w, &full_bar[i], (kb0 + i) * BK, n_blk * BW);
      }
      griddep_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        uint8_t* st = smem + stage * stage_bytes;
        if (kb - kb0 >= pre) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_2d(st, &tmap_w, &full_bar[stage], kb * BK, n_blk * BW);
        }
        for (int r = 0; r < a.m_pad; r += 32)  // token tile as 32-row boxes (rows >= M are zero-filled)
          tma_load_2d(st + BW * BK * 2 + r * 128, &tmap_x, &full_bar[stage], kb * BK, r);
        if (++stage == a.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && kb1 > kb0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t wa = smem_u32(smem + stage * stage_bytes);
        const uint32_t xa = wa + BW * BK * 2;
        for (int n0 = 0; n0 < a.m_pad; n0 += 256) {
          const int nn = min(256, a.m_pad - n0);
          const uint32_t idesc = make_idesc_bf16(BW, nn, 0, 0);
          const uint64_t ad = make_smem_desc(wa, 16, 1024, kLayoutSW128);
          const uint64_t bd = make_smem_desc(xa + n0 * 128, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(tmem_base + n0, ad + 2 * k, bd + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == a.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(acc_full);
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int epi_tid = threadIdx.x - 64;
    const int n_local = quad * 32 + lane;
    const int n = n_blk * BW + n_local;
    const bool n_ok = n < a.N;
    griddep_wait();
    if (kb1 > kb0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    bool finalize = true;
    const float* slots = nullptr;
    if (a.splits > 1) {
      float* mine = a.ws + (static_cast<size_t>(unit) * a.m_pad) * BW + n_local;
      for (int c = 0; c < a.m_pad / 32; ++c) {
        uint32_t r[32];
        if (kb1 > kb0) {
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) mine[static_cast<size_t>(c * 32 + j) * BW] = __uint_as_float(r[j]);
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (epi_tid == 0) {
        const int prev = atomicAdd(&a.counters[n_blk], 1);
        const int last = (prev == a.splits - 1) ? 1 : 0;
        if (last) a.counters[n_blk] = 0;
        *last_flag = last;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      finalize = (*last_flag != 0);
      if (finalize) __threadfence();
      slots = a.ws + (static_cast<size_t>(n_blk) * a.splits * a.m_pad) * BW + n_local;
    }
    if (finalize) {
      const float bias = (a.epi.bias != nullptr && n_ok) ? __bfloat162float(a.epi.bias[n]) : 0.f;
      for (int c = 0; c < a.m_pad / 32; ++c) {
        float v[32];
        if (a.splits > 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          for (int s = 0; s < a.splits; ++s) {  // fixed order
            const float* sp = slots + (static_cast<size_t>(s) * a.m_pad + c * 32) * BW;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __ldcg(sp + static_cast<size_t>(j) * BW);
          }
        } else {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int m = c * 32 + j;
          float x = v[j] + bias;
          if (a.epi.swiglu) {
            // lanes (2i, 2i+1) hold (gate_i, up_i) of the interleaved weight rows
            const float xb = bf16_round(x);
            const float other = __shfl_xor_sync(0xffffffffu, xb, 1);
            if (!(lane & 1) && n_ok && m < a.M)
              a.C[static_cast<size_t>(m) * a.ldc + (n >> 1)] =
                  __float2bfloat16(bf16_round(silu_f(xb)) * other);
            continue;
          }
          if (a.epi.act != ACT_NONE) {
            const float xb = bf16_round(x);
            x = a.epi.act == ACT_GELU_TANH ? gelu_tanh_f(xb)
                                           : (a.epi.act == ACT_GELU_ERF ? gelu_erf_f(xb) : silu_f(xb));
          }
          if (n_ok && m < a.M) {
            if (a.epi.residual != nullptr) {
              const int rr = a.epi.res_row_mod > 0 ? (m % a.epi.res_row_mod) : m;
              x = bf16_round(x) + __bfloat162float(a.epi.residual[static_cast<size_t>(rr) * a.epi.ld_res + n]);
            }
            a.C[static_cast<size_t>(m) * a.ldc + n] = __float2bfloat16(x);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// returns -1 if the shape is not handled here
int gemm_skinny_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, __nv_bfloat16* C,
                     int ldc, int M, int N, int K, const GemmEpilogue& epi, cudaStream_t stream) {
  if (M > 512 || M < 1) return -1;
  const int m_pad = (M + 31) / 32 * 32;
  const int sms = num_sms();
  const int n_blocks = (N + BW - 1) / BW;
  const int nkb = (K + BK - 1) / BK;
  int splits = 1;
  if (n_blocks < sms) {
    splits = (sms + n_blocks / 2) / n_blocks;
    if (splits > 8) splits = 8;
    while (splits > 1 && nkb / splits < 4) --splits;
  }
  const int kb_per_split = (nkb + splits - 1) / splits;
  splits = (nkb + kb_per_split - 1) / kb_per_split;  // no empty splits
  void* ws_ptr = nullptr;
  size_t ws_bytes = 0;
  get_workspace(&ws_ptr, &ws_bytes);
  const size_t need = kCounterBytes + static_cast<size_t>(n_blocks) * splits * m_pad * BW * 4;
  if (splits > 1 && (ws_ptr == nullptr || ws_bytes < need || n_blocks > 16384)) {
    splits = 1;
  }
  SkinnyArgs a;
  a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.m_pad = m_pad;
  a.splits = splits;
  a.kb_per_split = splits > 1 ? kb_per_split : nkb;
  a.counters = static_cast<int*>(ws_ptr);
  a.ws = ws_ptr ? reinterpret_cast<float*>(static_cast<char*>(ws_ptr) + kCounterBytes) : nullptr;
  a.epi = epi;
  const int stage_bytes = BW * BK * 2 + m_pad * BK * 2;
  int stages = (216 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return -1;
  a.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 8 * 8 * 2 + 64 + 1024;
  CUtensorMap tw, tx;
  if (make_tmap_2d_bf16(&tw, W, N, K, ldw, BW, BK, 128)) return 1;
  if (make_tmap_2d_bf16(&tx, A, M, K, lda, 32, BK, 128)) return 1;
  static bool attr = false;
  if (!attr) {
    VB_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  VB_CUDA(launch_pdl(gemm_skinny_kernel, dim3(n_blocks * splits), dim3(kThreads), smem, stream, tw, tx, a));
  return 0;
}

}  // namespace vb
