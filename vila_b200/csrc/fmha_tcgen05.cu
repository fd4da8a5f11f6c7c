// Flash-attention forward for sm_100a with tcgen05 tensor cores and TMEM accumulators.
//
// One CTA = one 128-row query tile of one (batch, head).  Warp roles:
//   warps 0-3 : softmax + output accumulation; thread t owns query row t (TMEM lane t)
//   warp  4   : TMA producer (Q once, then K/V blocks of 128 tokens, double buffered)
//   warp  5   : tcgen05.mma issuer + TMEM allocator
// Per KV block j:   S_j = Q K_j^T  (TMEM, double buffered so S_{j+1} overlaps softmax_j)
//                   P_j = exp2(S_j*scale - m)  -> bf16 -> shared (128B-swizzled, K-major A operand)
//                   O_j = P_j V_j  (TMEM, V is the MN-major B operand straight from its [tok, d] layout)
//                   o_acc = o_acc * alpha + O_j   (fp32 registers)
// Head dims that are not a multiple of the swizzle chunk (SigLIP d=72) are zero-padded for free
// by TMA out-of-bounds fill (72 -> 96 = 3 x 32-column SW64 chunks); the LLM d=128 uses 2 x SW128.
//
// Replaces flash_attn_func in SiglipFlashAttention2 (modeling_siglip.py:583-585, non-causal,
// scale 72^-0.5) and HF _flash_attention_forward for Qwen2 (modeling_qwen2.py:191-310; causal GQA).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vb {

namespace {

constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 128;  // kv rows per block (== KV page size)
constexpr int kThreads = 192;

template <int DP, int CW>
struct FmhaCfg {
  static_assert(DP % CW == 0, "");
  static constexpr int kChunks = DP / CW;
  static constexpr int kChunkBytes = 128 * CW * 2;  // [128 rows][CW] bf16
  static constexpr int kTileBytes = kChunks * kChunkBytes;
  static constexpr uint32_t kLayout = CW == 64 ? kLayoutSW128 : kLayoutSW64;
  static constexpr int kSwizzleBytes = CW * 2;
  static constexpr int kSBO = 8 * CW * 2;  // 8-row group pitch
  static constexpr int kPBytes = BQ * BKV * 2;
  static constexpr int kNumBars = 20;
  static constexpr int kSmem = kTileBytes * 5 + 2 * kPBytes + kNumBars * 8 + 16 + 1024;  // P double-buffered
  static constexpr int kTmemO = 256;  // S0 @ 0, S1 @ 128, O @ 256
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%4}; mov.b64 rc, {%5,%5}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{.reg .b64 ra, rd; mov.b64 ra, {%2,%3}; mov.b64 rd, {%0,%1}; add.rn.f32x2 rd, rd, ra; "
      "mov.b64 {%0,%1}, rd;}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}

struct FmhaKernelArgs {
  __nv_bfloat16* o;
  int64_t o_tok_stride, o_head_stride;
  const int32_t* page_table;
  int page_table_stride;
  int Sq, Sk, Hq, Hkv, D, causal, paged;
  float scale_log2;
  // split-KV decode mode (fmha_decode_split): blockIdx.z is a KV split of ONE sequence, not a batch
  // entry.  Split z covers tokens [z*split_tokens, min(Sk_total, (z+1)*split_tokens)); Sk_total is read
  // from device memory (*sk_dev + 1: the position of the token being decoded), the partial outputs go
  // to o_partial (fp32, normalised per split) with their log2-sum-exp in lse_out.
  const int32_t* sk_dev;
  int split_tokens;
  float* o_partial;  // [splits][Hq][Sq][D]
  float* lse_out;    // [splits][Hq][Sq]
  int* split_counters;  // [Hq] zero-initialised, self-cleaning; non-null: the last CTA of a head combines into o
};

template <int DP, int CW>
__global__ void __launch_bounds__(kThreads, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, FmhaKernelArgs a) {
  using C = FmhaCfg<DP, CW>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + C::kTileBytes;       // 2 stages
  uint8_t* v_s = k_s + 2 * C::kTileBytes;   // 2 stages
  uint8_t* p_s = v_s + 2 * C::kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * C::kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;   // [2]
  uint64_t* s_free = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;  // [2]
  uint64_t* p_free = bars + 15;  // [2]
  uint64_t* o_full = bars + 17;
  uint64_t* o_free = bars + 18;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // causal: heaviest (last) query tiles first, see fmha2_tcgen05.cu
  const int qt = a.causal ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x)
                          : static_cast<int>(blockIdx.x);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int hk = h / (a.Hq / a.Hkv);
  const bool split_mode = a.split_tokens > 0;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_free[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();
  griddep_wait();  // Q/K/V come from the predecessor; O may alias memory it still reads

  // KV extent of this CTA (after the dependency wait: in split mode it comes from device memory)
  int Sk = a.Sk;
  int blk0 = 0;              // first KV block (page-table index) of this CTA
  const int qb = split_mode ? 0 : b;  // batch entry the queries / outputs belong to
  if (split_mode) {
    const int total = *a.sk_dev + 1;
    const int start = b * a.split_tokens;
    blk0 = start / BKV;
    Sk = max(0, min(a.split_tokens, total - start));
  }
  const int off = Sk - a.Sq;  // causal diagonal offset
  int kv_end = Sk;
  if (a.causal) {
    int q_last = min((qt + 1) * BQ, a.Sq) - 1;
    kv_end = min(Sk, q_last + off + 1);
  }
  const int nblk = (kv_end + BKV - 1) / BKV;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0 && nblk > 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
      mbar_arrive_expect_tx(q_full, C::kTileBytes);
#pragma unroll
      for (int c = 0; c < C::kChunks; ++c)
        tma_load_4d(q_s + c * C::kChunkBytes, &tm_q, q_full, c * CW, h, qb * a.Sq + qt * BQ, 0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1;
        int tok, page;
        if (a.paged) {
          tok = 0;
          page = a.page_table ? a.page_table[split_mode ? blk0 + j : b * a.page_table_stride + j] : blk0 + j;
        } else {
          tok = b * a.Sk + j * BKV;
          page = 0;
        }
        mbar_wait(&k_empty[s], par);
        mbar_arrive_expect_tx(&k_full[s], C::kTileBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(k_s + s * C::kTileBytes + c * C::kChunkBytes, &tm_k, &k_full[s], c * CW, hk,
                      tok, page);
        mbar_wait(&v_empty[s], par);
        mbar_arrive_expect_tx(&v_full[s], C::kTileBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(v_s + s * C::kTileBytes + c * C::kChunkBytes, &tm_v, &v_full[s], c * CW, hk,
                      tok, page);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0 && nblk > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, DP, 0, 1);   // P (K-major) x V (MN-major)
      auto issue_s = [&](int j) {
        const int s = j & 1;
        const uint32_t u = (j >> 1) & 1;
        mbar_wait(&k_full[s], u);
        mbar_wait(&s_free[s], u ^ 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
#pragma unroll
          for (int k = 0; k < CW / 16; ++k) {
            const uint64_t ad =
                make_smem_desc(smem_u32(q_s + c * C::kChunkBytes) + k * 32, 16, C::kSBO, C::kLayout);
            const uint64_t bd = make_smem_desc(
                smem_u32(k_s + s * C::kTileBytes + c * C::kChunkBytes) + k * 32, 16, C::kSBO,
                C::kLayout);
            umma_f16(tmem_base + s * BKV, ad, bd, idesc_s, (c | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        const int s = j & 1;
        mbar_wait(&p_full[s], (j >> 1) & 1);
        mbar_wait(&v_full[s], (j >> 1) & 1);
        mbar_wait(o_free, (j & 1) ^ 1);  // O_{j-1} has been read out of TMEM
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A = P: two 64-column SW128 chunks of [128][64]; step 16 columns = 32 bytes
          const uint64_t ad = make_smem_desc(
              smem_u32(p_s + s * C::kPBytes + (k >> 2) * (BQ * 128)) + (k & 3) * 32, 16, 1024, kLayoutSW128);
          // B = V (MN-major): N spans the d-chunks (LBO = chunk pitch), K = 16 token rows per MMA
          const uint64_t bd =
              make_smem_desc(smem_u32(v_s + s * C::kTileBytes) + k * 16 * (CW * 2),
                             C::kChunkBytes, C::kSBO, C::kLayout);
          umma_f16(tmem_base + C::kTmemO, ad, bd, idesc_o, k != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[s]);
        umma_commit(&p_free[s]);
        umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / accumulate (thread == query row) =====================
    const int row = threadIdx.x;
    const int q_idx = qt * BQ + row;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float o_acc[DP];
#pragma unroll
    for (int i = 0; i < DP; ++i) o_acc[i] = 0.f;

    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1;
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + s * BKV;
      const int kv0 = j * BKV;
      const bool need_mask = (kv0 + BKV > Sk) || (a.causal && (kv0 + BKV - 1 > qt * BQ + off));
      const int kv_lim = a.causal ? min(Sk - 1, q_idx + off) : Sk - 1;  // last valid kv index

      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, r);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            mx = fmaxf(mx, (kv0 + c * 32 + i <= kv_lim) ? __uint_as_float(r[i]) : -INFINITY);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      float m_new = fmaxf(m, mx * a.scale_log2);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2(m - m_use);  // m == -inf -> 0

      // pass 2: P = exp2(S*scale - m) -> bf16 -> swizzled smem (double-buffered: P_j is written
      // while the tensor core is still reading P_{j-1})
      mbar_wait(&p_free[s], ((j >> 1) & 1) ^ 1);
      float rowsum = 0.f, rowsum1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, r);
        tmem_ld_wait();
        float p[32];
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = (kv0 + c * 32 + i <= kv_lim) ? ex2(__uint_as_float(r[i]) * a.scale_log2 - m_use) : 0.f;
            p[i] = e;
            rowsum += e;
          }
        } else {  // packed fp32 (FFMA2 / FADD2): the softmax warps are issue-slot bound
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0, a1;
            ffma2(a0, a1, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), a.scale_log2, -m_use);
            p[i] = ex2(a0);
            p[i + 1] = ex2(a1);
            fadd2(rowsum, rowsum1, p[i], p[i + 1]);
          }
        }
        uint8_t* prow = p_s + s * C::kPBytes + (c >> 1) * (BQ * 128) + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int piece = (c & 1) * 4 + g;
          uint4 v4;
          v4.x = pack_bf16(p[g * 8 + 0], p[g * 8 + 1]);
          v4.y = pack_bf16(p[g * 8 + 2], p[g * 8 + 3]);
          v4.z = pack_bf16(p[g * 8 + 4], p[g * 8 + 5]);
          v4.w = pack_bf16(p[g * 8 + 6], p[g * 8 + 7]);
          *reinterpret_cast<uint4*>(prow + ((piece ^ (row & 7)) << 4)) = v4;
        }
      }
      l = l * alpha + (rowsum + rowsum1);
      m = m_new;
      fence_proxy_async_smem();
      mbar_arrive(&p_full[s]);
      tc_fence_before();
      mbar_arrive(&s_free[s]);

      // Deferred accumulation: P_{j-1} V_{j-1} ran on the tensor core while this tile's softmax was
      // computed; fold it in now.  o_acc holds acc_j - P_j V_j = (o_acc + P_{j-1} V_{j-1}) * alpha_j.
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        const uint32_t o_addr = tmem_base + lane_addr + C::kTmemO;
#pragma unroll
        for (int c = 0; c < DP / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(o_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            o_acc[c * 32 + i] = (o_acc[c * 32 + i] + __uint_as_float(r[i])) * alpha;
        }
        tc_fence_before();
        mbar_arrive(o_free);
      }
    }
    if (nblk > 0) {  // the last tile's P V
      mbar_wait(o_full, (nblk - 1) & 1);
      tc_fence_after();
      const uint32_t o_addr = tmem_base + lane_addr + C::kTmemO;
#pragma unroll
      for (int c = 0; c < DP / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(o_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(r[i]);
      }
    }

    if (split_mode) {
      if (q_idx < a.Sq) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const int64_t r = (static_cast<int64_t>(b) * a.Hq + h) * a.Sq + q_idx;
        a.lse_out[r] = l > 0.f ? m + log2f(l) : -INFINITY;
        float4* dst = reinterpret_cast<float4*>(a.o_partial + r * a.D);
#pragma unroll
        for (int g = 0; g < DP / 4; ++g) {
          if (g * 4 < a.D)
            dst[g] = make_float4(o_acc[g * 4] * inv, o_acc[g * 4 + 1] * inv, o_acc[g * 4 + 2] * inv,
                                 o_acc[g * 4 + 3] * inv);
        }
      }
      if (a.split_counters != nullptr) {
        // Fused combine: the LAST split CTA of this KV head to finish merges all partials (fixed split
        // order -> deterministic) and writes the bf16 output, saving the separate combine launch.
        // MEASURED SLOWER than the separate 28-CTA combine kernel (one CTA per KV head walks
        // G x splits partial rows serially: 54 vs 28.6 us at 16.4K tokens, tools/bench_decode_attn.py);
        // kept as an option of the entry point, not used by the decoder.
        __shared__ int is_last_s;
        __threadfence();                                   // this thread's partial row is visible device-wide
        asm volatile("bar.sync 1, 128;" ::: "memory");    // the four softmax warps
        if (threadIdx.x == 0) {
          const int prev = atomicAdd(&a.split_counters[h], 1);
          is_last_s = (prev == static_cast<int>(gridDim.z) - 1);
          if (is_last_s) a.split_counters[h] = 0;          // re-arm for the next launch / graph replay
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (is_last_s) {
          __threadfence();
          const int nsp = static_cast<int>(gridDim.z);
          const int d = threadIdx.x;                       // 128 threads <-> D = 128 columns
          for (int g = 0; g < a.Sq; ++g) {
            float mx = -INFINITY;
            for (int s = 0; s < nsp; ++s)
              mx = fmaxf(mx, __ldcg(&a.lse_out[(static_cast<int64_t>(s) * a.Hq + h) * a.Sq + g]));
            float den = 0.f, acc = 0.f;
            for (int s = 0; s < nsp; ++s) {
              const int64_t r = (static_cast<int64_t>(s) * a.Hq + h) * a.Sq + g;
              const float v = __ldcg(&a.lse_out[r]);
              const float w = (v == -INFINITY) ? 0.f : exp2f(v - mx);
              den += w;
              acc += w * __ldcg(&a.o_partial[r * a.D + d]);
            }
            const float inv_den = den > 0.f ? 1.f / den : 0.f;  // same arithmetic as decode_combine_kernel
            a.o[static_cast<int64_t>(g) * a.o_tok_stride + static_cast<int64_t>(h) * a.o_head_stride + d] =
                __float2bfloat16(acc * inv_den);
          }
        }
      }
    } else if (q_idx < a.Sq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* dst = a.o + static_cast<int64_t>(b * a.Sq + q_idx) * a.o_tok_stride +
                           static_cast<int64_t>(h) * a.o_head_stride;
#pragma unroll
      for (int g = 0; g < DP / 8; ++g) {
        if (g * 8 < a.D) {
          uint4 v4;
          v4.x = pack_bf16(o_acc[g * 8 + 0] * inv, o_acc[g * 8 + 1] * inv);
          v4.y = pack_bf16(o_acc[g * 8 + 2] * inv, o_acc[g * 8 + 3] * inv);
          v4.z = pack_bf16(o_acc[g * 8 + 4] * inv, o_acc[g * 8 + 5] * inv);
          v4.w = pack_bf16(o_acc[g * 8 + 6] * inv, o_acc[g * 8 + 7] * inv);
          stg_v4(dst + g * 8, v4);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

struct SplitArgs {
  const int32_t* sk_dev;
  int split_tokens;
  float* o_partial;
  float* lse_out;
  int* counters;
};

template <int DP, int CW>
int launch_fmha(const FmhaParams& p, cudaStream_t stream, const SplitArgs* split = nullptr) {
  using C = FmhaCfg<DP, CW>;
  CUtensorMap tq, tk, tv;
  {
    // split mode: the queries are ONE set of Sq rows shared by all splits (blockIdx.z)
    const uint64_t q_rows = (uint64_t)(split ? 1 : p.B) * p.Sq;
    uint64_t dims[4] = {(uint64_t)p.D, (uint64_t)p.Hq, q_rows, 1};
    uint64_t str[3] = {(uint64_t)p.q_head_stride, (uint64_t)p.q_tok_stride,
                       (uint64_t)p.q_tok_stride * q_rows};
    uint32_t box[4] = {CW, 1, BQ, 1};
    if (make_tmap_nd_bf16(&tq, p.q, 4, dims, str, box, C::kSwizzleBytes)) return 1;
  }
  const bool paged = p.kv_page_stride != 0;
  {
    uint64_t dims[4], str[3];
    if (paged) {
      dims[0] = p.D; dims[1] = p.Hkv; dims[2] = BKV; dims[3] = p.kv_num_pages;
      str[0] = p.kv_head_stride; str[1] = p.kv_tok_stride; str[2] = p.kv_page_stride;
    } else {
      dims[0] = p.D; dims[1] = p.Hkv; dims[2] = (uint64_t)p.B * p.Sk; dims[3] = 1;
      str[0] = p.kv_head_stride; str[1] = p.kv_tok_stride;
      str[2] = (uint64_t)p.kv_tok_stride * p.B * p.Sk;
    }
    uint32_t box[4] = {CW, 1, BKV, 1};
    if (make_tmap_nd_bf16(&tk, p.k, 4, dims, str, box, C::kSwizzleBytes)) return 1;
    if (make_tmap_nd_bf16(&tv, p.v, 4, dims, str, box, C::kSwizzleBytes)) return 1;
  }
  FmhaKernelArgs a;
  a.o = p.o;
  a.o_tok_stride = p.o_tok_stride;
  a.o_head_stride = p.o_head_stride;
  a.page_table = p.page_table;
  a.page_table_stride = p.page_table_stride;
  a.Sq = p.Sq; a.Sk = p.Sk; a.Hq = p.Hq; a.Hkv = p.Hkv; a.D = p.D;
  a.causal = p.causal;
  a.paged = paged ? 1 : 0;
  a.scale_log2 = p.scale * 1.4426950408889634f;
  a.sk_dev = split ? split->sk_dev : nullptr;
  a.split_tokens = split ? split->split_tokens : 0;
  a.o_partial = split ? split->o_partial : nullptr;
  a.lse_out = split ? split->lse_out : nullptr;
  a.split_counters = split ? split->counters : nullptr;
  auto kern = fmha_fwd_kernel<DP, CW>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
  }
  dim3 grid((p.Sq + BQ - 1) / BQ, p.Hq, p.B);
  VB_CUDA(launch_pdl(kern, grid, dim3(kThreads), C::kSmem, stream, tq, tk, tv, a));
  return 0;
}

}  // namespace

int fmha_prefill(const FmhaParams& p, cudaStream_t stream) { return fmha_prefill_cfg(0, p, stream); }

// Split-KV attention of ONE long sequence for a handful of query rows (decode: the G query heads of a
// KV group are the "rows" of the 128-row tile, the KV heads are the "heads"): p.B = number of KV
// splits of split_tokens tokens each, the sequence length is *n_tok_minus_1 + 1 (device memory, read
// after the dependency wait, so one captured graph serves every decode position), K/V paged.  Writes
// normalised fp32 partials [B][Hq][Sq][D] and their log2-sum-exp [B][Hq][Sq]; non-causal.
int fmha_decode_split(const FmhaParams& p, const int32_t* n_tok_minus_1, int split_tokens,
                      float* o_partial, float* lse, int* counters, cudaStream_t stream) {
  VB_CHECK(p.D == 128, "fmha_decode_split: head dim must be 128");
  VB_CHECK(p.kv_page_stride != 0 && p.page_table != nullptr, "fmha_decode_split: K/V must be paged");
  VB_CHECK(split_tokens > 0 && split_tokens % BKV == 0, "fmha_decode_split: split_tokens %% 128 != 0");
  VB_CHECK(p.Sq >= 1 && p.Sq <= BQ && !p.causal, "fmha_decode_split: 1..128 query rows, non-causal");
  VB_CHECK(n_tok_minus_1 && o_partial && lse, "fmha_decode_split: null output / length pointer");
  VB_CHECK(counters == nullptr || p.o != nullptr, "fmha_decode_split: fused combine needs the output pointer");
  SplitArgs sa{n_tok_minus_1, split_tokens, o_partial, lse, counters};
  return launch_fmha<128, 64>(p, stream, &sa);
}

// variant: 0 = size heuristic, 1 = force the one-tile-per-CTA kernel, 2 = force the two-tile kernel
// (falls through to v1 only when v2 does not cover the head dim / Sq <= 128).
int fmha_prefill_cfg(int variant_in, const FmhaParams& p, cudaStream_t stream) {
  int variant = variant_in;
  VB_CHECK(variant >= 0 && variant <= 4, "fmha: unknown variant %d", variant);
  const int poly = variant == 3 ? 4 : (variant == 4 ? 2 : 0);  // 3 / 4: two-tile kernel with every 4th / 2nd exp2 as a polynomial
  if (variant >= 3) variant = 2;
  VB_CHECK(p.B > 0 && p.Sq > 0 && p.Sk > 0, "fmha: empty problem");
  VB_CHECK(p.Hq % p.Hkv == 0, "fmha: Hq (%d) must be a multiple of Hkv (%d)", p.Hq, p.Hkv);
  VB_CHECK(p.D % 8 == 0, "fmha: head dim must be a multiple of 8 (got %d)", p.D);
  VB_CHECK(p.q_tok_stride % 8 == 0 && p.q_head_stride % 8 == 0 && p.kv_tok_stride % 8 == 0 &&
               p.kv_head_stride % 8 == 0 && p.o_tok_stride % 8 == 0 && p.o_head_stride % 8 == 0,
           "fmha: strides must be multiples of 8 elements (16 bytes)");
  VB_CHECK(!p.causal || p.Sk >= p.Sq, "fmha: causal needs Sk >= Sq");
  {
    // v2 (two query tiles per CTA, ping-pong softmax warpgroups) once its 256-row CTAs fill the
    // machine; below that the one-tile-per-CTA kernel has twice the CTAs and the shorter critical
    // path (1 image, 1024 x 16 heads: 22.5 vs 28.2 us; 64 images: 1105 vs 729 us).
    const long v2_ctas = static_cast<long>((p.Sq + 255) / 256) * p.Hq * p.B;
    const bool use_v2 = variant == 2 ? p.Sq > 128
                                     : (variant == 1 ? false : (p.Sq > 128 && v2_ctas >= num_sms()));
    if (use_v2) {
      const int rc = fmha_prefill_v2(p, stream, poly);
      if (rc >= 0) return rc;
      VB_CHECK(variant != 2, "fmha: the two-tile kernel does not cover head dim %d", p.D);
    }
  }
  if (p.D == 128) return launch_fmha<128, 64>(p, stream);
  if (p.D <= 96 && p.D > 64) return launch_fmha<96, 32>(p, stream);
  if (p.D == 64) return launch_fmha<64, 64>(p, stream);
  set_last_error("fmha: unsupported head dim %d (supported: 64, 65..96, 128)", p.D);
  return 1;
}

}  // namespace vb
