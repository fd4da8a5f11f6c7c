// Flash-attention forward v2 for sm_100a: TWO 128-row query tiles per CTA with ping-pong softmax
// warpgroups, so the tensor core works on one tile while the other tile's softmax runs.
//
//   warps 0-3  (WG0): softmax + O accumulation of query tile 0   (thread == query row == TMEM lane)
//   warps 4-7  (WG1): the same for query tile 1
//   warp  8        : TMA producer — Q0,Q1 once, then ONE ring over the sequence K0,V0,K1,V1,...
//   warp  9        : tcgen05.mma issuer + TMEM allocator
//   TMEM: S0 @0, S1 @128 (fp32 128x128 each), O0 @256, O1 @384
//   MMA issue order per KV block j:  PV0(j) S0(j+1) PV1(j) S1(j+1)  — each softmax warpgroup always
//   has its next S tile being computed while it accumulates O, and the tensor pipe alternates tiles.
//
// Same math / rounding as v1 (fmha_tcgen05.cu): S = QK^T in fp32, online softmax in fp32 with exp2,
// P cast to bf16 for the PV MMA, O accumulated in TMEM by the PV MMA itself with a lazy (exact)
// rescale when a row maximum moves by more than 2^8, d=72 zero-padded to 96
// by TMA OOB fill (SW64 chunks), d=128 as two SW128 chunks, V consumed as the MN-major B operand.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace vb {
namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int kThreads2 = 320;

template <int DP, int CW>
struct Cfg2 {
  static constexpr int kChunks = DP / CW;
  static constexpr int kChunkBytes = 128 * CW * 2;
  static constexpr int kTileBytes = kChunks * kChunkBytes;  // one 128-row Q / K / V tile
  static constexpr uint32_t kLayout = CW == 64 ? kLayoutSW128 : kLayoutSW64;
  static constexpr int kSwizzleBytes = CW * 2;
  static constexpr int kSBO = 8 * CW * 2;
  static constexpr int kPBytes = BQ * BKV * 2;
  static constexpr int kRing = DP == 128 ? 3 : 4;  // K/V ring slots
  static constexpr int kNumBars = 2 + 2 * kRing + 12;
  static constexpr int kSmem = 2 * kTileBytes + kRing * kTileBytes + 2 * kPBytes + kNumBars * 8 + 16 + 1024;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, max relative error 7.5e-5,
// far below the bf16 rounding P gets anyway).  On B200 MUFU.EX2 runs at 16 lanes/clk/SM: the 2 x 16K
// exponentials of one KV block (two query tiles) take as long as its four 128x128x128 MMAs, so the
// softmax warpgroups and the tensor pipe are co-limiting; moving every 4th exponential to the FMA
// pipe takes the MUFU below the MMA time.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float xf = x + 12582912.f;          // 1.5 * 2^23: rounds x to an integer in the low mantissa bits
  const float fr = x - (xf - 12582912.f);   // fractional part in [-0.5, 0.5]
  float p = fmaf(fr, 0.0551716685f, 0.2426111251f);
  p = fmaf(fr, p, 0.6932609677f);
  p = fmaf(fr, p, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
// kPolyEvery (template parameter): every n-th exponential of a row uses ex2_poly (0: none).  MEASURED
// (tools/bench_fmha.py, S = 65.8K causal): 0 -> 963 TF/s, 4 -> 884, 2 -> 830: the softmax warps are
// issue-slot bound as much as MUFU bound, the ~10-instruction polynomial costs more than it frees.
// It stays as a selectable flavour (vila_fmha_cfg 3 / 4); the default is 0.

// Blackwell packed fp32 (two lanes per instruction: FFMA2 / FADD2) and 3-input max (FMNMX3): the
// softmax loop is issue-slot bound, these halve its arithmetic instruction count.
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
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

struct Args2 {
  __nv_bfloat16* o;
  int64_t o_tok_stride, o_head_stride;
  const int32_t* page_table;
  int page_table_stride;
  int Sq, Sk, Hq, Hkv, D, causal, paged;
  float scale_log2;
};

template <int DP, int CW, int kPolyEvery>
__global__ void __launch_bounds__(kThreads2, 1)
fmha2_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, Args2 a) {
  using C = Cfg2<DP, CW>;
  constexpr int R = C::kRing;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* q_s = smem;                              // 2 tiles
  uint8_t* ring_s = q_s + 2 * C::kTileBytes;        // R tiles
  uint8_t* p_s = ring_s + R * C::kTileBytes;        // 2 x [128][128] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * C::kPBytes);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* r_full = bars + 2;             // [R]
  uint64_t* r_empty = bars + 2 + R;        // [R]
  uint64_t* s_full = bars + 2 + 2 * R;     // [2]
  uint64_t* s_free = s_full + 2;           // [2]
  uint64_t* p_full = s_full + 4;           // [2]
  uint64_t* p_free = s_full + 6;           // [2]
  uint64_t* o_full = s_full + 8;           // [2]
  uint64_t* o_free = s_full + 10;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_full + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int hk = h / (a.Hq / a.Hkv);
  const int off = a.Sk - a.Sq;
  // causal: CTAs are dispatched in blockIdx order, so map the FIRST CTAs to the LAST (longest-KV)
  // query tiles — the tail of the grid is then made of the cheapest tiles
  const int qpair = a.causal ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x)
                             : static_cast<int>(blockIdx.x);

  int nblk[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int qt = qpair * 2 + t;
    int n = 0;
    if (qt * BQ < a.Sq) {
      int kv_end = a.Sk;
      if (a.causal) {
        const int q_last = min((qt + 1) * BQ, a.Sq) - 1;
        kv_end = min(a.Sk, q_last + off + 1);
      }
      n = (kv_end + BKV - 1) / BKV;
    }
    nblk[t] = n;
  }
  const int nmax = max(nblk[0], nblk[1]);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_free[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 128);
    }
    for (int i = 0; i < R; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0 && nmax > 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
      for (int t = 0; t < 2; ++t) {
        if (nblk[t] == 0) continue;
        mbar_arrive_expect_tx(&q_full[t], C::kTileBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(q_s + t * C::kTileBytes + c * C::kChunkBytes, &tm_q, &q_full[t], c * CW, h,
                      b * a.Sq + (qpair * 2 + t) * BQ, 0);
      }
      for (int i = 0; i < 2 * nmax; ++i) {  // item 2j = K(j), 2j+1 = V(j)
        const int j = i >> 1;
        const int slot = i % R;
        int tok, page;
        if (a.paged) {
          tok = 0;
          page = a.page_table ? a.page_table[b * a.page_table_stride + j] : j;
        } else {
          tok = b * a.Sk + j * BKV;
          page = 0;
        }
        mbar_wait(&r_empty[slot], ((i / R) & 1) ^ 1);
        mbar_arrive_expect_tx(&r_full[slot], C::kTileBytes);
        const CUtensorMap* tm = (i & 1) ? &tm_v : &tm_k;
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(ring_s + slot * C::kTileBytes + c * C::kChunkBytes, tm, &r_full[slot], c * CW,
                      hk, tok, page);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0 && nmax > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, DP, 0, 1);
      auto ring_wait = [&](int i) { mbar_wait(&r_full[i % R], (i / R) & 1); };
      auto ring_release = [&](int i) { umma_commit(&r_empty[i % R]); };
      auto issue_s = [&](int t, int j) {  // S_t(j) = Q_t K(j)^T
        mbar_wait(&s_free[t], (j & 1) ^ 1);
        tc_fence_after();
        const uint8_t* ks = ring_s + ((2 * j) % R) * C::kTileBytes;
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
#pragma unroll
          for (int k = 0; k < CW / 16; ++k) {
            const uint64_t ad = make_smem_desc(
                smem_u32(q_s + t * C::kTileBytes + c * C::kChunkBytes) + k * 32, 16, C::kSBO, C::kLayout);
            const uint64_t bd =
                make_smem_desc(smem_u32(ks + c * C::kChunkBytes) + k * 32, 16, C::kSBO, C::kLayout);
            umma_f16(tmem_base + t * BKV, ad, bd, idesc_s, (c | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {  // O_t(j) = P_t(j) V(j)
        mbar_wait(&p_full[t], j & 1);  // also orders this PV after any O rescale of block j
        tc_fence_after();
        const uint8_t* vs = ring_s + ((2 * j + 1) % R) * C::kTileBytes;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t ad = make_smem_desc(
              smem_u32(p_s + t * C::kPBytes + (k >> 2) * (BQ * 128)) + (k & 3) * 32, 16, 1024, kLayoutSW128);
          const uint64_t bd = make_smem_desc(smem_u32(vs) + k * 16 * (CW * 2), C::kChunkBytes, C::kSBO,
                                             C::kLayout);
          umma_f16(tmem_base + 256 + t * 128, ad, bd, idesc_o, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&p_free[t]);
        umma_commit(&o_full[t]);
      };
      // prologue: S_t(0)
      ring_wait(0);
      for (int t = 0; t < 2; ++t) {
        if (nblk[t] == 0) continue;
        mbar_wait(&q_full[t], 0);
        issue_s(t, 0);
      }
      ring_release(0);
      for (int j = 0; j < nmax; ++j) {
        // S(j+1) BEFORE PV(j): the softmax warpgroups release the S buffer as soon as they hold block
        // j's scores in registers, so the tensor pipe computes Q K(j+1)^T while they exponentiate
        if (j + 1 < nmax) {
          ring_wait(2 * (j + 1));                   // K(j+1)
          for (int t = 0; t < 2; ++t)
            if (j + 1 < nblk[t]) issue_s(t, j + 1);
          ring_release(2 * (j + 1));
        }
        ring_wait(2 * j + 1);                       // V(j)
        for (int t = 0; t < 2; ++t)
          if (j < nblk[t]) issue_pv(t, j);
        ring_release(2 * j + 1);
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int t = warp >> 2;                 // query tile of this warpgroup
    const int row = threadIdx.x & 127;
    const int qt = qpair * 2 + t;
    const int q_idx = qt * BQ + row;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int n = nblk[t];
    uint8_t* my_p = p_s + t * C::kPBytes;
    // O accumulates in TMEM across KV blocks (the PV MMA runs with accumulate=1).  Each row keeps a
    // reference maximum m_ref; the accumulator is rescaled (tcgen05.ld -> scale -> tcgen05.st) only
    // when a row's running maximum exceeds m_ref by more than 2^8 (exact: P, l and O all use the same
    // m_ref, the final division by l cancels it).  The decision is per warp (TMEM ld/st are
    // warp-collective on the warp's own 32 lanes).
    float m_ref = -INFINITY, l = 0.f;
    const uint32_t o_addr = tmem_base + lane_addr + 256 + t * 128;

    for (int j = 0; j < n; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + t * BKV;
      const int kv0 = j * BKV;
      const bool need_mask = (kv0 + BKV > a.Sk) || (a.causal && (kv0 + BKV - 1 > qt * BQ + off));
      const int kv_lim = a.causal ? min(a.Sk - 1, q_idx + off) : a.Sk - 1;

      // the whole 128-column score row of this thread in registers with ONE TMEM round trip, then the
      // S buffer goes back to the tensor pipe (was: two passes of four dependent 32-column loads)
      uint32_t r[BKV];
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c) tmem_ld_32x32b_x32(s_addr + c * 32, r + c * 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[t]);
      float mx = -INFINITY;
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < BKV; ++i) mx = fmaxf(mx, (kv0 + i <= kv_lim) ? __uint_as_float(r[i]) : -INFINITY);
      } else {
        // four independent max chains (two warps per scheduler: little latency hiding, so the
        // dependent chains of the row reductions are kept short)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < BKV; i += 8) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            m4[q] = fmax3(m4[q], __uint_as_float(r[i + 2 * q]), __uint_as_float(r[i + 2 * q + 1]));
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      const float m_blk = mx * a.scale_log2;
      if (j == 0) {
        m_ref = (m_blk == -INFINITY) ? 0.f : m_blk;
      } else {
        const bool need = m_blk > m_ref + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? m_blk : m_ref;
          const float alpha = ex2f(m_ref - m_new);  // 1 for rows that keep their reference
          mbar_wait(&o_full[t], (j - 1) & 1);       // PV(j-1) has landed in TMEM
          tc_fence_after();
          // rare path (a row maximum moved by more than 2^8): 8 columns at a time so that it adds no
          // register pressure to the hot loop, which holds the whole score row
#pragma unroll 1
          for (int c = 0; c < DP / 8; ++c) {
            uint32_t o8[8];
            tmem_ld_32x32b_x8(o_addr + c * 8, o8);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * alpha);
            tmem_st_32x32b_x8(o_addr + c * 8, o8);
          }
          tmem_st_wait();
          l *= alpha;
          m_ref = m_new;
        }
      }

      mbar_wait(&p_free[t], (j & 1) ^ 1);
      float rowsum = 0.f, rowsum1 = 0.f, rs2 = 0.f, rs3 = 0.f;  // 4 independent accumulation chains
      const float neg_m = -m_ref;
      // 8 columns at a time: exponentials -> bf16 -> one 16-byte store (keeps the live set small: the
      // 128 score registers + 8 temporaries; a 32-wide temporary array spilled to local memory)
#pragma unroll
      for (int g8 = 0; g8 < BKV / 8; ++g8) {
        float p[8];
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float arg = __uint_as_float(r[g8 * 8 + i]) * a.scale_log2 - m_ref;
            const float e = (kv0 + g8 * 8 + i <= kv_lim) ? ex2f(arg) : 0.f;
            p[i] = e;
            rowsum += e;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float a0, a1;
            ffma2(a0, a1, __uint_as_float(r[g8 * 8 + i]), __uint_as_float(r[g8 * 8 + i + 1]), a.scale_log2, neg_m);
            constexpr int kP = kPolyEvery > 0 ? kPolyEvery : 1;
            p[i] = (kPolyEvery > 0 && (i % kP) == kP - 1) ? ex2_poly(a0) : ex2f(a0);
            p[i + 1] = (kPolyEvery > 0 && ((i + 1) % kP) == kP - 1) ? ex2_poly(a1) : ex2f(a1);
            if ((i & 2) == 0) fadd2(rowsum, rowsum1, p[i], p[i + 1]);
            else fadd2(rs2, rs3, p[i], p[i + 1]);
          }
        }
        // P layout: two [128 rows][64 cols] SW128 chunks; 16-byte piece (g8 & 7) of chunk (g8 >> 3)
        uint8_t* prow = my_p + (g8 >> 3) * (BQ * 128) + row * 128;
        uint4 v4;
        v4.x = pack_bf16(p[0], p[1]);
        v4.y = pack_bf16(p[2], p[3]);
        v4.z = pack_bf16(p[4], p[5]);
        v4.w = pack_bf16(p[6], p[7]);
        *reinterpret_cast<uint4*>(prow + (((g8 & 7) ^ (row & 7)) << 4)) = v4;
      }
      rowsum += rowsum1 + (rs2 + rs3);
      l += rowsum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }

    if (n > 0) {
      mbar_wait(&o_full[t], (n - 1) & 1);
      tc_fence_after();
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* dst = a.o + static_cast<int64_t>(b * a.Sq + q_idx) * a.o_tok_stride +
                           static_cast<int64_t>(h) * a.o_head_stride;
#pragma unroll 1
      for (int c = 0; c < DP / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(o_addr + c * 32, r);
        tmem_ld_wait();
        if (q_idx < a.Sq) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c * 32 + g * 8 < a.D) {
              uint4 v4;
              v4.x = pack_bf16(__uint_as_float(r[g * 8 + 0]) * inv, __uint_as_float(r[g * 8 + 1]) * inv);
              v4.y = pack_bf16(__uint_as_float(r[g * 8 + 2]) * inv, __uint_as_float(r[g * 8 + 3]) * inv);
              v4.z = pack_bf16(__uint_as_float(r[g * 8 + 4]) * inv, __uint_as_float(r[g * 8 + 5]) * inv);
              v4.w = pack_bf16(__uint_as_float(r[g * 8 + 6]) * inv, __uint_as_float(r[g * 8 + 7]) * inv);
              stg_v4(dst + c * 32 + g * 8, v4);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int DP, int CW, int kPolyEvery>
int launch_fmha2(const FmhaParams& p, cudaStream_t stream) {
  using C = Cfg2<DP, CW>;
  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[4] = {(uint64_t)p.D, (uint64_t)p.Hq, (uint64_t)p.B * p.Sq, 1};
    uint64_t str[3] = {(uint64_t)p.q_head_stride, (uint64_t)p.q_tok_stride,
                       (uint64_t)p.q_tok_stride * p.B * p.Sq};
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
  Args2 a;
  a.o = p.o;
  a.o_tok_stride = p.o_tok_stride;
  a.o_head_stride = p.o_head_stride;
  a.page_table = p.page_table;
  a.page_table_stride = p.page_table_stride;
  a.Sq = p.Sq; a.Sk = p.Sk; a.Hq = p.Hq; a.Hkv = p.Hkv; a.D = p.D;
  a.causal = p.causal;
  a.paged = paged ? 1 : 0;
  a.scale_log2 = p.scale * 1.4426950408889634f;
  auto kern = fmha2_fwd_kernel<DP, CW, kPolyEvery>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
  }
  const int q_tiles = (p.Sq + BQ - 1) / BQ;
  dim3 grid((q_tiles + 1) / 2, p.Hq, p.B);
  VB_CUDA(launch_pdl(kern, grid, dim3(kThreads2), C::kSmem, stream, tq, tk, tv, a));
  return 0;
}

}  // namespace

// returns -1 when the shape is not handled by v2 (caller uses v1)
int fmha_prefill_v2(const FmhaParams& p, cudaStream_t stream, int poly_every) {  // default 0, see kPolyEvery
  if (p.D == 128) {
    if (poly_every == 4) return launch_fmha2<128, 64, 4>(p, stream);
    if (poly_every == 2) return launch_fmha2<128, 64, 2>(p, stream);
    return launch_fmha2<128, 64, 0>(p, stream);
  }
  if (p.D <= 96 && p.D > 64) {
    if (poly_every == 4) return launch_fmha2<96, 32, 4>(p, stream);
    if (poly_every == 2) return launch_fmha2<96, 32, 2>(p, stream);
    return launch_fmha2<96, 32, 0>(p, stream);
  }
  return -1;
}

}  // namespace vb
