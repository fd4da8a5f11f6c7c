// Skinny-M ("swap-AB") bf16 GEMM for sm_100a:  C[M,N] = epi(X[M,K] · W[N,K]^T),  M <= 512 tokens.
//
// At prefill of a short prompt (M ~ 280) or a small decode batch the weights dominate the traffic and
// a 128-row token tile wastes the tensor core (280 rows -> 3 tiles = 384 rows), re-reads the weight
// tile once per token tile through the ~50 B/clk L2->SM port, and leaves SMs idle.  Here the roles
// are swapped: the UMMA M dimension (TMEM lanes) carries WEIGHT rows and the UMMA N dimension carries
// ALL tokens (padded only to a multiple of 32, issued as one or two equal UMMA N chunks), so every
// weight byte enters an SM exactly once, and the K dimension is split across CTAs so that the machine
// is filled.  Partial sums of the splits are parked in fp32 workspace slots and added in a fixed
// order by the last-arriving CTA (deterministic).
//
// Two flavours of the same kernel:
//   kPair = false : one CTA = 128 weight rows, tcgen05.mma.cta_group::1, loads the whole token tile
//   kPair = true  : a 2-CTA cluster = 256 weight rows, ONE tcgen05.mma.cta_group::2 (M = 256) issued
//                   by the leader CTA; each CTA stages its own 128 weight rows plus HALF of the token
//                   tile (the tensor cores of both SMs read both halves), which halves the token
//                   bytes per SM — the term that bounds the single-CTA flavour.  Both CTAs' TMA loads
//                   signal the leader's "full" barrier; the leader's tcgen05.commit multicasts the
//                   "stage free" / "accumulator ready" arrivals to both CTAs.
//
//   warp 0: TMA producer (W tile 128x64 + token rows x 64 per k-block, 128B swizzle)
//   warp 1: tcgen05.mma issuer (leader CTA only in pair mode) + TMEM allocator
//           (accumulator: 128 lanes x M_pad fp32 columns per CTA)
//   warps 2-9: epilogue, two warpgroups taking alternate 32-token chunks — TMEM lane == output
//              feature n, column == token m.  Each chunk is transposed through a shared-memory
//              staging tile so that C[m, n0..] leaves the SM as 16-byte vectors in 256-byte rows;
//              bias / GELU / residual / SwiGLU with the rounding points of gemm_tcgen05.cu.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace vb {
namespace {

constexpr int BW = 128;      // weight rows per CTA (TMEM lanes)
constexpr int BK = 64;
constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 two epilogue warpgroups
constexpr size_t kCounterBytes = 64 * 1024;
constexpr int kStgPitch = BW * 2 + 16;           // staging row (one token, 128 features) + pad
constexpr int kStgBytes = 2 * 32 * kStgPitch;    // double buffer of 32 tokens


struct SkinnyArgs {
  __nv_bfloat16* C;
  int ldc, M, N, K;
  int m_pad;        // tokens padded to a multiple of 32 (<= 512)
  int chunk;        // tokens per UMMA instruction: m_pad (<= 256) or m_pad / 2
  int stages;
  int splits;       // k-splits per weight block (pair mode: per pair of weight blocks)
  int kb_per_split;
  int cluster_k;    // 1: the splits of a weight block form a cluster and reduce through DSMEM
  float* ws;        // [n_blk][split][m_pad][128] fp32 partials
  int* counters;    // [n_blocks]
  GemmEpilogue epi;
  int dbg_skip;     // profiling aid: 1 = skip the global stores, 2 = skip the transposed read-out
  long long* dbg;   // profiling aid (VILA_B200_GEMM_DEBUG): 16 counters per CTA, or nullptr
};

__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void sts_bf16(uint32_t addr, float x) {
  const __nv_bfloat16 h = __float2bfloat16(x);
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const uint16_t*>(&h)) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(addr));
  return r;
}
// named barrier of one epilogue warpgroup (128 threads): ids 1 and 2
__device__ __forceinline__ void epi_group_sync(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}

// Transposed read-out of one staged 32-token chunk: kVpr 16-byte vectors per token row, all
// trip counts compile-time so the shared loads / residual loads / stores of a thread are batched.
template <int kVpr>
__device__ __forceinline__ void readout(const uint8_t* buf, int epi_tid, int c, int nc0, int n_out,
                                        const SkinnyArgs& a) {
  constexpr int kIters = 32 * kVpr / 128;
  const uint32_t base = smem_u32(buf);
  uint4 o[kIters];
#pragma unroll
  for (int i = 0; i < kIters; ++i) {
    const int q = epi_tid + i * 128;
    o[i] = lds_v4(base + (q / kVpr) * kStgPitch + (q % kVpr) * 16);
  }
  if (a.epi.residual != nullptr) {
    uint4 b[kIters];
#pragma unroll
    for (int i = 0; i < kIters; ++i) {
      const int q = epi_tid + i * 128;
      const int m = c * 32 + q / kVpr, nn = nc0 + (q % kVpr) * 8;
      b[i] = make_uint4(0u, 0u, 0u, 0u);
      if (m < a.M && nn < n_out) {
        const int rr = a.epi.res_row_mod > 0 ? (m % a.epi.res_row_mod) : m;
        b[i] = ldg_v4(a.epi.residual + static_cast<size_t>(rr) * a.epi.ld_res + nn);
      }
    }
#pragma unroll
    for (int i = 0; i < kIters; ++i) {
      o[i].x = pack_bf16(bf_lo(o[i].x) + bf_lo(b[i].x), bf_hi(o[i].x) + bf_hi(b[i].x));
      o[i].y = pack_bf16(bf_lo(o[i].y) + bf_lo(b[i].y), bf_hi(o[i].y) + bf_hi(b[i].y));
      o[i].z = pack_bf16(bf_lo(o[i].z) + bf_lo(b[i].z), bf_hi(o[i].z) + bf_hi(b[i].z));
      o[i].w = pack_bf16(bf_lo(o[i].w) + bf_lo(b[i].w), bf_hi(o[i].w) + bf_hi(b[i].w));
    }
  }
#pragma unroll
  for (int i = 0; i < kIters; ++i) {
    const int q = epi_tid + i * 128;
    const int m = c * 32 + q / kVpr, nn = nc0 + (q % kVpr) * 8;
    if (m < a.M && nn < n_out && a.dbg_skip != 1) stg_v4(a.C + static_cast<size_t>(m) * a.ldc + nn, o[i]);
  }
}

// Read-out of one staged 32-token chunk of ONE attention head (this CTA's 128 features) for the fused
// q/k/v projection: HF apply_rotary_pos_emb (rotate-half; cos/sin from the per-request bf16 table, every
// product rounded to bf16 before the sum — exactly rope_kv_kernel in data_movement.cu) on q and k
// heads, then q -> C, k / v -> the paged KV pool (DynamicCache.update as a scatter).
__device__ __forceinline__ void readout_rope(const uint8_t* buf, int epi_tid, int c, int head,
                                             const SkinnyArgs& a) {
  const uint32_t base = smem_u32(buf);
  const GemmEpilogue& e = a.epi;
  const bool is_q = head < e.rope_hq;
  const bool is_v = head >= e.rope_hq + e.rope_hkv;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = epi_tid + i * 128;  // 32 tokens x 8 segments of 8 rotation pairs
    const int tok = q >> 3, seg = q & 7;
    const int m = c * 32 + tok;
    if (m >= a.M) continue;
    uint4 lo = lds_v4(base + tok * kStgPitch + seg * 16);         // features [8 seg, 8 seg + 8)
    uint4 hi = lds_v4(base + tok * kStgPitch + 128 + seg * 16);   // features 64 + [8 seg, 8 seg + 8)
    if (!is_v) {
      const uint4 cv = ldg_v4(e.rope_table + static_cast<size_t>(m) * 128 + seg * 8);
      const uint4 sv = ldg_v4(e.rope_table + static_cast<size_t>(m) * 128 + 64 + seg * 8);
      uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
      uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
      const uint32_t* cw = reinterpret_cast<const uint32_t*>(&cv);
      const uint32_t* sw = reinterpret_cast<const uint32_t*>(&sv);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        float y0[2], y1[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const float x0 = hf ? bf_hi(lw[w]) : bf_lo(lw[w]);
          const float x1 = hf ? bf_hi(hw[w]) : bf_lo(hw[w]);
          const float cs = hf ? bf_hi(cw[w]) : bf_lo(cw[w]);
          const float sn = hf ? bf_hi(sw[w]) : bf_lo(sw[w]);
          y0[hf] = bf16_round(x0 * cs) + bf16_round(-x1 * sn);
          y1[hf] = bf16_round(x1 * cs) + bf16_round(x0 * sn);
        }
        lw[w] = pack_bf16(y0[0], y0[1]);
        hw[w] = pack_bf16(y1[0], y1[1]);
      }
    }
    __nv_bfloat16* dst;
    if (is_q || e.k_pool == nullptr) {
      dst = a.C + static_cast<size_t>(m) * a.ldc + head * 128;
    } else {
      const int cpos = e.cache_pos0 + m;
      const int page = e.page_table[cpos >> 7];
      const int hk = (head - e.rope_hq) % e.rope_hkv;
      dst = (is_v ? e.v_pool : e.k_pool) +
            ((static_cast<size_t>(page) * 128 + (cpos & 127)) * e.rope_hkv + hk) * 128;
    }
    stg_v4(dst + seg * 8, lo);
    stg_v4(dst + 64 + seg * 8, hi);
  }
}

template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                   SkinnyArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int x_rows = kPair ? a.m_pad / 2 : a.m_pad;  // token rows staged by THIS CTA per k-block
  const int stage_bytes = BW * BK * 2 + x_rows * BK * 2;
  uint8_t* stg = smem + a.stages * stage_bytes;
  uint8_t* bar_base = stg + kStgBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* acc_full = empty_bar + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  uint32_t* last_flag = tmem_ptr + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int unit = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int nb = unit / a.splits, split = unit - nb * a.splits;
  const int n_blk = kPair ? nb * 2 + static_cast<int>(rank) : nb;
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
  if (warp == 1) {
    if (kPair) tmem_alloc_pair<512>(tmem_ptr);
    else tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();  // peer barriers are initialised before any remote signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();

  long long* dbg = a.dbg ? a.dbg + 16 * blockIdx.x : nullptr;
  long long e0 = 0, e1 = 0;
  if (dbg && threadIdx.x == 0) {
    dbg[0] = gtime_ns();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    dbg[6] = smid;
  }
  if (warp == 0) {
    if (lane == 0) {
      long long w_empty = 0;
      // every load of the pair reports to the LEADER's full barrier
      auto full_addr = [&](int i) {
        const uint32_t own = smem_u32(&full_bar[i]);
        return kPair ? mapa_u32(own, 0) : own;
      };
      auto load = [&](void* dst, const CUtensorMap* tm, int i, int c0, int c1) {
        if (kPair) tma_load_2d_pair(dst, tm, full_addr(i), c0, c1);
        else tma_load_2d(dst, tm, &full_bar[i], c0, c1);
      };
      const uint32_t tx_bytes = static_cast<uint32_t>(stage_bytes) * (kPair ? 2u : 1u);
      // weights are parameters: their first stages are requested before the dependency wait
      const int pre = a.epi.static_w ? min(a.stages, kb1 - kb0) : 0;
      for (int i = 0; i < pre; ++i) {
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[i], tx_bytes);
        load(smem + i * stage_bytes, &tmap_w, i, (kb0 + i) * BK, n_blk * BW);
      }
      griddep_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        uint8_t* st = smem + stage * stage_bytes;
        if (kb - kb0 >= pre) {
          const long long c0 = dbg ? clock64() : 0;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (dbg) w_empty += clock64() - c0;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          load(st, &tmap_w, stage, kb * BK, n_blk * BW);
        }
        uint8_t* xs = st + BW * BK * 2;
        // one box per UMMA token chunk (rows >= M are zero-filled); a pair CTA stages its half
        const int box = kPair ? a.chunk / 2 : a.chunk;
        for (int n0 = 0, o = 0; n0 < a.m_pad; n0 += a.chunk, o += box)
          load(xs + o * 128, &tmap_x, stage, kb * BK, n0 + static_cast<int>(rank) * box);
        if (++stage == a.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (dbg) dbg[4] = w_empty;
    }
  } else if (warp == 1) {
    if (lane == 0 && kb1 > kb0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_full = 0;
      const long long m0 = dbg ? clock64() : 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const long long c0 = dbg ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase);
        if (dbg) w_full += clock64() - c0;
        tc_fence_after();
        const uint32_t wa = smem_u32(smem + stage * stage_bytes);
        const uint32_t xa = wa + BW * BK * 2;
        const uint32_t idesc = make_idesc_bf16(kPair ? 256 : 128, a.chunk, 0, 0);
        for (int n0 = 0; n0 < a.m_pad; n0 += a.chunk) {
          const uint64_t ad = make_smem_desc(wa, 16, 1024, kLayoutSW128);
          const uint64_t bd = make_smem_desc(xa + (kPair ? n0 / 2 : n0) * 128, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t acc = (kb != kb0 || k != 0) ? 1u : 0u;
            if (kPair) umma_f16_pair(tmem_base + n0, ad + 2 * k, bd + 2 * k, idesc, acc);
            else umma_f16(tmem_base + n0, ad + 2 * k, bd + 2 * k, idesc, acc);
          }
        }
        if (kPair) umma_commit_pair(&empty_bar[stage]);
        else umma_commit(&empty_bar[stage]);
        if (++stage == a.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (kPair) umma_commit_pair(acc_full);
      else umma_commit(acc_full);
      if (dbg) {
        dbg[2] = w_full;
        dbg[3] = clock64() - m0;
      }
    }
  }
  // ===================== epilogue: two warpgroups (warps 2..5 and 6..9) =====================
  // Both groups cover all four TMEM lane quadrants; they take alternate 32-token chunks, each with
  // its own staging tile and named barrier, so one group's TMEM read / activation math overlaps the
  // other group's transposed read-out.
  const int quad = warp & 3;
  const int wg = warp >= 6 ? 1 : 0;
  const int epi_tid = static_cast<int>(threadIdx.x) - 64 - wg * 128;
  const int n_local = quad * 32 + lane;
  const int n = n_blk * BW + n_local;
  const bool n_ok = n < a.N;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const int n_chunks = a.m_pad / 32;
  bool finalize = true;
  const float* slots = nullptr;
  float* dump = reinterpret_cast<float*>(smem);  // cluster split-K: [token][128] fp32 partial
  if (warp >= 2) {
    // ---- phase 1: wait for the accumulator; split-K partials leave TMEM ----
    griddep_wait();
    e0 = dbg ? clock64() : 0;
    if (kb1 > kb0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    e1 = dbg ? clock64() : 0;
    if (a.splits > 1 && a.cluster_k) {
      // All of THIS CTA's MMAs have retired -> its pipeline stages are dead; park the partial there.
      // (Peers PULL it after the cluster barrier: pushing into a peer would race with the peer's
      //  still-running main loop.)
#pragma unroll 1
      for (int c = wg; c < n_chunks; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) dump[(c * 32 + j) * BW + n_local] = __uint_as_float(r[j]);
      }
    } else if (a.splits > 1) {
      const int my_unit = n_blk * a.splits + split;
      float* mine = a.ws + (static_cast<size_t>(my_unit) * a.m_pad) * BW + n_local;
#pragma unroll 1
      for (int c = wg; c < n_chunks; c += 2) {
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
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (threadIdx.x == 64) {
        const int prev = atomicAdd(&a.counters[n_blk], 1);
        const int last = (prev == a.splits - 1) ? 1 : 0;
        if (last) a.counters[n_blk] = 0;
        *last_flag = last;
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
      finalize = (*last_flag != 0);
      if (finalize) __threadfence();
      slots = a.ws + (static_cast<size_t>(n_blk) * a.splits * a.m_pad) * BW + n_local;
    }
  }
  // cluster split-K: the `splits` CTAs of one weight block form a cluster; once every partial is
  // parked, CTA r reduces the 32-token chunks c with c % splits == r straight out of its peers'
  // shared memory (DSMEM loads) in a fixed order — no workspace round trip, no counters, deterministic
  const long long p1 = dbg ? clock64() : 0;
  if (a.cluster_k) cluster_sync_all();
  const long long p2 = dbg ? clock64() : 0;
  if (warp >= 2 && finalize) {
    // ---- phase 2: bias / activation / transposition / store ----
    const bool swiglu = a.epi.swiglu != 0;
    const int act = a.epi.act;
    const bool dsmem = a.cluster_k && a.splits > 1;
    const float bias = (a.epi.bias != nullptr && n_ok) ? __bfloat162float(a.epi.bias[n]) : 0.f;
    // output row geometry of this CTA: features [nc0, nc0 + 128 or 64)
    const int nc0 = swiglu ? (n_blk * BW) >> 1 : n_blk * BW;
    const int n_out = swiglu ? a.N >> 1 : a.N;
    const int c_first = dsmem ? split : 0, c_step = dsmem ? a.splits : 1;
    uint8_t* buf = stg + wg * (32 * kStgPitch);
#pragma unroll 1
    for (int c = c_first + wg * c_step; c < n_chunks; c += 2 * c_step) {
      if (c * 32 >= a.M) break;
      float v[32];
      if (dsmem) {
        // rank order 0..splits-1 (lowest k first); this CTA's own partial comes from local smem
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        const uint32_t off = smem_u32(dump) + static_cast<uint32_t>((c * 32) * BW + n_local) * 4u;
#pragma unroll 1
        for (int s = 0; s < a.splits; ++s) {
          if (s == split) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += dump[(c * 32 + j) * BW + n_local];
          } else {
            const uint32_t base = mapa_u32(off, static_cast<uint32_t>(s));
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += ld_dsmem_f32(base + j * (BW * 4));
          }
        }
      } else if (a.splits > 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll 1
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
      if (swiglu) {
        // lanes (2i, 2i+1) hold (gate_i, up_i) of the interleaved weight rows.  The pair splits the
        // 32 tokens: the even (gate) lane finishes tokens 0..15, the odd (up) lane tokens 16..31, so
        // every lane does useful SiLU work and one shuffle serves two tokens.
        const bool odd = (lane & 1) != 0;
        const uint32_t dst = smem_u32(buf) + ((n_local >> 1) + (odd ? 16 * (kStgPitch / 2) : 0)) * 2;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a0 = bf16_round(v[j] + bias);       // this lane's row, token j
          const float a1 = bf16_round(v[16 + j] + bias);  // this lane's row, token 16 + j
          const float recv = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
          const float gate = odd ? recv : a0;
          const float up = odd ? a1 : recv;
          sts_bf16(dst + j * kStgPitch, bf16_round(silu_f(gate)) * up);
        }
      } else {
        const uint32_t dst = smem_u32(buf) + n_local * 2;
        if (act == ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sts_bf16(dst + j * kStgPitch, v[j] + bias);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float xb = bf16_round(v[j] + bias);
            const float y = act == ACT_GELU_TANH ? gelu_tanh_f(xb)
                                                 : (act == ACT_GELU_ERF ? gelu_erf_f(xb) : silu_f(xb));
            sts_bf16(dst + j * kStgPitch, y);
          }
        }
      }
      epi_group_sync(wg);
      // transposed read-out: 16-byte vectors along the feature dimension
      if (a.dbg_skip != 2) {
        if (a.epi.rope_table != nullptr) readout_rope(buf, epi_tid, c, n_blk, a);
        else if (swiglu) readout<8>(buf, epi_tid, c, nc0, n_out, a);
        else readout<16>(buf, epi_tid, c, nc0, n_out, a);
      }
      epi_group_sync(wg);  // the staging tile may be overwritten
    }
  }

  if (dbg && threadIdx.x == 64) {
    dbg[8] = p1 - e1;            // phase 1 (partial leaves TMEM)
    dbg[9] = p2 - p1;            // cluster barrier
    dbg[7] = e1 - e0;            // epilogue warps waiting for the accumulator
    dbg[5] = clock64() - e1;     // epilogue proper
  }
  tc_fence_before();
  // pair: the leader's MMAs read the peer's shared memory until the end; cluster split-K: peers read
  // this CTA's parked partial until they are done
  if (kPair || a.cluster_k) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
  if (dbg && threadIdx.x == 0) dbg[1] = gtime_ns();
}

template <bool kPair>
int launch_skinny(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, __nv_bfloat16* C,
                  int ldc, int M, int N, int K, const GemmEpilogue& epi, cudaStream_t stream) {
  const int m_pad = (M + 31) / 32 * 32;
  const int sms = num_sms();
  const int n_blocks = (N + BW - 1) / BW;
  const int nb_units = kPair ? (n_blocks + 1) / 2 : n_blocks;  // scheduling units along N
  const int slots = kPair ? sms / 2 : sms;
  const int nkb = (K + BK - 1) / BK;
  int splits = 1;
  if (nb_units < slots) {
    splits = (slots + nb_units / 2) / nb_units;
    if (splits > 8) splits = 8;
    while (splits > 1 && nkb / splits < 4) --splits;
  }
  int kb_per_split = (nkb + splits - 1) / splits;
  splits = (nkb + kb_per_split - 1) / kb_per_split;  // no empty splits
  const int want_splits = splits;
  void* ws_ptr = nullptr;
  size_t ws_bytes = 0;
  get_workspace(&ws_ptr, &ws_bytes);
  const int n_blocks_alloc = kPair ? nb_units * 2 : n_blocks;
  const size_t need = kCounterBytes + static_cast<size_t>(n_blocks_alloc) * splits * m_pad * BW * 4;
  if (splits > 1 && (ws_ptr == nullptr || ws_bytes < need || n_blocks_alloc > 16384)) {
    splits = 1;
    kb_per_split = nkb;
  }
  const int x_rows = kPair ? m_pad / 2 : m_pad;
  const int stage_bytes = BW * BK * 2 + x_rows * BK * 2;
  const int budget = 227 * 1024 - 1024 - kStgBytes - 256;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return -1;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + kStgBytes + 256 + 1024;
  auto kern = gemm_skinny_kernel<kPair>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  // cluster split-K (single-CTA flavour): largest split count whose clusters are all co-resident
  int cluster_k = 0;
  if (!kPair && want_splits > 1) {
    static int active[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
    for (int sp = want_splits; sp >= 2; --sp) {
      const int kbps = (nkb + sp - 1) / sp;
      if ((nkb + kbps - 1) / kbps != sp) continue;
      // the fp32 partial (m_pad x 128) is parked in the dead pipeline stages
      if (static_cast<size_t>(m_pad) * BW * 4 > static_cast<size_t>(stages) * stage_bytes) continue;
      if (active[sp] < 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sp * 64);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = sp;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n_act = 0;
        if (cudaOccupancyMaxActiveClusters(&n_act, kern, &cfg) != cudaSuccess) {
          (void)cudaGetLastError();
          n_act = 0;
        }
        active[sp] = n_act;
      }
      if (active[sp] >= n_blocks) {
        splits = sp;
        kb_per_split = kbps;
        cluster_k = 1;
        break;
      }
    }
  }
  SkinnyArgs a;
  a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.cluster_k = cluster_k;
  a.m_pad = m_pad;
  a.chunk = m_pad > 256 ? m_pad / 2 : m_pad;  // multiple of 16; pair halves are multiples of 8 rows
  a.splits = splits;
  a.kb_per_split = kb_per_split;
  a.counters = static_cast<int*>(ws_ptr);
  a.ws = ws_ptr ? reinterpret_cast<float*>(static_cast<char*>(ws_ptr) + kCounterBytes) : nullptr;
  a.epi = epi;
  a.dbg = nullptr;
  a.dbg_skip = 0;
#ifdef VILA_B200_DEBUG_HOOKS  // profiling aids, compiled OUT of the shipped library (an env var
                              // must not be able to make a kernel write through an arbitrary pointer)
  if (const char* e = getenv("VILA_B200_GEMM_DEBUG_SKIP")) a.dbg_skip = atoi(e);
  if (const char* e = getenv("VILA_B200_GEMM_DEBUG")) {  // hex device pointer (tools/skinny_phase_times.py)
    unsigned long long ptr = 0;
    if (sscanf(e, "%llx", &ptr) == 1) a.dbg = reinterpret_cast<long long*>(ptr);
  }
#endif
  a.stages = stages;
  CUtensorMap tw, tx;
  if (make_tmap_2d_bf16(&tw, W, N, K, ldw, BW, BK, 128)) return 1;
  if (make_tmap_2d_bf16(&tx, A, M, K, lda, kPair ? a.chunk / 2 : a.chunk, BK, 128)) return 1;
  const int grid = nb_units * splits * (kPair ? 2 : 1);
  VB_CUDA(launch_pdl_cluster(kern, dim3(grid), dim3(kThreads), smem, stream,
                             dim3(kPair ? 2 : (cluster_k ? splits : 1), 1, 1), tw, tx, a));
  return 0;
}

}  // namespace

// returns -1 if the shape is not handled here.  pair: 0 = one CTA per 128 weight rows,
// 1 = CTA pairs (tcgen05 cta_group::2) per 256 weight rows
int gemm_skinny_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, __nv_bfloat16* C,
                     int ldc, int M, int N, int K, const GemmEpilogue& epi, int pair, cudaStream_t stream) {
  if (M > 512 || M < 1) return -1;
  if (epi.rope_table != nullptr) {
    VB_CHECK(N == (epi.rope_hq + 2 * epi.rope_hkv) * 128 && !epi.swiglu &&
                 epi.act == ACT_NONE && epi.residual == nullptr &&
                 (epi.k_pool == nullptr || (epi.v_pool != nullptr && epi.page_table != nullptr)),
             "gemm_skinny: the fused q/k/v RoPE epilogue needs head_dim 128, N = (Hq + 2 Hkv) * 128 and "
             "no other epilogue");
  }
  return pair ? launch_skinny<true>(A, lda, W, ldw, C, ldc, M, N, K, epi, stream)
              : launch_skinny<false>(A, lda, W, ldw, C, ldc, M, N, K, epi, stream);
}

}  // namespace vb
