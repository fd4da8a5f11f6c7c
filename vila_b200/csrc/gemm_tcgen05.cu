// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epi(A[M,K] · W[N,K]^T)
//
//   * operands staged global -> shared by TMA (cp.async.bulk.tensor, 128B swizzle, K-major)
//   * tcgen05.mma (cta_group::1, kind::f16, 128 x BLOCK_N x 16) issued by ONE thread
//   * fp32 accumulators live in TMEM, double-buffered so the epilogue of one work segment overlaps
//     the main loop of the next
//   * epilogue warps read TMEM with tcgen05.ld (lane == output row) and fuse
//     bias / GELU / SwiGLU / residual (+ broadcast "row modulo" residual for position embeddings)
//   * scheduling: data-parallel over output tiles, or STREAM-K when the tile count does not fill the
//     148 SMs evenly (skinny prefill GEMMs, M = 280): the (tile, k-block) iteration space is cut
//     into equal contiguous ranges, each partial tile is parked in its own fp32 workspace slot and
//     the last-arriving CTA sums the slots in a fixed order (deterministic) and applies the epilogue
//   * programmatic dependent launch: barrier init / TMEM alloc / descriptor prefetch and — for
//     parameter matrices — the first pipeline stages of W overlap the predecessor kernel's tail
//
// Replaces the cuBLAS calls behind nn.Linear on the reference hot path:
//   SigLIP q/k/v/out_proj, fc1/fc2      (modeling_siglip.py:384-387,707-715)
//   mm_projector Linear layers          (base_projector.py:145-162)
//   Qwen2 q/k/v/o, gate/up/down, lm_head (modeling_qwen2.py:164-176,223-226)
//   patch-embed conv as im2col GEMM     (modeling_siglip.py:269-275,322-328)
#include "common.cuh"
#include "kernels.h"

namespace vb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one 128B-swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 192;  // warp0: TMA, warp1: MMA + TMEM alloc, warps 2..5: epilogue

// kPair: a 2-CTA cluster computes one (2*BLOCK_M) x BLOCK_N tile with tcgen05.mma.cta_group::2 —
// each CTA stages its own 128 rows of A and HALF of the W tile (BLOCK_N/2 rows); the tensor cores of
// both SMs read both halves, so the bytes each SM pulls through its L2 port per MAC drop by 1/3
// (128x256 tile: 48 KB -> 32 KB per k-block), which is what bounds the single-CTA kernel.
//
// kMode 2 (split-K pair): a 2-CTA cluster computes one 128 x BLOCK_N tile, each CTA half of the K
// range with ordinary cta_group::1 MMAs; the two fp32 partials are exchanged through distributed
// shared memory (each CTA parks the half of the columns it does NOT finalise, the peer pulls it),
// so under-filled GEMMs (N = 1152 ViT out_proj / fc2: 72 tiles) use all 148 SMs.  One tile per pair.
enum { kModeSingle = 0, kModePair = 1, kModeSplitK2 = 2 };

template <int BLOCK_N, int kStages, int kMode = kModeSingle>
struct GemmSmem {
  static constexpr bool kPair = kMode == kModePair;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBRows = kPair ? BLOCK_N / 2 : BLOCK_N;  // W rows staged by one CTA
  static constexpr int kBBytes = kBRows * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;  // + align slack
};


__device__ __forceinline__ float4 ld_cg_v4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// Tile rasterisation: tiles are numbered so that the CTAs working at the same time share operand
// panels in L2.  Plain M-fastest numbering streams the whole A matrix once per N block (ncu, ViT qkv
// at M = 65536: 2.5 GB of DRAM traffic for 0.61 GB of operands + output).  Tiles are therefore grouped
// in bands of kBandM M-blocks; inside a band M runs fastest, then N: the ~74-148 concurrent tiles
// touch kBandM A panels and a sliding window of W panels.  With <= kBandM M-blocks (every prefill
// shape up to 2048 tokens) this IS the M-fastest order.
constexpr int kBandM = 8;
__device__ __forceinline__ void tile_coords(int t, int num_m_blocks, int num_n_blocks, int& m_blk, int& n_blk) {
  const int per_band = kBandM * num_n_blocks;
  const int band = t / per_band;
  const int r = t - band * per_band;
  const int m0 = band * kBandM;
  const int bm = min(kBandM, num_m_blocks - m0);
  n_blk = r / bm;
  m_blk = m0 + (r - n_blk * bm);
}

// Work decomposition shared by the three warp roles: a CTA walks a sequence of segments
// (tile, [kb0, kb1)).  Data-parallel: whole tiles blockIdx.x, +gridDim.x, ...  Stream-K: the
// contiguous iteration range [it0, it1) of the (tile, k-block) space.
struct Sched {
  int nkb, num_m_blocks, num_tiles;
  int stream_k;
  long it, it_end;  // stream-K
  int tile;         // data-parallel
  int stride;  // data-parallel: tiles between two visits of this CTA (or CTA pair)
  // tile_m: rows of one scheduling tile (128, or 256 for a CTA pair); worker/n_workers: index and
  // number of the units that walk the tile list (CTAs, or CTA pairs)
  __device__ __forceinline__ Sched(int M, int N, int K, int block_n, int stream_k_, int tile_m = BLOCK_M,
                                   int worker = blockIdx.x, int n_workers = gridDim.x) {
    nkb = (K + BLOCK_K - 1) / BLOCK_K;
    num_m_blocks = (M + tile_m - 1) / tile_m;
    num_tiles = num_m_blocks * ((N + block_n - 1) / block_n);
    stream_k = stream_k_;
    const long total = static_cast<long>(num_tiles) * nkb;
    it = total * worker / n_workers;
    it_end = total * (worker + 1) / n_workers;
    tile = worker;
    stride = n_workers;
  }
  // returns false when done; otherwise the next segment
  __device__ __forceinline__ bool next(int& t, int& kb0, int& kb1) {
    if (stream_k) {
      if (it >= it_end) return false;
      t = static_cast<int>(it / nkb);
      kb0 = static_cast<int>(it - static_cast<long>(t) * nkb);
      const long rem = it_end - it;
      kb1 = (nkb - kb0) < rem ? nkb : kb0 + static_cast<int>(rem);
      it += kb1 - kb0;
      return true;
    }
    if (tile >= num_tiles) return false;
    t = tile;
    kb0 = 0;
    kb1 = nkb;
    tile += stride;
    return true;
  }
};

template <int BLOCK_N, int kStages, int kMode>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_w, __nv_bfloat16* __restrict__ C,
                         int ldc, int M, int N, int K, GemmEpilogue epi) {
  using S = GemmSmem<BLOCK_N, kStages, kMode>;
  constexpr bool kPair = kMode == kModePair;
  constexpr bool kSplit2 = kMode == kModeSplitK2;
  constexpr bool kCluster = kPair || kSplit2;
  constexpr int TILE_M = kPair ? 2 * BLOCK_M : BLOCK_M;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint32_t* last_flag = tmem_ptr + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m_blocks = (M + TILE_M - 1) / TILE_M;
  const int num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
  const int stream_k = (!kCluster && epi.split_k > 1) ? 1 : 0;
  const uint32_t rank = kCluster ? cluster_ctarank() : 0u;        // CTA within the pair
  const uint32_t mrank = kPair ? rank : 0u;                       // pair mode: which half of the operands
  const int worker = kCluster ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_workers = kCluster ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // split-K pair: this CTA's k-block range of every tile
  const int nkb_all = (K + BLOCK_K - 1) / BLOCK_K;
  const int sk_lo = (kSplit2 && rank == 1) ? (nkb_all + 1) / 2 : 0;
  const int sk_hi = (kSplit2 && rank == 0) ? (nkb_all + 1) / 2 : nkb_all;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    // pair: one arrival per epilogue warp of BOTH CTAs, on the leader's barrier
    mbar_init(&tmem_empty[0], kPair ? 8 : 128);
    mbar_init(&tmem_empty[1], kPair ? 8 : 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_pair<2 * BLOCK_N>(tmem_ptr);
    else tmem_alloc<2 * BLOCK_N>(tmem_ptr);
  }
  tc_fence_before();
  if (kCluster) cluster_sync_all();  // the peer's barriers exist before any remote signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      Sched sch(M, N, K, BLOCK_N, stream_k, TILE_M, worker, n_workers);
      // every load of a pair reports to the LEADER's full barrier, which expects both CTAs' bytes
      auto load = [&](void* dst, const CUtensorMap* tm, int i, int c0, int c1) {
        if (kPair) tma_load_2d_pair(dst, tm, mapa_u32(smem_u32(&full_bar[i]), 0), c0, c1);
        else tma_load_2d(dst, tm, &full_bar[i], c0, c1);
      };
      constexpr uint32_t kTxBytes = S::kStageBytes * (kPair ? 2 : 1);
      const int a_row_off = static_cast<int>(mrank) * BLOCK_M;
      const int w_row_off = static_cast<int>(mrank) * S::kBRows;
      int stage = 0;
      uint32_t phase = 0;
      int t, kb0, kb1;
      bool waited = false;
      // W is a parameter: fetch its first stages before the dependency wait (A comes after)
      int pre = 0;
      int pt = 0, pkb0 = 0, pkb1 = 0;
      Sched peek = sch;
      const bool have_first = peek.next(pt, pkb0, pkb1);
      if (kSplit2) {
        pkb0 = sk_lo;
        pkb1 = sk_hi;
      }
      if (epi.static_w && have_first) {
        pre = min(kStages, pkb1 - pkb0);
        int m_blk_unused, n_blk;
        tile_coords(pt, num_m_blocks, num_n_blocks, m_blk_unused, n_blk);
        for (int i = 0; i < pre; ++i) {
          if (mrank == 0) mbar_arrive_expect_tx(&full_bar[i], kTxBytes);
          load(smem_b + i * S::kBBytes, &tmap_w, i, (pkb0 + i) * BLOCK_K, n_blk * BLOCK_N + w_row_off);
        }
      }
      griddep_wait();
      waited = true;
      (void)waited;
      bool first = true;
      while (sch.next(t, kb0, kb1)) {
        int m_blk, n_blk;
        tile_coords(t, num_m_blocks, num_n_blocks, m_blk, n_blk);
        if (kSplit2) {
          kb0 = sk_lo;
          kb1 = sk_hi;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          if (first && (kb - kb0) < pre) {
            // W already in flight for this stage: only A is missing
            load(smem_a + stage * S::kABytes, &tmap_a, stage, kb * BLOCK_K, m_blk * TILE_M + a_row_off);
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (mrank == 0) mbar_arrive_expect_tx(&full_bar[stage], kTxBytes);
            load(smem_a + stage * S::kABytes, &tmap_a, stage, kb * BLOCK_K, m_blk * TILE_M + a_row_off);
            load(smem_b + stage * S::kBBytes, &tmap_w, stage, kb * BLOCK_K, n_blk * BLOCK_N + w_row_off);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0 && mrank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BLOCK_N, 0, 0);
      Sched sch(M, N, K, BLOCK_N, stream_k, TILE_M, worker, n_workers);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int t, kb0, kb1;
      while (sch.next(t, kb0, kb1)) {
        if (kSplit2) {
          kb0 = sk_lo;
          kb1 = sk_hi;
        }
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc =
              make_smem_desc(smem_u32(smem_a + stage * S::kABytes), 16, 1024, kLayoutSW128);
          const uint64_t b_desc =
              make_smem_desc(smem_u32(smem_b + stage * S::kBBytes), 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the 128B swizzle atom: +2 in addr>>4 units
            const uint32_t acc = (kb != kb0 || k != 0) ? 1u : 0u;
            if (kPair) umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
            else umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
          }
          // frees this smem stage (in both CTAs of a pair) when the MMAs retire
          if (kPair) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (kPair) umma_commit_pair(&tmem_full[as]);  // accumulator complete -> epilogue(s)
        else umma_commit(&tmem_full[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps (TMEM -> regs -> global) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int epi_tid = threadIdx.x - 64;
    Sched sch(M, N, K, BLOCK_N, stream_k, TILE_M, worker, n_workers);
    int as = 0;
    uint32_t aphase = 0;
    int t, kb0, kb1;
    griddep_wait();  // C / residual / workspace may still be in use by the predecessor
    while (sch.next(t, kb0, kb1)) {
      int m_blk, n_blk;
      tile_coords(t, num_m_blocks, num_n_blocks, m_blk, n_blk);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row = m_blk * TILE_M + static_cast<int>(mrank) * BLOCK_M + quad * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N;
      const bool partial = !kSplit2 && ((kb0 != 0) || (kb1 != sch.nkb));
      // split-K pair: park the column half the PEER finalises ([col/4][row] float4 -> conflict-free
      // for the writer here and for the peer's DSMEM reads), then meet the peer at the cluster barrier
      constexpr int kHalfChunks = BLOCK_N / 64;  // 32-column chunks per half
      float4* xchg = reinterpret_cast<float4*>(smem_a);  // the pipeline stages are dead by now
      if (kSplit2) {
        const int c_peer = static_cast<int>(1u - rank) * kHalfChunks;
#pragma unroll 1
        for (int c = 0; c < kHalfChunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + (c_peer + c) * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            xchg[(c * 8 + g) * BLOCK_M + quad * 32 + lane] =
                make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                            __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
        }
        cluster_sync_all();  // (warps 0/1 execute the matching barrier after their loops)
      }
      const uint32_t xchg_peer = kSplit2 ? mapa_u32(smem_u32(xchg), 1u - rank) : 0u;
      bool finalize = true;       // apply the epilogue and write C
      bool from_ws = false;       // accumulator comes from the fp32 workspace
      int n_slots = 0;
      const float* slot0 = nullptr;  // first partial slot of this tile, this thread's row
      if (partial) {
        // ---- stream-K partial tile: park the fp32 partial in this segment's own workspace slot
        // (slots are summed later in a fixed order -> deterministic, no atomics on data) ----
        const long total = static_cast<long>(sch.num_tiles) * sch.nkb;
        const long i0 = static_cast<long>(t) * sch.nkb;
        auto cta_of = [&](long i) {
          long g = i * gridDim.x / total;
          while (g + 1 < static_cast<long>(gridDim.x) && total * (g + 1) / gridDim.x <= i) ++g;
          return static_cast<int>(g);
        };
        const int c_first = cta_of(i0);
        n_slots = cta_of(i0 + sch.nkb - 1) - c_first + 1;
        const int my_slot = static_cast<int>(blockIdx.x) - c_first;
        const size_t tile_base = static_cast<size_t>(t) * epi.split_k * (BLOCK_M * BLOCK_N);
        const size_t row_off = static_cast<size_t>(quad * 32 + lane) * BLOCK_N;
        slot0 = epi.splitk_ws + tile_base + row_off;
        float* mine = epi.splitk_ws + tile_base + static_cast<size_t>(my_slot) * (BLOCK_M * BLOCK_N) + row_off;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<uint4*>(mine + c * 32 + g * 4) =
                  make_uint4(r[g * 4 + 0], r[g * 4 + 1], r[g * 4 + 2], r[g * 4 + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[as]);  // TMEM stage is free again
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_tid == 0) {
          const int add = kb1 - kb0;
          const int prev = atomicAdd(&epi.splitk_counters[t], add);
          const int last = (prev + add == sch.nkb) ? 1 : 0;
          if (last) epi.splitk_counters[t] = 0;  // self-cleaning
          *last_flag = last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        finalize = (*last_flag != 0);
        from_ws = true;
        if (finalize) __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone has read last_flag
      }
      if (finalize) {
        const __nv_bfloat16* res_row = nullptr;
        if (epi.residual != nullptr && row_ok) {
          const int rr = epi.res_row_mod > 0 ? (row % epi.res_row_mod) : row;
          res_row = epi.residual + static_cast<size_t>(rr) * epi.ld_res;
        }
        const int c_lo = kSplit2 ? static_cast<int>(rank) * kHalfChunks : 0;
        const int c_hi = kSplit2 ? c_lo + kHalfChunks : BLOCK_N / 32;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          float v[32];
          const int n0 = n_blk * BLOCK_N + c * 32;
          if (!from_ws) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (kSplit2) {
              // + the peer's partial of the same columns (two operands: order-independent)
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 q = ld_dsmem_v4f(
                    xchg_peer + static_cast<uint32_t>(((c - c_lo) * 8 + g) * BLOCK_M + quad * 32 + lane) * 16u);
                v[4 * g + 0] += q.x; v[4 * g + 1] += q.y; v[4 * g + 2] += q.z; v[4 * g + 3] += q.w;
              }
            }
          } else {
            if (n0 >= N) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            if (row_ok) {
              for (int s = 0; s < n_slots; ++s) {  // fixed order: slot 0 (lowest k) first
                const float* wsp = slot0 + static_cast<size_t>(s) * (BLOCK_M * BLOCK_N) + c * 32;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  const float4 q = ld_cg_v4(wsp + g * 4);
                  v[g * 4 + 0] += q.x; v[g * 4 + 1] += q.y; v[g * 4 + 2] += q.z; v[g * 4 + 3] += q.w;
                }
              }
            }
          }
          if (n0 >= N) continue;
          if (epi.bias != nullptr) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (n0 + g * 8 < N) {
                const uint4 b = ldg_v4(epi.bias + n0 + g * 8);
                v[g * 8 + 0] += bf_lo(b.x);
                v[g * 8 + 1] += bf_hi(b.x);
                v[g * 8 + 2] += bf_lo(b.y);
                v[g * 8 + 3] += bf_hi(b.y);
                v[g * 8 + 4] += bf_lo(b.z);
                v[g * 8 + 5] += bf_hi(b.z);
                v[g * 8 + 6] += bf_lo(b.w);
                v[g * 8 + 7] += bf_hi(b.w);
              }
            }
          }
          if (epi.swiglu) {
            // interleaved (gate, up) column pairs -> 16 outputs: silu(gate) * up,
            // rounding to bf16 at the points the reference's unfused ops do.
            if (row_ok) {
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float g0 = bf16_round(v[4 * j + 0]), u0 = bf16_round(v[4 * j + 1]);
                float g1 = bf16_round(v[4 * j + 2]), u1 = bf16_round(v[4 * j + 3]);
                float a0 = bf16_round(silu_f(g0)) * u0;
                float a1 = bf16_round(silu_f(g1)) * u1;
                o[j] = pack_bf16(a0, a1);
              }
              __nv_bfloat16* dst = C + static_cast<size_t>(row) * ldc + (n0 >> 1);
              if (n0 + 16 <= N) stg_v4(dst, make_uint4(o[0], o[1], o[2], o[3]));
              if (n0 + 32 <= N) stg_v4(dst + 8, make_uint4(o[4], o[5], o[6], o[7]));
            }
            continue;
          }
          if (epi.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = bf16_round(v[j]);
              v[j] = epi.act == ACT_GELU_TANH
                         ? gelu_tanh_f(x)
                         : (epi.act == ACT_GELU_ERF ? gelu_erf_f(x) : silu_f(x));
            }
          }
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (n0 + g * 8 < N) {
                float* w = v + g * 8;
                if (res_row != nullptr) {
                  const uint4 b = ldg_v4(res_row + n0 + g * 8);
                  w[0] = bf16_round(w[0]) + bf_lo(b.x);
                  w[1] = bf16_round(w[1]) + bf_hi(b.x);
                  w[2] = bf16_round(w[2]) + bf_lo(b.y);
                  w[3] = bf16_round(w[3]) + bf_hi(b.y);
                  w[4] = bf16_round(w[4]) + bf_lo(b.z);
                  w[5] = bf16_round(w[5]) + bf_hi(b.z);
                  w[6] = bf16_round(w[6]) + bf_lo(b.w);
                  w[7] = bf16_round(w[7]) + bf_hi(b.w);
                }
                uint4 o;
                o.x = pack_bf16(w[0], w[1]);
                o.y = pack_bf16(w[2], w[3]);
                o.z = pack_bf16(w[4], w[5]);
                o.w = pack_bf16(w[6], w[7]);
                stg_v4(C + static_cast<size_t>(row) * ldc + n0 + g * 8, o);
              }
            }
          }
        }
      }
      if (!partial) {
        tc_fence_before();
        if (kPair) {
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));
        } else {
          mbar_arrive(&tmem_empty[as]);
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  if (kSplit2 && warp < 2) cluster_sync_all();  // matches the epilogue warps' exchange barrier
  // pair: the leader's MMAs read the peer's shared memory until the end; split-K pair: the peer
  // pulls this CTA's parked partial
  if (kCluster) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair<2 * BLOCK_N>(tmem_base);
    else tmem_dealloc<2 * BLOCK_N>(tmem_base);
  }
}

// caller-registered scratch for stream-K partial sums (vila_set_workspace)
struct Workspace {
  void* ptr = nullptr;
  size_t bytes = 0;
};
// one registration per DEVICE (a process may drive several GPUs; the scratch is device memory)
Workspace g_ws_dev[64];
inline Workspace& cur_ws() {
  int d = 0;
  cudaGetDevice(&d);
  return g_ws_dev[(d < 0 || d >= 64) ? 0 : d];
}
#define g_ws cur_ws()
constexpr size_t kCounterBytes = 64 * 1024;  // 16384 tile counters

template <int BLOCK_N, int kStages, int kMode = kModeSingle>
int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, __nv_bfloat16* C,
                int ldc, int M, int N, int K, GemmEpilogue epi, int force_stream_k,
                cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N, kStages, kMode>;
  constexpr bool kPair = kMode == kModePair;
  constexpr int TILE_M = kPair ? 2 * BLOCK_M : BLOCK_M;
  CUtensorMap ta, tw;
  if (make_tmap_2d_bf16(&ta, A, M, K, lda, BLOCK_M, BLOCK_K, 128)) return 1;
  if (make_tmap_2d_bf16(&tw, W, N, K, ldw, S::kBRows, BLOCK_K, 128)) return 1;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, kStages, kMode>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  }
  const int sms = num_sms();
  const int tiles = ((M + TILE_M - 1) / TILE_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  if (kMode == kModeSplitK2) {
    // one 128 x BLOCK_N tile per CTA pair, half of K each
    if (tiles > sms / 2 || (K + BLOCK_K - 1) / BLOCK_K < 2) {
      set_last_error("gemm: split-K pairs need <= %d tiles and >= 2 k-blocks (tiles=%d)", sms / 2, tiles);
      return 1;
    }
    epi.split_k = 1;
    VB_CUDA(launch_pdl_cluster(kern, dim3(2 * tiles), dim3(kNumThreads), S::kTotal, stream,
                               dim3(2, 1, 1), ta, tw, C, ldc, M, N, K, epi));
    return 0;
  }
  if (kPair) {
    // data-parallel over 256 x BLOCK_N tiles, one tile at a time per CTA pair
    const int pairs = tiles < sms / 2 ? tiles : sms / 2;
    epi.split_k = 1;
    VB_CUDA(launch_pdl_cluster(kern, dim3(2 * pairs), dim3(kNumThreads), S::kTotal, stream,
                               dim3(2, 1, 1), ta, tw, C, ldc, M, N, K, epi));
    return 0;
  }
  const int nkb = (K + BLOCK_K - 1) / BLOCK_K;
  // stream-K when whole tiles would leave SMs idle or produce a ragged last wave
  bool sk = false;
  const long total_it = static_cast<long>(tiles) * nkb;
  const int sk_grid = total_it < sms ? static_cast<int>(total_it) : sms;
  const long ipc = total_it / sk_grid;  // k-iterations per CTA (floor)
  const int max_slots = static_cast<int>((nkb + ipc - 1) / ipc) + 1;
  const size_t need = kCounterBytes + static_cast<size_t>(tiles) * max_slots * BLOCK_M * BLOCK_N * 4;
  const bool ws_ok = g_ws.ptr != nullptr && g_ws.bytes >= need && tiles <= 16384;
  if (force_stream_k >= 0) {
    sk = force_stream_k != 0;
  }
  // (stream-K stays opt-in: on the NVILA shapes the partial-tile fix-up costs more than the SM
  //  under-fill it removes — see profiles/r01_gemm_configs.md)
  if (sk && !ws_ok) {
    set_last_error("gemm: stream-K needs a registered workspace of >= %zu bytes (vila_set_workspace)",
                   need);
    return 1;
  }
  int grid = tiles < sms ? tiles : sms;
  epi.split_k = 1;
  if (sk) {
    grid = sk_grid;
    epi.split_k = max_slots;  // > 1 selects the stream-K schedule; = workspace slots per tile
    epi.splitk_counters = static_cast<int*>(g_ws.ptr);
    epi.splitk_ws = reinterpret_cast<float*>(static_cast<char*>(g_ws.ptr) + kCounterBytes);
  }
  VB_CUDA(launch_pdl(kern, dim3(grid), dim3(kNumThreads), S::kTotal, stream, ta, tw, C, ldc, M, N,
                     K, epi));
  return 0;
}

int check_args(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, __nv_bfloat16* C,
               int ldc, int M, int N, int K, const GemmEpilogue& epi) {
  VB_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VB_CHECK(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0,
           "gemm: K, lda, ldw must be multiples of 8 (TMA 16-byte strides): K=%d lda=%d ldw=%d", K,
           lda, ldw);
  VB_CHECK(N % 8 == 0 && ldc % 8 == 0, "gemm: N and ldc must be multiples of 8: N=%d ldc=%d", N,
           ldc);
  VB_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(C) & 15) == 0,
           "gemm: pointers must be 16-byte aligned");
  if (epi.swiglu) VB_CHECK(N % 32 == 0, "gemm: swiglu epilogue needs N %% 32 == 0 (N=%d)", N);
  return 0;
}

}  // namespace

// q/k/v projection with RoPE + KV-cache append fused into the epilogue (swap-AB kernel, M <= 384,
// head_dim 128).  Returns -1 when the shape is not covered (caller: plain GEMM + rope_kv_append).
int gemm_qkv_rope_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, __nv_bfloat16* C,
                       int ldc, int M, int N, int K, const GemmEpilogue& epi, cudaStream_t stream) {
  if (check_args(A, lda, W, ldw, C, ldc, M, N, K, epi)) return 1;
  if (M > 384) return -1;
  return gemm_skinny_bf16(A, lda, W, ldw, C, ldc, M, N, K, epi, 0, stream);
}

void get_workspace(void** ptr, size_t* bytes) {
  *ptr = g_ws.ptr;
  *bytes = g_ws.bytes;
}

int set_workspace(void* ptr, size_t bytes) {
  VB_CHECK(ptr == nullptr || bytes >= kCounterBytes + 1024, "workspace too small (%zu bytes)", bytes);
  VB_CHECK((reinterpret_cast<uintptr_t>(ptr) & 255) == 0, "workspace must be 256-byte aligned");
  g_ws.ptr = ptr;
  g_ws.bytes = ptr ? bytes : 0;
  return 0;
}

int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, __nv_bfloat16* C,
              int ldc, int M, int N, int K, const GemmEpilogue& epi, cudaStream_t stream) {
  if (check_args(A, lda, W, ldw, C, ldc, M, N, K, epi)) return 1;
  VB_CHECK(epi.rope_table == nullptr, "gemm: the fused RoPE epilogue is only available through gemm_qkv_rope_bf16");
  // Tile-shape heuristic. N=256 tiles keep the MMA (not shared-memory bandwidth) the limiter when
  // there is enough work; N=128 otherwise; stream-K (inside launch_gemm) fixes SM under-fill.
  // Measured on B200 (tools/bench_gemm.py): 128x256 tiles win as soon as they fill ~70 % of the
  // SMs (the 256-wide tile keeps the MMA, not shared-memory bandwidth, the limiter); otherwise
  // 128x128; 64-wide tiles only for narrow outputs.
  // Few tokens (short-prompt prefill, projector): the weights dominate the traffic -> swap-AB kernel
  // (every weight byte enters an SM once; cluster split-K fills the machine).  B200, M = 279:
  // down-proj 52 vs 90 us, gate/up 83 vs 86, qkv/o 21-23 vs 22 (profiles/r01_gemm_configs.md)
  const int sms = num_sms();
  const int mb = (M + BLOCK_M - 1) / BLOCK_M;
  // Long-K GEMMs with at most one 128x128 tile per SM pair (ViT fc2: 72 tiles, K = 4304; projector):
  // split K over a CTA pair with a DSMEM exchange -> all SMs stream operands (fc2 16.9 vs 24.7 us).
  const long tiles128 = static_cast<long>(mb) * ((N + 127) / 128);
  const bool split2 = tiles128 <= sms / 2 && K >= 2048 && M * 10 >= mb * BLOCK_M * 9;
  if (split2) return launch_gemm<128, 6, kModeSplitK2>(A, lda, W, ldw, C, ldc, M, N, K, epi, -1, stream);
  if (M <= 384) {
    const int rc = gemm_skinny_bf16(A, lda, W, ldw, C, ldc, M, N, K, epi, 0, stream);
    if (rc >= 0) return rc;
  }
  // CTA pairs (256 x 256 tiles, cta_group::2) once they fill ~85 % of the SM pairs: fewer bytes per
  // MAC through each SM's L2 port and shared memory (8192^3: 752 vs 859 us; 2048x37888x3584: 355 vs 396)
  const long pair_tiles = static_cast<long>((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((N + 255) / 256);
  const long m_pad_pair = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) * (2L * BLOCK_M), m_pad_single = static_cast<long>(mb) * BLOCK_M;
  if (pair_tiles * 20 >= 17L * (sms / 2) && m_pad_pair * 8 <= m_pad_single * 9)  // no extra row padding
    return launch_gemm<256, 6, kModePair>(A, lda, W, ldw, C, ldc, M, N, K, epi, -1, stream);
  const long tiles256 = static_cast<long>(mb) * ((N + 255) / 256);
  if (tiles256 * 10 >= 7L * sms)
    return launch_gemm<256, 4>(A, lda, W, ldw, C, ldc, M, N, K, epi, -1, stream);
  if (N < 128) return launch_gemm<64, 8>(A, lda, W, ldw, C, ldc, M, N, K, epi, -1, stream);
  return launch_gemm<128, 6>(A, lda, W, ldw, C, ldc, M, N, K, epi, -1, stream);
}

// test hook: force a tile configuration (block_n in {64,128,256}; +1000 forces stream-K, +2000 off;
// 3000 / 3001: swap-AB skinny kernel, single CTA / CTA pair; 4128 / 4256: CTA-pair 256 x BLOCK_N
// tiles; 5128: split-K CTA pairs on 128 x 128 tiles)
int gemm_bf16_cfg(int block_n, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw,
                  __nv_bfloat16* C, int ldc, int M, int N, int K, const GemmEpilogue& epi,
                  cudaStream_t stream) {
  if (check_args(A, lda, W, ldw, C, ldc, M, N, K, epi)) return 1;
  if (block_n == 4256) return launch_gemm<256, 6, kModePair>(A, lda, W, ldw, C, ldc, M, N, K, epi, -1, stream);
  if (block_n == 4128) return launch_gemm<128, 8, kModePair>(A, lda, W, ldw, C, ldc, M, N, K, epi, -1, stream);
  if (block_n == 5128) return launch_gemm<128, 6, kModeSplitK2>(A, lda, W, ldw, C, ldc, M, N, K, epi, -1, stream);
  if (block_n == 3000 || block_n == 3001) {
    const int rc = gemm_skinny_bf16(A, lda, W, ldw, C, ldc, M, N, K, epi, block_n - 3000, stream);
    if (rc < 0) set_last_error("gemm_bf16_cfg: skinny kernel does not handle M=%d", M);
    return rc < 0 ? 1 : rc;
  }
  int fsk = -1;
  if (block_n >= 2000) {
    fsk = 0;
    block_n -= 2000;
  } else if (block_n >= 1000) {
    fsk = 1;
    block_n -= 1000;
  }
  switch (block_n) {
    case 64: return launch_gemm<64, 8>(A, lda, W, ldw, C, ldc, M, N, K, epi, fsk, stream);
    case 128: return launch_gemm<128, 6>(A, lda, W, ldw, C, ldc, M, N, K, epi, fsk, stream);
    case 256: return launch_gemm<256, 4>(A, lda, W, ldw, C, ldc, M, N, K, epi, fsk, stream);
    default: set_last_error("gemm_bf16_cfg: unsupported block_n %d", block_n); return 1;
  }
}

}  // namespace vb
