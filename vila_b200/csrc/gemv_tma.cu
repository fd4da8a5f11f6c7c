// Weight-streaming GEMV with TMA bulk copies into per-warp shared-memory rings (sm_100a).
//
//   y = W x  (M == 1 decode):  every byte of W is read exactly once from HBM, so the kernel is a pure
//   bandwidth problem.  Instead of register-staged LDG (in-flight bytes limited by registers) each of
//   the 8 consumer warps owns a ring of kStages shared-memory slots that its lane 0 keeps filled with
//   cp.async.bulk (1-D TMA) copies of (row, k-part) chunks; completion is signalled on per-slot
//   mbarriers.  ~100-190 KB of weight data are in flight per SM with zero address arithmetic on the
//   load path, and — because the weights are parameters — the first kStages chunks are issued BEFORE
//   the programmatic-dependent-launch wait, overlapping the predecessor kernel's tail and this
//   kernel's own RMSNorm prologue.
//
//   Fusions: RMSNorm(x) prologue, bias, residual, SwiGLU (interleaved gate/up rows), greedy argmax.
//   Partial sums of the k-parts of a row are parked in slots and added in a fixed order.
//
// Replaces cuBLAS GEMV behind nn.Linear + ATen RMSNorm / SiLU / mul / add / argmax at decode time
// (modeling_qwen2.py:81-95,164-176,223-226; HF lm_head + greedy argmax).
#include "common.cuh"
#include "kernels.h"

namespace vb {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxStages = 6;


__device__ __forceinline__ float dot8(const uint4& w, const uint4& x, float acc) {
  acc = fmaf(bf_lo(w.x), bf_lo(x.x), acc);
  acc = fmaf(bf_hi(w.x), bf_hi(x.x), acc);
  acc = fmaf(bf_lo(w.y), bf_lo(x.y), acc);
  acc = fmaf(bf_hi(w.y), bf_hi(x.y), acc);
  acc = fmaf(bf_lo(w.z), bf_lo(x.z), acc);
  acc = fmaf(bf_hi(w.z), bf_hi(x.z), acc);
  acc = fmaf(bf_lo(w.w), bf_lo(x.w), acc);
  acc = fmaf(bf_hi(w.w), bf_hi(x.w), acc);
  return acc;
}

__device__ __forceinline__ uint32_t float_order(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// 1-D bulk copy global -> shared, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

struct TmaGemvLayout {
  int chunk_elems;   // K / ksplit
  int stages;
  int x_off, nw_off, acc_off, ring_off, bar_off, total;
};

__global__ void __launch_bounds__(kThreads, 1)
gemv_tma_kernel(GemvParams p, int rows_per_block, int ksplit, TmaGemvLayout L) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint4* xs = reinterpret_cast<uint4*>(smem + L.x_off);
  float* acc = reinterpret_cast<float*>(smem + L.acc_off);
  uint8_t* ring = smem + L.ring_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  __shared__ float red[32];
  __shared__ unsigned long long best_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * rows_per_block;
  const int nrows = min(rows_per_block, p.N - row0);
  if (nrows <= 0) return;
  const int nvec = p.K >> 3;
  const int chunk_bytes = L.chunk_elems * 2;
  const int chunk_vecs = L.chunk_elems >> 3;
  const int items = nrows * ksplit;
  const int n_my = items > warp ? (items - warp + kWarps - 1) / kWarps : 0;
  uint8_t* my_ring = ring + static_cast<size_t>(warp) * L.stages * chunk_bytes;
  uint64_t* my_bars = bars + warp * kMaxStages;

  uint64_t* x_bar = bars + kWarps * kMaxStages;  // [0]: x arrived, [1]: norm weight arrived
  if (lane == 0) {
    for (int s = 0; s < L.stages; ++s) mbar_init(&my_bars[s], 1);
    if (warp == 0) {
      mbar_init(&x_bar[0], 1);
      mbar_init(&x_bar[1], 1);
    }
    fence_barrier_init();
  }
  __syncwarp();
  griddep_launch_dependents();

  auto issue = [&](int j) {  // lane 0 only: copy item j of this warp into slot j % stages
    const int item = warp + j * kWarps;
    const int r = item / ksplit, part = item - r * ksplit;
    const __nv_bfloat16* src = p.w + static_cast<size_t>(row0 + r) * p.K + part * L.chunk_elems;
    const int s = j % L.stages;
    mbar_arrive_expect_tx(&my_bars[s], chunk_bytes);
    bulk_g2s(my_ring + static_cast<size_t>(s) * chunk_bytes, src, chunk_bytes, &my_bars[s]);
  };

  const bool early = (p.flags & 2) != 0;  // static weights: stream before the dependency wait
  const uint32_t x_bytes = static_cast<uint32_t>(p.K) * 2;
  uint4* nws = reinterpret_cast<uint4*>(smem + L.nw_off);
  int issued = 0;
  if (early && lane == 0) {
    for (; issued < L.stages && issued < n_my; ++issued) issue(issued);
    if (warp == 0 && p.norm_w != nullptr) {  // the norm weight is a parameter as well
      mbar_arrive_expect_tx(&x_bar[1], x_bytes);
      bulk_g2s(nws, p.norm_w, x_bytes, &x_bar[1]);
    }
  }
  griddep_wait();
  // ---- prologue: x arrives as ONE bulk copy (a register-staged loop of dependent LDG -> STS round
  // trips cost ~0.6 us per 4 KB slice: 6 us for the down projection's 38 KB activation vector, during
  // which the full weight rings stalled the HBM stream) ----
  if (lane == 0) {
    if (warp == 0) {
      mbar_arrive_expect_tx(&x_bar[0], x_bytes);
      bulk_g2s(xs, p.x, x_bytes, &x_bar[0]);
      if (!early && p.norm_w != nullptr) {
        mbar_arrive_expect_tx(&x_bar[1], x_bytes);
        bulk_g2s(nws, p.norm_w, x_bytes, &x_bar[1]);
      }
    }
    if (!early)
      for (; issued < L.stages && issued < n_my; ++issued) issue(issued);
  }
  if (threadIdx.x == 0) best_s = 0ull;
  __syncthreads();  // barrier initialisation by warp 0 is visible to every warp
  mbar_wait(&x_bar[0], 0);
  if (p.norm_w != nullptr) {
    float s = 0.f;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const uint4 v = xs[i];
      float t;
      t = bf_lo(v.x); s += t * t;
      t = bf_hi(v.x); s += t * t;
      t = bf_lo(v.y); s += t * t;
      t = bf_hi(v.y); s += t * t;
      t = bf_lo(v.z); s += t * t;
      t = bf_hi(v.z); s += t * t;
      t = bf_lo(v.w); s += t * t;
      t = bf_hi(v.w); s += t * t;
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float t = lane < kWarps ? red[lane] : 0.f;
    t = warp_sum(t);
    const float rstd = rsqrtf(t / p.K + p.norm_eps);
    mbar_wait(&x_bar[1], 0);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {  // each thread rewrites only its own slots
      const uint4 v = xs[i], g = nws[i];
      uint4 o;
      o.x = pack_bf16(bf16_round(bf_lo(v.x) * rstd) * bf_lo(g.x), bf16_round(bf_hi(v.x) * rstd) * bf_hi(g.x));
      o.y = pack_bf16(bf16_round(bf_lo(v.y) * rstd) * bf_lo(g.y), bf16_round(bf_hi(v.y) * rstd) * bf_hi(g.y));
      o.z = pack_bf16(bf16_round(bf_lo(v.z) * rstd) * bf_lo(g.z), bf16_round(bf_hi(v.z) * rstd) * bf_hi(g.z));
      o.w = pack_bf16(bf16_round(bf_lo(v.w) * rstd) * bf_lo(g.w), bf16_round(bf_hi(v.w) * rstd) * bf_hi(g.w));
      xs[i] = o;
    }
  }
  __syncthreads();

  // ---- main loop: consume this warp's ring ----
  for (int j = 0; j < n_my; ++j) {
    const int s = j % L.stages;
    const int item = warp + j * kWarps;
    const int r = item / ksplit, part = item - r * ksplit;
    mbar_wait(&my_bars[s], (j / L.stages) & 1);
    const uint4* wv = reinterpret_cast<const uint4*>(my_ring + static_cast<size_t>(s) * chunk_bytes);
    const uint4* xv = xs + part * chunk_vecs;
    float s0 = 0.f, s1 = 0.f;
    int v = lane;
    for (; v + 32 < chunk_vecs; v += 64) {
      const uint4 a = wv[v], b = wv[v + 32];
      s0 = dot8(a, xv[v], s0);
      s1 = dot8(b, xv[v + 32], s1);
    }
    if (v < chunk_vecs) s0 = dot8(wv[v], xv[v], s0);
    const float tot = warp_sum(s0 + s1);
    __syncwarp();  // every lane is done reading slot s
    if (lane == 0) {
      acc[item] = tot;  // slot = row * ksplit + part
      if (j + L.stages < n_my) issue(j + L.stages);
    }
  }
  __syncthreads();
  if (ksplit > 1) {
    // fixed-order reduction of the k-parts, in place: acc[i] <- sum_q acc[i*ksplit + q].  Rows are
    // processed in phases of blockDim.x; writing row i only overwrites slots of rows <= i, which
    // have been read in this or an earlier phase.
    for (int base = 0; base < nrows; base += blockDim.x) {
      const int i = base + threadIdx.x;
      float tot = 0.f;
      if (i < nrows)
        for (int q = 0; q < ksplit; ++q) tot += acc[i * ksplit + q];
      __syncthreads();
      if (i < nrows) acc[i] = tot;
      __syncthreads();
    }
  }

  // ---- epilogue ----
  if (p.flags & 1) {
    for (int j = threadIdx.x; j < (nrows >> 1); j += blockDim.x) {
      float g = acc[2 * j], u = acc[2 * j + 1];
      if (p.bias) {
        g += __bfloat162float(p.bias[row0 + 2 * j]);
        u += __bfloat162float(p.bias[row0 + 2 * j + 1]);
      }
      g = bf16_round(g);
      u = bf16_round(u);
      p.y[(row0 >> 1) + j] = __float2bfloat16(bf16_round(silu_f(g)) * u);
    }
    return;
  }
  unsigned long long best = 0ull;
  for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
    float v = acc[r];
    if (p.bias) v += __bfloat162float(p.bias[row0 + r]);
    v = bf16_round(v);
    if (p.residual) v = bf16_round(v + __bfloat162float(p.residual[row0 + r]));
    if (p.y) p.y[row0 + r] = __float2bfloat16(v);
    if (p.argmax_key) {
      const unsigned long long key =
          (static_cast<unsigned long long>(float_order(v)) << 32) |
          static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(row0 + r));
      best = key > best ? key : best;
    }
  }
  if (p.argmax_key) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) atomicMax(&best_s, best);
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(p.argmax_key, best_s);
  }
}

}  // namespace

// returns 0 on launch, -1 if the shape does not fit this kernel (caller falls back to the LSU kernel)
int gemv_tma_bf16(const GemvParams& p, cudaStream_t stream) {
  const int sms = num_sms();
  int rows_per_block = (p.N + sms - 1) / sms;
  if ((p.flags & 1) && (rows_per_block & 1)) rows_per_block += 1;
  const int grid = (p.N + rows_per_block - 1) / rows_per_block;
  // Shared-memory budget: (almost) the whole SM.  Measured on B200: under a saturated HBM pipe the
  // load latency is ~2.5 us, so ~110+ KB must be in flight per SM to sustain the full rate; halving
  // the rings to let the next kernel's CTA co-reside (PDL) dropped the gate/up GEMV from 97 % to
  // 76 % of the measured HBM peak (profiles/r01_gemv_variants.md).
  constexpr int kSmemBudget = 220 * 1024;
  // x and the norm weight arrive by bulk copy: 16-byte aligned sources (else: register-staged kernel)
  if ((reinterpret_cast<uintptr_t>(p.x) & 15) || (p.norm_w && (reinterpret_cast<uintptr_t>(p.norm_w) & 15)))
    return -1;
  const int x1_bytes = (p.K * 2 + 127) / 128 * 128;
  const int x_bytes = x1_bytes * (p.norm_w ? 2 : 1);  // x (+ norm weight staging)
  TmaGemvLayout L;
  int ksplit = -1;
  for (int ks = 1; ks <= 128; ++ks) {
    if (p.K % ks) continue;
    const int ce = p.K / ks;
    if (ce % 8) continue;
    const int cb = ce * 2;
    if (cb > 8192) continue;
    if (cb < 512) break;
    const int acc_bytes = (rows_per_block * ks * 4 + 127) / 128 * 128;
    const int ring_budget = kSmemBudget - x_bytes - acc_bytes - (kWarps * kMaxStages + 2) * 8 - 256;
    int st = ring_budget / (kWarps * cb);
    if (st > kMaxStages) st = kMaxStages;
    if (st < 3) continue;  // need a few chunks in flight per warp
    ksplit = ks;
    L.stages = st;
    if (static_cast<long>(rows_per_block) * ks >= 6L * kWarps) break;  // enough items per warp
  }
  if (ksplit < 0) return -1;
  L.chunk_elems = p.K / ksplit;
  const int chunk_bytes = L.chunk_elems * 2;
  L.x_off = 0;
  L.nw_off = x1_bytes;
  L.acc_off = x_bytes;
  L.ring_off = (L.acc_off + rows_per_block * ksplit * 4 + 127) / 128 * 128;
  {
    // recompute stages for the final ksplit (the loop may have broken on an earlier candidate)
    const int ring_budget = kSmemBudget - L.ring_off - (kWarps * kMaxStages + 2) * 8 - 256;
    int st = ring_budget / (kWarps * chunk_bytes);
    if (st > kMaxStages) st = kMaxStages;
    if (st < 2) return -1;
    L.stages = st;
  }
  L.bar_off = (L.ring_off + kWarps * L.stages * chunk_bytes + 127) / 128 * 128;
  L.total = L.bar_off + (kWarps * kMaxStages + 2) * 8;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    VB_CUDA(cudaFuncSetAttribute(gemv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 224 * 1024));
  }
  VB_CUDA(launch_pdl(gemv_tma_kernel, dim3(grid), dim3(kThreads), static_cast<size_t>(L.total), stream,
                     p, rows_per_block, ksplit, L));
  return 0;
}

}  // namespace vb
