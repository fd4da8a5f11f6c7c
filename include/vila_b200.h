/*
 * vila_b200 — C-ABI of the B200-native (sm_100a) VILA multimodal forward hot path.
 *
 * The reference (NVlabs/VILA) has no FFI on this path: every GPU op is a library call made from
 * Python (flash_attn, cuBLAS through nn.Linear, cuDNN through nn.Conv2d, ATen elementwise).  This
 * header is the boundary a maintainer would bind instead of those calls; each entry point names the
 * reference call site (paths relative to the VILA repo root) it replaces.  The only native-extension
 * precedent in the tree (llava/model/coat/optimizer/kernels/bindings.cpp:6-10) takes tensors,
 * mutates in place, returns void and runs on the current stream; we keep "caller owns all memory,
 * in-place/out-param, stream-ordered", but with plain pointers so any host language can bind it.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers to bf16 (uint16 storage) unless stated otherwise,
 *     row-major, 16-byte aligned; leading dimensions are in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success; non-zero on error, message via vila_last_error()
 *     (thread-local). Nothing falls back to the CPU: without an sm_100a device calls fail.
 *   - no hidden state beyond (a) the per-device stream-K scratch registered with vila_set_workspace and
 *     (b) per-device "function attributes set" flags: KV pool, page tables, workspaces and counters
 *     are caller-allocated.
 */
#ifndef VILA_B200_H_
#define VILA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VILA_ACT_NONE 0
#define VILA_ACT_GELU_TANH 1 /* SigLIP "gelu_pytorch_tanh"  (modeling_siglip.py:707-715) */
#define VILA_ACT_GELU_ERF 2  /* nn.GELU() in mm_projector   (base_projector.py:145-162)  */
#define VILA_ACT_SILU 3

/* `flags` of vila_linear / vila_gemv:
 *   VILA_FLAG_SWIGLU    weight rows are interleaved (gate_0, up_0, gate_1, ...); out is [M, N/2] =
 *                       silu(gate) * up
 *   VILA_FLAG_STATIC_W  the weight matrix is a parameter that no earlier kernel in this stream is
 *                       still writing: the kernel may fetch it BEFORE the programmatic-dependent-launch
 *                       wait (all vila_* kernels are launched with PDL so that prologues and weight
 *                       prefetch overlap the predecessor's tail). Leave it clear for weights produced
 *                       on the fly. */
#define VILA_FLAG_SWIGLU 1
#define VILA_FLAG_STATIC_W 2

const char* vila_last_error(void);
int vila_abi_version(void); /* 2 */
/* Register caller-owned, ZERO-INITIALISED device scratch (256-byte aligned) used by vila_linear's
 * stream-K schedule for fp32 partial sums (64 KiB of counters + M*N*4 bytes per call that uses it;
 * calls that do not fit simply use the data-parallel schedule).  The kernels leave it zeroed again.
 * The registration is per DEVICE (the current device at the time of the call); one workspace serves
 * one stream at a time.  ptr == NULL unregisters. */
int vila_set_workspace(void* ptr, uint64_t bytes);
/* fills SM count and compute capability of the current device */
int vila_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * vila_linear — out[M,N] = epilogue(x[M,K] · w[N,K]^T)        (tcgen05 / TMEM / TMA GEMM)
 *   epilogue: (+bias[N]) -> act -> (+residual[row % res_row_mod or row, :])   or SwiGLU:
 *   flags & VILA_FLAG_SWIGLU: w rows are interleaved (gate_0, up_0, gate_1, up_1, ...) and out is
 *   [M, N/2] = silu(gate) * up.
 * Replaces nn.Linear (+ the ATen GELU / SiLU*mul / residual add that follows it):
 *   SigLIP q/k/v/out_proj, fc1/fc2 : llava/model/multimodal_encoder/siglip/modeling_siglip.py:384-387,707-715,752,757
 *   patch-embed Conv2d (as im2col GEMM, residual = position embedding with res_row_mod = #patches): :269-275,322-328
 *   mm_projector Linear layers     : llava/model/multimodal_projector/base_projector.py:145-162
 *   Qwen2 q/k/v/o, gate/up/down, lm_head: transformers Qwen2 (in-tree copy
 *     llava/eval/vision_niah_vila/zigzag_ring_attn/modeling_qwen2.py:164-176,223-226,633-706)
 * ------------------------------------------------------------------------------------------- */
int vila_linear(const void* x, int64_t ldx, const void* w, int64_t ldw, const void* bias,
                const void* residual, int64_t ld_res, int res_row_mod, void* out, int64_t ldo,
                int M, int N, int K, int act, int flags, void* stream);
/* test hook: same, forcing the kernel configuration instead of the size heuristic:
 *   64 / 128 / 256   single-CTA tiles 128 x block_n (+1000: deterministic stream-K, +2000: stream-K off)
 *   4128 / 4256      CTA-pair tiles 256 x {128,256} (tcgen05 cta_group::2)
 *   5128             split-K CTA pairs on 128 x 128 tiles (needs <= #SM/2 tiles, K > 64)
 *   3000 / 3001      swap-AB skinny kernel (M <= 512), one CTA / CTA pair per weight block */
int vila_linear_cfg(int block_n, const void* x, int64_t ldx, const void* w, int64_t ldw,
                    const void* bias, const void* residual, int64_t ld_res, int res_row_mod,
                    void* out, int64_t ldo, int M, int N, int K, int act, int flags, void* stream);

/* nn.LayerNorm over the last dim (modeling_siglip.py:723,725,746,755; base_projector.py:147) */
int vila_layernorm(const void* x, const void* weight, const void* bias, void* out, int rows,
                   int cols, float eps, void* stream);
/* Qwen2RMSNorm (modeling_qwen2.py:81-95). If residual_add != NULL: x += residual_add first (in place). */
int vila_rmsnorm(void* x_inout, const void* residual_add, const void* weight, void* out, int rows,
                 int cols, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * vila_fmha — flash attention forward, tcgen05 QK^T / PV with TMEM accumulators.
 *   q/o element (token t, head h, dim d) at  base + t*tok_stride + h*head_stride + d
 *   k/v: if kv_page_stride != 0 the KV cache is paged (128 tokens per page):
 *          element (page p, row r, head h, d) at base + p*kv_page_stride + r*kv_tok_stride + h*kv_head_stride + d
 *          and page_table[b*page_table_stride + j] gives the page of block j (NULL: identity);
 *        else k/v are [B*Sk, Hkv, D] views like q.
 * Replaces flash_attn_func(q,k,v,causal=False) in SiglipFlashAttention2 (modeling_siglip.py:583-585)
 * and HF _flash_attention_forward for Qwen2/Llama (patched at llava/train/sequence_parallel/
 * monkey_patch.py:133-239, llava/model/utils/packing.py:36), causal GQA.
 * ------------------------------------------------------------------------------------------- */
typedef struct vila_fmha_params {
  const void* q;
  int64_t q_tok_stride, q_head_stride;
  const void* k;
  const void* v;
  int64_t kv_page_stride, kv_tok_stride, kv_head_stride, kv_num_pages;
  const int32_t* page_table;
  int32_t page_table_stride;
  void* o;
  int64_t o_tok_stride, o_head_stride;
  int32_t B, Sq, Sk, Hq, Hkv, D, causal;
  float scale;
} vila_fmha_params;
int vila_fmha(const vila_fmha_params* p, void* stream);
/* test hook (like vila_linear_cfg): same, forcing the kernel flavour instead of the size heuristic:
 *   0 heuristic; 1 one query tile per CTA (fmha_fwd_kernel); 2 two query tiles per CTA with ping-pong
 *   softmax warpgroups and the O accumulator in TMEM (fmha2_fwd_kernel; needs Sq > 128, D in 72..96 or 128);
 *   3 / 4: as 2 with every 4th / every 2nd exponential as an FMA-pipe polynomial (measured slower) */
int vila_fmha_cfg(int variant, const vila_fmha_params* p, void* stream);

/* im2col for Conv2d(3,1152,k=14,s=14) (modeling_siglip.py:269-275): pixels [B,C,H,W] ->
 * out [B*(H/P)*(W/P), k_pad], column = (c, ky, kx); columns >= C*P*P are zero. */
int vila_patch_im2col(const void* pixels, void* out, int B, int C, int H, int W, int patch,
                      int k_pad, void* stream);
/* Image preprocessing on the device.  Replaces, per resize grid of
 * mm_utils.process_image (llava/mm_utils.py:442-522; dynamic_preprocess :299-338, dynamic_s2_preprocess
 * :341-405), PIL `image.resize((out_w, out_h))` (bicubic; Pillow Resample.c 8bpc fixed point) + the crop
 * into tile x tile blocks + SiglipImageProcessor.preprocess (rescale 1/255, normalise (x-mean)/std):
 *   src      uint8 [H, W, 3] (device)      tmp  uint8 scratch [H, out_w, 3]
 *   coef_*   int32 [out, ksize] 22-bit filter taps, bounds_* int32 [out, 2] = (first input index, count)
 *            (host: vila_b200.model.media.bicubic_coeffs == Pillow precompute_coeffs/normalize_coeffs_8bpc)
 *   out      bf16 [n_tiles, 3, tile, tile]; pixel (Y, X) of the resized image lands in tile
 *            tile_index0 + (Y / tile) * (out_w / tile) + X / tile.  Bit-identical to the PIL path. */
int vila_resize_bicubic_tiles(const uint8_t* src, int H, int W, int out_w, int out_h,
                              const int32_t* coef_x, const int32_t* bounds_x, int ksize_x,
                              const int32_t* coef_y, const int32_t* bounds_y, int ksize_y, uint8_t* tmp,
                              void* out_tiles, int tile, int tile_index0, float mean, float stdv,
                              void* stream);
/* DownSampleBlock.flat_square / flat_square_2x2 / flat_square_3x3 (base_projector.py:58-123):
 * x [B, h*w, C] -> out [B, ceil(h/r)*ceil(w/r), r*r*C], zero padded. */
int vila_space_to_depth(const void* x, void* out, int B, int h, int w, int C, int r, void* stream);
/* merge_features_for_dynamic_s2 + split_chessboard for one image (llava_arch.py:282-364). */
int vila_s2_merge(const void* tiles, void* out, int side, int C, int n_scales,
                  const int* splits_h, const int* splits_w, int out_bh, int out_bw, int share_tile,
                  void* stream);
/* merge_chessboard + "(h w) c" flatten of projected tiles (llava_arch.py:255-280,384-390). */
int vila_chessboard_merge(const void* tiles, void* out, int bh, int bw, int s, int C, void* stream);
/* TSPVideoEncoder pooling (llava/model/encoders/video/tsp.py:11-12,28-51). */
int vila_tsp_pool(const void* x, void* out, int T, int h, int w, int C, int pt, int ph, int pw,
                  void* stream);
/* text/media embedding splice of LlavaMetaForCausalLM._embed (llava_arch.py:429,457-479):
 * out[i] = src[i] >= 0 ? table[src[i]] : media[-(src[i]+1)].  src is int32 on the device. */
int vila_embed_splice(const void* table, const void* media, const int32_t* src, void* out, int rows,
                      int cols, void* stream);
/* HF apply_rotary_pos_emb (modeling_qwen2.py:99-160) in place on q,k of qkv [S,(Hq+2Hkv)*D] and
 * DynamicCache.update as a scatter into the paged pool (k_pool may be NULL: RoPE only).
 * inv_freq: fp32 [D/2] on the device. positions: int32 [S] on the device. */
int vila_rope_kv_append(void* qkv, const int32_t* positions, int S, int Hq, int Hkv, int D,
                        const float* inv_freq, void* k_pool, void* v_pool,
                        const int32_t* page_table, int cache_pos0, void* stream);
/* (cache_pos0 < 0: the cache slot of row s is positions[s] itself — decode, position on the device) */
/* vila_rope_kv_append with cos / sin taken from vila_rope_table(positions) (computed once per request
 * and shared by all layers and heads; 16-byte accesses): the long-prefill form, bit-identical. */
int vila_rope_kv_append_table(void* qkv, const void* rope_table, int S, int Hq, int Hkv, int D,
                              void* k_pool, void* v_pool, const int32_t* page_table, int cache_pos0,
                              void* stream);
/* Fused q/k/v projection for a short prefill chunk (M <= 384 tokens, head_dim 128):
 * qkv = x @ w^T + bias; RoPE on the q and k heads; q heads -> qkv_out[:, :Hq*128]; k / v heads ->
 * the paged pools (k_pool NULL: they stay in qkv_out) — vila_linear followed by vila_rope_kv_append
 * in ONE kernel, bit-identical to the two calls (rope_table: vila_rope_table of the chunk's positions).  Replaces Qwen2Attention's q_proj/k_proj/v_proj +
 * apply_rotary_pos_emb + past_key_value.update (modeling_qwen2.py:223-226,99-160,262-266 of the
 * in-tree copy).  Returns 3 (and sets vila_last_error) when the shape is not covered: the caller
 * then issues the two separate calls. */
int vila_linear_qkv_rope(const void* x, int64_t ldx, const void* w, int64_t ldw, const void* bias,
                         void* qkv_out, int64_t ldo, int M, int K, int Hq, int Hkv, int D,
                         const void* rope_table, void* k_pool, void* v_pool,
                         const int32_t* page_table, int cache_pos0, int flags, void* stream);
/* cos / sin table of Qwen2RotaryEmbedding.forward (modeling_qwen2.py:99-143) for one request,
 * computed once and shared by all layers: table [S, D] bf16 = cos(pos*inv_freq[0..D/2)) | sin(...). */
int vila_rope_table(const int32_t* positions, int S, int D, const float* inv_freq, void* table,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * decode (one token): weight-streaming GEMV with fused RMSNorm prologue and bias / residual /
 * SwiGLU / greedy-argmax epilogues; split-KV paged GQA attention with fused RoPE + KV append.
 * ------------------------------------------------------------------------------------------- */
typedef struct vila_gemv_params {
  const void* x;
  const void* w;
  const void* bias;
  const void* norm_w;
  float norm_eps;
  const void* residual;
  void* y;
  int32_t N, K;
  int32_t flags; /* VILA_FLAG_SWIGLU: rows interleaved (gate, up), y is [N/2];
                    VILA_FLAG_STATIC_W: see below */
  unsigned long long* argmax_key; /* device u64, zero before the first launch */
} vila_gemv_params;
int vila_gemv(const vila_gemv_params* p, void* stream);

int vila_argmax_finalize(unsigned long long* key, int32_t* token_out, int32_t* token_hist,
                         int32_t* step_counter, int32_t* position, const void* embed_table,
                         void* x_next, int hidden, void* stream);

typedef struct vila_decode_attn_params {
  void* qkv;
  const int32_t* position;
  void* k_pool;
  void* v_pool;
  const int32_t* page_table;
  void* out;
  float* ws;         /* >= Hkv*num_splits*G*(D+2) floats */
  int32_t* counters; /* Hkv ints, zero-initialised once */
  const float* inv_freq;
  int32_t Hq, Hkv, D, num_splits; /* num_splits 0: one CTA per query head, no split: the cache holds at
                                     most 1024 tokens (page_table has >= 8 entries); 1..8: KV splits as a
                                     thread-block cluster; 9..64: splits combined through ws/counters */
  float scale;
} vila_decode_attn_params;
int vila_decode_attention(const vila_decode_attn_params* p, void* stream);

/* Batched decode attention for continuous batching over ONE shared paged pool: `batch` sequences, sequence
 * b uses qkv + b*qkv_stride, out + b*out_stride, position[b] (< 0: idle slot, skipped) and the page-table
 * row page_table + b*pt_stride (max_pages <= 32 valid entries: contexts up to 4096 tokens).  ws / counters /
 * num_splits of the struct are ignored.  One CTA per (query head, sequence); RoPE + KV append fused. */
int vila_decode_attention_batch(const vila_decode_attn_params* p, int batch, int qkv_stride,
                                int out_stride, int pt_stride, int max_pages, void* stream);

/* Long-context decode attention (video: 16K-66K cached tokens = 34-135 MB of K/V per layer): RoPE +
 * KV append for the new token, then the tcgen05 FMHA kernel in split-KV mode (the G query heads of a
 * KV group are the query rows of its 128-row tile; K/V pages stream through TMA on
 * Hkv * num_splits CTAs), then a deterministic combine.  split j covers tokens
 * [j*split_tokens, (j+1)*split_tokens) (multiple of 128; num_splits*split_tokens >= max context).
 * Replaces the same HF calls as vila_decode_attention (apply_rotary_pos_emb + DynamicCache.update +
 * flash-attn decode, modeling_qwen2.py:99-160,262-310). */
typedef struct vila_decode_attn_split_params {
  void* qkv;
  const int32_t* position;
  void* k_pool;
  void* v_pool;
  const int32_t* page_table;
  int64_t kv_num_pages;
  void* out;
  float* o_partial; /* >= num_splits*Hq*D floats */
  float* lse;       /* >= num_splits*Hq floats */
  int32_t* counters; /* Hkv ints, zero before the first launch (self-cleaning): the last split CTA of a KV head
                        merges the partials itself; NULL: a separate combine kernel is launched */
  const float* inv_freq;
  int32_t Hq, Hkv, D, num_splits, split_tokens;
  float scale;
} vila_decode_attn_split_params;
int vila_decode_attention_split(const vila_decode_attn_split_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * vila_decode_mega — n_tokens greedy decode steps of the whole LLM in ONE persistent launch
 * (one CTA per SM, weights streamed through per-warp TMA rings that run ahead across layer and token
 * boundaries, grid barriers between phases).  Replaces the per-token HF generate loop
 * (llava_arch.py:833 -> GenerationMixin: ~400 launches and one D2H sync per token).
 * `layers` is a DEVICE array of num_layers entries.  `barrier` and `epoch` are device u32 words that
 * must both be zero before the first launch (the kernel keeps them consistent afterwards).
 * Head dim is 128.  The new tokens are appended to hist[*step ...], *position and *step advance.
 * ------------------------------------------------------------------------------------------- */
typedef struct vila_mega_layer {
  const void* qkv_w;
  const void* qkv_b;
  const void* o_w;
  const void* gu_w;
  const void* down_w;
  const void* ln1_w;
  const void* ln2_w;
  void* k_pool;
  void* v_pool;
} vila_mega_layer;
typedef struct vila_mega_params {
  const vila_mega_layer* layers;
  int32_t num_layers;
  const void* final_norm_w;
  const void* lm_head_w;
  const void* embed;
  int32_t hidden, inter, Hq, Hkv, vocab;
  float eps, scale;
  const float* inv_freq;
  const int32_t* page_table;
  void* x;
  void* qkv;
  void* act;
  float* attn_ws;
  int32_t* attn_counters; /* Hkv ints, zero before the first launch */
  unsigned long long* key;
  int32_t* token;
  int32_t* hist;
  int32_t* step;
  int32_t* position;
  uint32_t* barrier;
  uint32_t* epoch;
  int32_t n_tokens, splits;
} vila_mega_params;
int vila_decode_mega(const vila_mega_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VILA_B200_H_ */
