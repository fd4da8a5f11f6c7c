"""ORACLE — test infrastructure only.  NOT part of the product path.

CPU (or any-device) restatement in plain PyTorch of the reference algorithm for the VILA multimodal
forward hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; vila_b200/ never does.

Every function cites the reference file:line it follows (paths relative to the NVlabs/VILA root,
commit b760c34b).  Pinning status: the reference ships NO tests / golden vectors for this path
(SURVEY.md §4), so the oracle is pinned instead against the reference's own modules executed in the
authoring container (oracle/validate_against_reference.py: vendored SigLIP + MultimodalProjector
loaded by file path, llava_arch glue compared function by function, and the installed
`transformers` Qwen2ForCausalLM for the third-party LLM arithmetic) and the resulting tensors are
committed under tests/golden/ (oracle/gen_golden.py).

All functions are dtype-agnostic: run them in fp32 for the "truth" and in bf16 to reproduce the
reference's rounding points (the reference runs unfused bf16/fp16 torch ops).
"""
from __future__ import annotations

import math
from collections import deque
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100  # llava/constants.py


# =================================================================================================
# configs (architecture constants from scripts/NVILA/*.sh)
# =================================================================================================
@dataclass
class SiglipCfg:
    hidden_size: int = 1152
    intermediate_size: int = 4304
    num_hidden_layers: int = 27
    num_attention_heads: int = 16
    image_size: int = 448
    patch_size: int = 14
    num_channels: int = 3
    layer_norm_eps: float = 1e-6

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size


@dataclass
class Qwen2Cfg:
    hidden_size: int = 3584
    intermediate_size: int = 18944
    num_hidden_layers: int = 28
    num_attention_heads: int = 28
    num_key_value_heads: int = 4
    vocab_size: int = 152064
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    head_dim: int = 128


# =================================================================================================
# SigLIP vision tower  (llava/model/multimodal_encoder/siglip/modeling_siglip.py)
# =================================================================================================
def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """ACT2FN["gelu_pytorch_tanh"] used by SiglipMLP (modeling_siglip.py:707-715)."""
    return F.gelu(x, approximate="tanh")


def siglip_embeddings(pixels, w_patch, b_patch, pos_emb, patch: int):
    """SiglipVisionEmbeddings.forward (modeling_siglip.py:320-329): conv(k=s=patch) -> flatten ->
    transpose -> + position_embedding(position_ids)."""
    x = F.conv2d(pixels, w_patch, b_patch, stride=patch)  # :322
    x = x.flatten(2).transpose(1, 2)  # :323
    return x + pos_emb.unsqueeze(0)  # :328


def siglip_attention(x, p: Dict[str, torch.Tensor], prefix: str, num_heads: int):
    """SiglipAttention.forward (modeling_siglip.py:389-439); the flash / sdpa variants
    (:461-587,:639-692) compute the same function: softmax in fp32, scale = head_dim**-0.5."""
    B, N, Cc = x.shape
    hd = Cc // num_heads
    q = F.linear(x, p[prefix + "q_proj.weight"], p[prefix + "q_proj.bias"])
    k = F.linear(x, p[prefix + "k_proj.weight"], p[prefix + "k_proj.bias"])
    v = F.linear(x, p[prefix + "v_proj.weight"], p[prefix + "v_proj.bias"])
    q = q.view(B, N, num_heads, hd).transpose(1, 2)
    k = k.view(B, N, num_heads, hd).transpose(1, 2)
    v = v.view(B, N, num_heads, hd).transpose(1, 2)
    attn = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)  # :407
    attn = F.softmax(attn, dim=-1, dtype=torch.float32).to(q.dtype)  # :423
    o = torch.matmul(attn, v).transpose(1, 2).reshape(B, N, Cc)  # :425-434
    return F.linear(o, p[prefix + "out_proj.weight"], p[prefix + "out_proj.bias"])  # :436


def siglip_layer(x, p, prefix: str, cfg: SiglipCfg):
    """SiglipEncoderLayer.forward (modeling_siglip.py:728-764)."""
    r = x
    h = F.layer_norm(x, (cfg.hidden_size,), p[prefix + "layer_norm1.weight"],
                     p[prefix + "layer_norm1.bias"], cfg.layer_norm_eps)
    h = siglip_attention(h, p, prefix + "self_attn.", cfg.num_attention_heads)
    x = r + h
    r = x
    h = F.layer_norm(x, (cfg.hidden_size,), p[prefix + "layer_norm2.weight"],
                     p[prefix + "layer_norm2.bias"], cfg.layer_norm_eps)
    h = F.linear(h, p[prefix + "mlp.fc1.weight"], p[prefix + "mlp.fc1.bias"])
    h = gelu_tanh(h)
    h = F.linear(h, p[prefix + "mlp.fc2.weight"], p[prefix + "mlp.fc2.bias"])
    return r + h


def siglip_tower(pixels, p: Dict[str, torch.Tensor], cfg: SiglipCfg, select_layer: int = -2):
    """SiglipVisionTransformer.forward with output_hidden_states=True (modeling_siglip.py:1172-1211,
    encoder loop :994-1017) followed by VisionTower.feature_select with `cls_patch`
    (vision_encoder.py:44-52): returns hidden_states[select_layer].  hidden_states has
    num_layers+1 entries (embeddings first); [-2] is the INPUT of the last layer, so only
    num_layers-1 layers are evaluated here (the reference computes and discards the last one)."""
    pre = "vision_model."
    x = siglip_embeddings(pixels, p[pre + "embeddings.patch_embedding.weight"],
                          p[pre + "embeddings.patch_embedding.bias"],
                          p[pre + "embeddings.position_embedding.weight"], cfg.patch_size)
    n_states = cfg.num_hidden_layers + 1
    idx = select_layer if select_layer >= 0 else n_states + select_layer
    for i in range(idx):
        x = siglip_layer(x, p, f"{pre}encoder.layers.{i}.", cfg)
    return x


# =================================================================================================
# mm_projector (llava/model/multimodal_projector/base_projector.py)
# =================================================================================================
def flat_square(x: torch.Tensor, r: int) -> torch.Tensor:
    """DownSampleBlock.flat_square / flat_square_2x2 (:58-97) for r=2, flat_square_3x3 (:110-123)
    for r=3.  x: [n, w, h, c] -> [n, ceil(w/r), ceil(h/r), r*r*c] with zero padding."""
    n, w, h, c = x.shape
    if w % r != 0:
        x = torch.cat([x, x.new_zeros((n, r - w % r, h, c))], dim=1)
        n, w, h, c = x.shape
    if h % r != 0:
        x = torch.cat([x, x.new_zeros((n, w, r - h % r, c))], dim=2)
        n, w, h, c = x.shape
    x = x.contiguous().view(n, w, h // r, c * r)
    x = x.permute(0, 2, 1, 3).contiguous()
    x = x.view(n, h // r, w // r, c * r * r)
    x = x.permute(0, 2, 1, 3).contiguous()
    return x


def downsample(x: torch.Tensor, r: int) -> torch.Tensor:
    """DownSampleBlock / DownSample2x2BlockFix / DownSample3x3BlockFix .forward (:49-56,74-81,100-107)."""
    h = w = int(x.shape[1] ** 0.5)
    x = x.reshape(x.shape[0], h, w, -1)
    x = flat_square(x, r)
    return x.reshape(x.shape[0], -1, x.shape[-1])


def projector(x, p: Dict[str, torch.Tensor], kind: str):
    """MultimodalProjector.forward (:248-252) for the Sequential stacks defined at :145-176.
    nn.GELU() is the exact (erf) GELU; LayerNorm eps is the default 1e-5."""
    if kind in ("mlp_downsample", "mlp_downsample_2x2_fix"):
        x = downsample(x, 2)
        x = F.layer_norm(x, (x.shape[-1],), p["layers.1.weight"], p["layers.1.bias"], 1e-5)
        x = F.linear(x, p["layers.2.weight"], p["layers.2.bias"])
        x = F.gelu(x)
        return F.linear(x, p["layers.4.weight"], p["layers.4.bias"])
    if kind == "mlp_downsample_3x3_fix":
        x = downsample(x, 3)
        x = F.layer_norm(x, (x.shape[-1],), p["layers.1.weight"], p["layers.1.bias"], 1e-5)
        x = F.linear(x, p["layers.2.weight"], p["layers.2.bias"])
        x = F.gelu(x)
        x = F.layer_norm(x, (x.shape[-1],), p["layers.4.weight"], p["layers.4.bias"], 1e-5)
        x = F.linear(x, p["layers.5.weight"], p["layers.5.bias"])
        x = F.gelu(x)
        return F.linear(x, p["layers.7.weight"], p["layers.7.bias"])
    raise ValueError(f"Unknown projector type: {kind}")


# =================================================================================================
# llava_arch.py glue: chessboard / dynamic-S2 / encode_images
# =================================================================================================
def merge_chessboard(x: torch.Tensor, num_split_h: int, num_split_w: int) -> torch.Tensor:
    """LlavaMetaModel.merge_chessboard (llava_arch.py:255-280). x: b*n*c or b*c*h*w -> b*c*H*W"""
    B = x.shape[0]
    if x.dim() == 3:
        N = x.shape[1]
        s = int(N ** 0.5)
        x = x.reshape(B, s, s, x.shape[2]).permute(0, 3, 1, 2)  # "b (h w) c -> b c h w"
    assert B % (num_split_h * num_split_w) == 0
    b = B // (num_split_h * num_split_w)
    rows = []
    for i in range(num_split_h):
        rows.append(torch.cat([x[(i * num_split_w + j) * b:(i * num_split_w + j + 1) * b]
                               for j in range(num_split_w)], dim=-1))
    return torch.cat(rows, dim=-2)


def split_chessboard(x: torch.Tensor, num_split_h: int, num_split_w: int) -> torch.Tensor:
    """LlavaMetaModel.split_chessboard (llava_arch.py:282-296)."""
    B, Cc, H, W = x.shape
    assert H % num_split_h == 0 and W % num_split_w == 0
    h, w = H // num_split_h, W // num_split_w
    return torch.cat([x[:, :, i * h:(i + 1) * h, j * w:(j + 1) * w]
                      for i in range(num_split_h) for j in range(num_split_w)], dim=0)


def merge_features_for_dynamic_s2(image_features, block_sizes, scales: Sequence[int],
                                  resize_output_to_scale_idx: int):
    """LlavaMetaModel.merge_features_for_dynamic_s2 (llava_arch.py:298-364)."""
    out, new_block_sizes = [], []
    cnt = 0
    for bs in block_sizes:
        if bs is None:
            cur = image_features[cnt:cnt + 1]
            s = int(cur.shape[1] ** 0.5)
            cur = cur.reshape(1, s, s, -1).permute(0, 3, 1, 2)
            cur = cur.repeat(1, len(scales), 1, 1)
            out.append(cur)
            new_block_sizes.append((1, 1))
            cnt += 1
            continue
        per_scale = []
        for scale in scales[:-1]:
            nb = (scale // scales[0]) ** 2
            per_scale.append(merge_chessboard(image_features[cnt:cnt + nb], scale // scales[0],
                                              scale // scales[0]))
            cnt += nb
        nb = bs[0] * bs[1]
        per_scale.append(merge_chessboard(image_features[cnt:cnt + nb], bs[0], bs[1]))
        cnt += nb
        output_size = per_scale[resize_output_to_scale_idx].shape[-2:]
        cur = torch.cat([F.interpolate(f.to(torch.float32), size=output_size, mode="area").to(f.dtype)
                         for f in per_scale], dim=1)  # :336-345
        out.append(cur)
        if resize_output_to_scale_idx == len(scales) - 1 or resize_output_to_scale_idx == -1:
            new_block_sizes.append(bs)
        else:
            k = scales[resize_output_to_scale_idx] // scales[0]
            new_block_sizes.append((k, k))
    assert cnt == len(image_features)
    return out, new_block_sizes


def encode_images(images, tower_fn, projector_fn, *, dynamic_s2: bool = False, block_sizes=None,
                  scales: Sequence[int] = (448, 896, 1344), resize_output_to_scale_idx: int = -1):
    """LlavaMetaModel.encode_images (llava_arch.py:366-394)."""
    if block_sizes is None:
        block_sizes = [None] * len(images)
    if not dynamic_s2:
        return projector_fn(tower_fn(images))
    feats = tower_fn(images)
    feats, nbs = merge_features_for_dynamic_s2(feats, block_sizes, scales, resize_output_to_scale_idx)
    feats = [split_chessboard(x, b[0], b[1]) for x, b in zip(feats, nbs)]
    feats = torch.cat([x.flatten(2).transpose(1, 2) for x in feats], dim=0)  # "b c h w -> b (h w) c"
    feats = projector_fn(feats)
    feats = list(feats.split([b[0] * b[1] for b in nbs], dim=0))
    feats = [merge_chessboard(x, b[0], b[1]) for x, b in zip(feats, nbs)]
    feats = [x[0].flatten(1).transpose(0, 1) for x in feats]  # "1 c h w -> (h w) c"
    if all(f.shape[0] == feats[0].shape[0] for f in feats):
        feats = torch.stack(feats, dim=0)
    return feats


# =================================================================================================
# media encoders (llava/model/encoders)
# =================================================================================================
def image_encoder(features, end_embeds: Optional[torch.Tensor], start_embeds=None):
    """BasicImageEncoder._process_features per image (encoders/image/basic.py:29-39,73-79)."""
    outs = []
    for f in features:
        if start_embeds is not None:
            f = torch.cat([start_embeds, f], dim=0)
        if end_embeds is not None:
            f = torch.cat([f, end_embeds], dim=0)
        outs.append(f)
    return outs


def video_encoder(features, end_embeds, start_embeds=None):
    """BasicVideoEncoder._process_features (encoders/video/basic.py:29-41): per-frame start/end
    tokens, then flatten(0,1). features: [T, N, C]."""
    if start_embeds is not None:
        features = torch.cat([torch.stack([start_embeds] * features.shape[0], 0), features], dim=1)
    if end_embeds is not None:
        features = torch.cat([features, torch.stack([end_embeds] * features.shape[0], 0)], dim=1)
    return features.flatten(0, 1)


def tsp_pool(x: torch.Tensor, size: int, dim: int) -> torch.Tensor:
    """encoders/video/tsp.py:11-12."""
    return x.view(x.shape[:dim] + (-1, size) + x.shape[dim + 1:]).mean(dim + 1)


def tsp_video_encoder(inputs, pool_sizes, end_embeds, start_embeds=None, sep_embeds=None):
    """TSPVideoEncoder._process_features (encoders/video/tsp.py:28-51). inputs: [T, N, C]."""
    nt, ns = inputs.shape[:2]
    nl = int(ns ** 0.5)
    outs = []
    for pool_size in pool_sizes:
        f = inputs.view(nt, nl, nl, -1)
        for dim, pp in enumerate(pool_size):
            f = tsp_pool(f, pp, dim=dim)
        f = f.flatten(1, 2)
        f = video_encoder(f, end_embeds, start_embeds)
        if sep_embeds is not None:
            f = torch.cat([f, sep_embeds], dim=0)
        outs.append(f)
    return torch.cat(outs, dim=0)


# =================================================================================================
# _embed: text/media splice (llava_arch.py:412-490, 528-555)
# =================================================================================================
def embed_splice(input_ids: torch.Tensor, embed_table: torch.Tensor,
                 media_embeds: Dict[str, List[torch.Tensor]], media_token_ids: Dict[str, int],
                 labels: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                 padding_side: str = "right"):
    """LlavaMetaForCausalLM._embed after the encoders ran (llava_arch.py:419-490) +
    __batchify_sequence (:528-555). Returns (inputs_embeds, labels, attention_mask)."""
    labels = labels if labels is not None else torch.full_like(input_ids, IGNORE_INDEX)
    attention_mask = (attention_mask if attention_mask is not None
                      else torch.ones_like(input_ids, dtype=torch.bool))
    text_embeds = F.embedding(input_ids, embed_table)  # :429
    queues = {k: deque(v) for k, v in media_embeds.items()}
    bsz = labels.shape[0]
    text = [text_embeds[k][attention_mask[k]] for k in range(bsz)]  # :447
    labs = [labels[k][attention_mask[k]] for k in range(bsz)]
    tok2name = {tid: name for name, tid in media_token_ids.items()}
    inputs_m, labels_m = [], []
    for k in range(bsz):
        ids_k = input_ids[k].tolist()  # NOTE the reference indexes the UNMASKED input_ids (:463)
        ins, lbs = [], []
        pos = 0
        while pos < len(labs[k]):
            if ids_k[pos] in tok2name:
                end = pos + 1
                inp = queues[tok2name[ids_k[pos]]].popleft()
                lab = torch.full([inp.shape[0]], IGNORE_INDEX, dtype=labs[k].dtype, device=labs[k].device)
            else:
                end = pos
                while end < len(labs[k]) and ids_k[end] not in tok2name:
                    end += 1
                inp = text[k][pos:end]
                lab = labs[k][pos:end]
            ins.append(inp)
            lbs.append(lab)
            pos = end
        inputs_m.append(torch.cat(ins, dim=0))
        labels_m.append(torch.cat(lbs, dim=0))
    for name, q in queues.items():
        if q:
            raise ValueError(f"Not all {name} embeddings are consumed!")  # :484
    # __batchify_sequence
    hidden = inputs_m[0].shape[1]
    max_len = max(x.shape[0] for x in inputs_m)
    mask = torch.ones((bsz, max_len), dtype=torch.bool, device=inputs_m[0].device)
    ins_p, lab_p = [], []
    for k in range(bsz):
        n = inputs_m[k].shape[0]
        pad_i = torch.zeros((max_len - n, hidden), dtype=inputs_m[k].dtype, device=inputs_m[k].device)
        pad_l = torch.full((max_len - n,), IGNORE_INDEX, dtype=labels_m[k].dtype, device=labels_m[k].device)
        if padding_side == "right":
            mask[k, n:] = False
            ins_p.append(torch.cat([inputs_m[k], pad_i], 0))
            lab_p.append(torch.cat([labels_m[k], pad_l], 0))
        else:
            mask[k, :max_len - n] = False
            ins_p.append(torch.cat([pad_i, inputs_m[k]], 0))
            lab_p.append(torch.cat([pad_l, labels_m[k]], 0))
    return torch.stack(ins_p, 0), torch.stack(lab_p, 0), mask


# =================================================================================================
# Qwen2 LLM (third-party transformers==4.46.0; formulas per the in-tree copy
# llava/eval/vision_niah_vila/zigzag_ring_attn/modeling_qwen2.py)
# =================================================================================================
def rms_norm(x, w, eps: float):
    """Qwen2RMSNorm.forward (modeling_qwen2.py:89-95)."""
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return w * h.to(dt)


def rope_inv_freq(head_dim: int, theta: float) -> torch.Tensor:
    """Qwen2RotaryEmbedding.__init__ (modeling_qwen2.py:99-111)."""
    return 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))


def rope_cos_sin(position_ids: torch.Tensor, head_dim: int, theta: float, dtype):
    """cos/sin caches (modeling_qwen2.py:113-134): fp32 outer product, cat, cos/sin, cast."""
    inv = rope_inv_freq(head_dim, theta).to(position_ids.device)
    freqs = position_ids.to(torch.float32)[:, None] * inv[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    """modeling_qwen2.py:137-141."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q, k, cos, sin):
    """apply_rotary_pos_emb (modeling_qwen2.py:144-160). q,k: [H, S, D]; cos/sin: [S, D]."""
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


SCORE_BUDGET = 1 << 28  # elements of one attention-score block (memory bound only; the result does not depend on it)


def qwen2_attention(x, p, prefix: str, cfg: Qwen2Cfg, position_ids, past_kv=None):
    """Qwen2Attention.forward (modeling_qwen2.py:191-310): q/k/v proj (+bias), RoPE, KV cache
    append, repeat_kv (:179-188), causal softmax(QK^T/sqrt(d)) in fp32, PV, o_proj (no bias)."""
    S = x.shape[0]
    H, Hk, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    q = F.linear(x, p[prefix + "q_proj.weight"], p[prefix + "q_proj.bias"]).view(S, H, D).transpose(0, 1)
    k = F.linear(x, p[prefix + "k_proj.weight"], p[prefix + "k_proj.bias"]).view(S, Hk, D).transpose(0, 1)
    v = F.linear(x, p[prefix + "v_proj.weight"], p[prefix + "v_proj.bias"]).view(S, Hk, D).transpose(0, 1)
    cos, sin = rope_cos_sin(position_ids, D, cfg.rope_theta, x.dtype)
    q, k = apply_rope(q, k, cos, sin)
    if past_kv is not None:
        k = torch.cat([past_kv[0], k], dim=1)
        v = torch.cat([past_kv[1], v], dim=1)
    new_kv = (k, v)
    rep = H // Hk
    kk = k.repeat_interleave(rep, dim=0)
    vv = v.repeat_interleave(rep, dim=0)
    Sk = kk.shape[1]
    # Same arithmetic per (head, query row); heads and query rows are processed in blocks only to bound
    # the [h, q, Sk] score tensor for long sequences (64-frame video: S = 16.5K -> 1 GB per head in fp32;
    # 256 frames: S = 65.8K -> 17 GB per head, hence the row blocks).
    hstep = max(1, min(H, SCORE_BUDGET // max(1, S * Sk)))
    qstep = S if S * Sk <= SCORE_BUDGET else max(1, SCORE_BUDGET // Sk)
    kv_idx = torch.arange(Sk, device=x.device)
    outs = []
    for h0 in range(0, H, hstep):
        rows = []
        for q0 in range(0, S, qstep):
            q1 = min(S, q0 + qstep)
            att = torch.matmul(q[h0:h0 + hstep, q0:q1], kk[h0:h0 + hstep].transpose(1, 2)) / math.sqrt(D)  # :273
            # causal mask: query row i (global position Sk - S + i) sees kv <= its position
            mask = torch.zeros((q1 - q0, Sk), dtype=torch.float32, device=x.device)
            mask.masked_fill_(kv_idx[None, :] > (torch.arange(q0, q1, device=x.device)[:, None] + (Sk - S)),
                              float("-inf"))
            att = att + mask.to(att.dtype)
            att = F.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)  # :290
            rows.append(torch.matmul(att, vv[h0:h0 + hstep]))
            del att, mask
        outs.append(torch.cat(rows, dim=1) if len(rows) > 1 else rows[0])
    o = torch.cat(outs, dim=0).transpose(0, 1).reshape(S, H * D)
    return F.linear(o, p[prefix + "o_proj.weight"]), new_kv


def qwen2_mlp(x, p, prefix: str):
    """Qwen2MLP.forward (modeling_qwen2.py:164-176): down(silu(gate(x)) * up(x))."""
    return F.linear(F.silu(F.linear(x, p[prefix + "gate_proj.weight"])) *
                    F.linear(x, p[prefix + "up_proj.weight"]), p[prefix + "down_proj.weight"])


def qwen2_forward(inputs_embeds, p, cfg: Qwen2Cfg, position_ids=None, past=None,
                  last_only: bool = False, return_hidden: bool = False):
    """Qwen2ForCausalLM.forward for one sequence (decoder layer: modeling_qwen2.py:633-706).
    inputs_embeds [S, hidden] -> (logits [S or 1, V], new_past)."""
    S = inputs_embeds.shape[0]
    past_len = 0 if past is None else past[0][0].shape[1]
    if position_ids is None:
        position_ids = torch.arange(past_len, past_len + S, device=inputs_embeds.device)
    x = inputs_embeds
    new_past = []
    for i in range(cfg.num_hidden_layers):
        pre = f"model.layers.{i}."
        h = rms_norm(x, p[pre + "input_layernorm.weight"], cfg.rms_norm_eps)
        a, kv = qwen2_attention(h, p, pre + "self_attn.", cfg, position_ids,
                                None if past is None else past[i])
        new_past.append(kv)
        x = x + a
        h = rms_norm(x, p[pre + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        x = x + qwen2_mlp(h, p, pre + "mlp.")
    hidden = x  # last decoder layer's output (input of the final norm)
    x = rms_norm(x, p["model.norm.weight"], cfg.rms_norm_eps)
    if last_only:
        x = x[-1:]
    if return_hidden:
        return F.linear(x, p["lm_head.weight"]), new_past, hidden
    return F.linear(x, p["lm_head.weight"]), new_past


def greedy_generate(inputs_embeds, p, cfg: Qwen2Cfg, max_new_tokens: int,
                    eos_token_ids: Sequence[int] = ()):
    """HF GenerationMixin greedy loop as called from llava_arch.py:833 (`inputs_embeds=` prompt,
    DynamicCache, argmax, stop on eos).  Returns only the NEW ids, and the per-step logits."""
    logits, past = qwen2_forward(inputs_embeds, p, cfg, last_only=True)
    ids, all_logits = [], []
    for _ in range(max_new_tokens):
        all_logits.append(logits[-1])
        tok = int(torch.argmax(logits[-1].float()))
        ids.append(tok)
        if tok in eos_token_ids:
            break
        emb = p["model.embed_tokens.weight"][tok][None, :]
        logits, past = qwen2_forward(emb, p, cfg, past=past, last_only=True)
    return ids, torch.stack(all_logits)


# =================================================================================================
# end-to-end: LlavaLlamaModel.forward / generate for one sample
# =================================================================================================
@dataclass
class VilaOracleModel:
    """Holds the three state dicts with the reference's names (`llm.*`, `vision_tower.vision_tower.*`,
    `mm_projector.*`; llava_arch.py:158-204) and evaluates the path."""
    vcfg: SiglipCfg
    lcfg: Qwen2Cfg
    projector_type: str
    vision: Dict[str, torch.Tensor]
    proj: Dict[str, torch.Tensor]
    llm: Dict[str, torch.Tensor]
    image_token_id: int = 151649
    newline_token_ids: Tuple[int, ...] = (198,)  # tokenizer("\n").input_ids for Qwen2
    dynamic_s2: bool = False
    s2_scales: Tuple[int, ...] = (448, 896, 1344)
    s2_resize_output_to_scale_idx: int = -1

    def tower(self, images):
        return siglip_tower(images, self.vision, self.vcfg, -2)

    def project(self, feats):
        return projector(feats, self.proj, self.projector_type)

    def encode_images(self, images, block_sizes=None):
        return encode_images(images, self.tower, self.project, dynamic_s2=self.dynamic_s2,
                             block_sizes=block_sizes, scales=self.s2_scales,
                             resize_output_to_scale_idx=self.s2_resize_output_to_scale_idx)

    def embed(self, input_ids, images: List[torch.Tensor], block_sizes=None):
        """_embed for media = {"image": images} (llava_arch.py:412-490)."""
        table = self.llm["model.embed_tokens.weight"]
        end = F.embedding(torch.tensor(self.newline_token_ids, device=table.device), table)
        media = {}
        if images:
            feats = self.encode_images(torch.stack(images, 0), block_sizes)
            media["image"] = image_encoder(list(feats), end)
        return embed_splice(input_ids.to(table.device), table, media, {"image": self.image_token_id})

    def forward_logits(self, input_ids, images, block_sizes=None):
        emb, _, _ = self.embed(input_ids, images, block_sizes)
        logits, _ = qwen2_forward(emb[0], self.llm, self.lcfg)
        return logits

    def generate(self, input_ids, images, max_new_tokens: int, eos=(), block_sizes=None):
        emb, _, _ = self.embed(input_ids, images, block_sizes)
        return greedy_generate(emb[0], self.llm, self.lcfg, max_new_tokens, eos)
