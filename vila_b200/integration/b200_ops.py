"""Operator-level binding of libvila_b200.so as a maintainer of the reference would add it
(INTEGRATION.md option B; in the reference tree this file would be llava/model/b200_ops.py).

Self-contained on purpose: plain ctypes on the C-ABI of include/vila_b200.h, no import of the
vila_b200 Python package.  Two seams of the reference are served:
  * `flash_attn_func(q, k, v, dropout_p, softmax_scale, causal)` — the call SiglipFlashAttention2 makes
    (llava/model/multimodal_encoder/siglip/modeling_siglip.py:583-585; import at :43-45);
  * the `attn_implementation=` constructor kwarg (llava/model/language_model/builder.py:67,74;
    llava/train/train.py:560,569): `register_hf_attention()` registers "vila_b200" with transformers'
    attention-function registry, so `AutoModel...(attn_implementation="vila_b200")` routes every
    attention layer (SigLIP non-causal, Qwen2 causal GQA) to vila_fmha.
Executed by tests/test_integration_gpu.py.
"""
import ctypes
import os

import torch

_LIB = os.environ.get("VILA_B200_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      "libvila_b200.so")
_lib = ctypes.CDLL(_LIB)
_lib.vila_last_error.restype = ctypes.c_char_p


class _Fmha(ctypes.Structure):  # == vila_fmha_params in include/vila_b200.h
    _fields_ = [("q", ctypes.c_void_p), ("q_tok_stride", ctypes.c_int64), ("q_head_stride", ctypes.c_int64),
                ("k", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("kv_page_stride", ctypes.c_int64), ("kv_tok_stride", ctypes.c_int64),
                ("kv_head_stride", ctypes.c_int64), ("kv_num_pages", ctypes.c_int64),
                ("page_table", ctypes.c_void_p), ("page_table_stride", ctypes.c_int32),
                ("o", ctypes.c_void_p), ("o_tok_stride", ctypes.c_int64), ("o_head_stride", ctypes.c_int64),
                ("B", ctypes.c_int32), ("Sq", ctypes.c_int32), ("Sk", ctypes.c_int32), ("Hq", ctypes.c_int32),
                ("Hkv", ctypes.c_int32), ("D", ctypes.c_int32), ("causal", ctypes.c_int32),
                ("scale", ctypes.c_float)]


def flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False):
    """flash_attn.flash_attn_func as called at modeling_siglip.py:583-585.
    q [B, Sq, Hq, D], k/v [B, Sk, Hkv, D]: bf16 CUDA tensors, last dim contiguous, batch-major rows."""
    assert dropout_p == 0.0, "inference path: no attention dropout"
    B, Sq, Hq, D = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    q, k, v = (t if t.stride(3) == 1 and t.stride(0) == t.shape[1] * t.stride(1) else t.contiguous() for t in (q, k, v))
    assert k.stride() == v.stride()
    o = torch.empty((B, Sq, Hq, D), dtype=q.dtype, device=q.device)
    p = _Fmha(q.data_ptr(), q.stride(1), q.stride(2), k.data_ptr(), v.data_ptr(),
              0, k.stride(1), k.stride(2), 0, None, 0, o.data_ptr(), o.stride(1), o.stride(2),
              B, Sq, Sk, Hq, Hkv, D, int(causal), softmax_scale if softmax_scale is not None else D ** -0.5)
    if _lib.vila_fmha(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)):
        raise RuntimeError(_lib.vila_last_error().decode())
    return o


def hf_attention_forward(module, query, key, value, attention_mask, dropout=0.0, scaling=None,
                         is_causal=None, **kwargs):
    """transformers attention-function signature (query/key/value [B, H, S, D]) -> (out [B, S, H, D], None).
    Padding masks are not supported here (the reference un-pads before flash-attn as well)."""
    if is_causal is None:
        is_causal = bool(getattr(module, "is_causal", False)) and query.shape[2] > 1
    out = flash_attn_func(query.transpose(1, 2), key.transpose(1, 2), value.transpose(1, 2),
                          softmax_scale=scaling, causal=is_causal)
    return out, None


def register_hf_attention(name: str = "vila_b200") -> str:
    from transformers import AttentionInterface
    AttentionInterface.register(name, hf_attention_forward)
    return name
