"""Shared helpers for the parity tests (oracle side only — never imported by vila_b200)."""
from __future__ import annotations

from typing import Dict, List

import torch

from oracle import vila_oracle as O


def oracle_from_state_dict(sd: Dict[str, torch.Tensor], cfg, dtype=torch.float32) -> O.VilaOracleModel:
    """Build the CPU oracle from a reference-named state dict (values are bf16 on any device)."""
    def sub(prefix):
        return {k[len(prefix):]: v.detach().to("cpu").to(dtype) for k, v in sd.items()
                if k.startswith(prefix)}

    v, l = cfg.vision_tower_cfg, cfg.llm_cfg
    return O.VilaOracleModel(
        vcfg=O.SiglipCfg(v.hidden_size, v.intermediate_size, v.num_hidden_layers, v.num_attention_heads,
                         v.image_size, v.patch_size, v.num_channels, v.layer_norm_eps),
        lcfg=O.Qwen2Cfg(l.hidden_size, l.intermediate_size, l.num_hidden_layers, l.num_attention_heads,
                        l.num_key_value_heads, l.vocab_size, l.rms_norm_eps, l.rope_theta, l.head_dim),
        projector_type=cfg.mm_projector_type,
        vision=sub("vision_tower.vision_tower."), proj=sub("mm_projector."), llm=sub("llm."),
        image_token_id=cfg.image_token_id, newline_token_ids=tuple(cfg.newline_token_ids),
        dynamic_s2=cfg.dynamic_s2, s2_scales=tuple(sorted(cfg.s2_scales)),
        s2_resize_output_to_scale_idx=cfg.s2_resize_output_to_scale_idx)


def check_close(name: str, got: torch.Tensor, truth: torch.Tensor, ref_lowp: torch.Tensor = None,
                rel_floor: float = 1e-3, factor: float = 2.0) -> float:
    """Tolerance model for bf16 kernels (stated once here, used by every e2e parity test):

        err(got) <= factor * err(reference's own bf16 path) + rel_floor * max|truth|

    `truth` is the oracle in fp32 on bf16-rounded weights; `ref_lowp` is the oracle run in bf16 with
    the reference's rounding points (unfused torch ops).  I.e. the CUDA path may not be further from
    the fp32 truth than twice the reference's own bf16 noise (+1e-3 relative).  Without `ref_lowp`
    the bound is 2^-7 relative (one bf16 ulp of the largest value) + rel_floor."""
    got, truth = got.detach().float().cpu(), truth.detach().float().cpu()
    scale = truth.abs().max().item()
    err = (got - truth).abs().max().item()
    if ref_lowp is not None:
        ref_err = (ref_lowp.detach().float().cpu() - truth).abs().max().item()
        bound = factor * ref_err + rel_floor * scale
    else:
        bound = (2 ** -7 + rel_floor) * scale
    assert err <= bound, f"{name}: err {err:.4e} > bound {bound:.4e} (scale {scale:.3e})"
    return err


def greedy_ids_match(got: List[int], oracle_ids: List[int], oracle_logits: torch.Tensor,
                     margin: float) -> None:
    """Greedy token ids must be identical, except that the comparison stops at the first step where
    the oracle's own top-2 logit margin is below `margin` (a bf16-level tie: either choice is a
    faithful greedy decode and the continuations legitimately differ)."""
    for i, (a, b) in enumerate(zip(got, oracle_ids)):
        if a == b:
            continue
        top2 = torch.topk(oracle_logits[i].float(), 2).values
        gap = (top2[0] - top2[1]).item()
        assert gap < margin, f"token {i}: got {a}, oracle {b}, oracle top-2 margin {gap:.4f} >= {margin}"
        return
