"""Configuration objects mirroring the fields of the reference's LlavaConfig that the hot path reads
(llava/model/configuration_llava.py:23-112): `mm_vision_select_layer`,
`mm_vision_select_feature`, `dynamic_s2`, `s2_scales`, `s2_resize_output_to_scale_idx`,
`mm_projector_type`, `image_aspect_ratio`, `num_video_frames`, `model_dtype`.
Architecture constants come from the reference launch scripts (scripts/NVILA/*.sh).
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import List, Optional, Tuple


@dataclass
class SiglipVisionConfig:
    hidden_size: int = 1152
    intermediate_size: int = 4304
    num_hidden_layers: int = 27
    num_attention_heads: int = 16
    image_size: int = 448
    patch_size: int = 14
    num_channels: int = 3
    layer_norm_eps: float = 1e-6
    hidden_act: str = "gelu_pytorch_tanh"
    model_type: str = "siglip_vision_model"

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid


@dataclass
class Qwen2Config:
    hidden_size: int = 3584
    intermediate_size: int = 18944
    num_hidden_layers: int = 28
    num_attention_heads: int = 28
    num_key_value_heads: int = 4
    vocab_size: int = 152064
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    max_position_embeddings: int = 32768
    tie_word_embeddings: bool = False
    model_type: str = "qwen2"

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class LlavaConfig:
    """Top-level config (reference: LlavaConfig with llm_cfg / vision_tower_cfg / mm_projector_cfg)."""
    llm_cfg: Qwen2Config = field(default_factory=Qwen2Config)
    vision_tower_cfg: SiglipVisionConfig = field(default_factory=SiglipVisionConfig)
    mm_projector_type: str = "mlp_downsample"
    mm_vision_select_layer: int = -2
    mm_vision_select_feature: str = "cls_patch"
    image_aspect_ratio: str = "resize"
    dynamic_s2: bool = False
    s2_scales: Tuple[int, ...] = (448, 896, 1344)
    s2_max_split_size: int = 448
    s2_resize_output_to_scale_idx: int = -1
    min_tiles: int = 1
    max_tiles: int = 12
    num_video_frames: int = 8
    video_max_tiles: int = 1
    video_encoder: str = "basic"            # "basic" | "tsp"
    tsp_pool_sizes: Tuple[Tuple[int, int, int], ...] = ((8, 1, 1),)
    model_dtype: str = "torch.bfloat16"
    model_max_length: int = 32768
    # token ids (Qwen2 tokenizer + VILA media tokens; llava/model/language_model/builder.py:207-211)
    image_token_id: int = 151649
    video_token_id: int = 151650
    newline_token_ids: Tuple[int, ...] = (198,)
    eos_token_ids: Tuple[int, ...] = (151645,)
    pad_token_id: int = 151643

    @property
    def hidden_size(self) -> int:
        return self.llm_cfg.hidden_size

    @property
    def mm_hidden_size(self) -> int:
        n = len(self.s2_scales) if self.dynamic_s2 else 1
        return self.vision_tower_cfg.hidden_size * n

    def to_dict(self):
        return asdict(self)


def nvila_8b(**kw) -> LlavaConfig:
    """NVILA-8B: SigLIP-so400m/14-448 + mlp_downsample + Qwen2.5-7B (scripts/NVILA/stage1_9tile.sh)."""
    return LlavaConfig(**kw)


def nvila_8b_dynamic_s2(**kw) -> LlavaConfig:
    return LlavaConfig(dynamic_s2=True, image_aspect_ratio="dynamic_s2", **kw)


def nvila_video_8b(**kw) -> LlavaConfig:
    """NVILA-Video-8B: mlp_downsample_2x2_fix, 64 frames (scripts/NVILA/stage4.sh:42,50-51)."""
    return LlavaConfig(mm_projector_type="mlp_downsample_2x2_fix", num_video_frames=64, **kw)


def nvila_lite_3b(**kw) -> LlavaConfig:
    """NVILA-Lite-3B: mlp_downsample_3x3_fix + (assumed) Qwen2.5-3B (scripts/NVILA-Lite/sft.sh)."""
    llm = Qwen2Config(hidden_size=2048, intermediate_size=11008, num_hidden_layers=36,
                      num_attention_heads=16, num_key_value_heads=2, vocab_size=151936,
                      tie_word_embeddings=True)
    return LlavaConfig(llm_cfg=llm, mm_projector_type="mlp_downsample_3x3_fix",
                       image_aspect_ratio="dynamic", **kw)


def tiny_test_config(projector: str = "mlp_downsample", dynamic_s2: bool = False,
                     image_size: int = 112, vis_layers: int = 3, llm_layers: int = 2,
                     heads: Tuple[int, int] = (4, 2), **kw) -> LlavaConfig:
    """Small architecture with the SAME head dims as NVILA (ViT d=72, LLM d=128) for parity tests."""
    vis = SiglipVisionConfig(hidden_size=144, intermediate_size=272, num_hidden_layers=vis_layers,
                             num_attention_heads=2, image_size=image_size, patch_size=14)
    llm = Qwen2Config(hidden_size=128 * heads[0], intermediate_size=1024, num_hidden_layers=llm_layers,
                      num_attention_heads=heads[0], num_key_value_heads=heads[1], vocab_size=1024)
    scales = (image_size, image_size * 2, image_size * 3)
    return LlavaConfig(llm_cfg=llm, vision_tower_cfg=vis, mm_projector_type=projector,
                       dynamic_s2=dynamic_s2, s2_scales=scales, s2_max_split_size=image_size,
                       image_token_id=1000, video_token_id=1001, newline_token_ids=(13,),
                       eos_token_ids=(2,), pad_token_id=0, **kw)
