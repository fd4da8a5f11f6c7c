"""Host-side glue around generate_content (CPU, PIL): media extraction, SigLIP preprocessing,
dynamic-S2 tiling and conversation tokenisation.  These are CALLERS of the hot path;
they stay on the host exactly like the reference:
  extract_media               llava/utils/media.py:93-123
  process_image(s)            llava/mm_utils.py:442-541
  find_closest_aspect_ratio   llava/mm_utils.py:283-296
  dynamic_s2_preprocess       llava/mm_utils.py:341-405
  tokenize_conversation       llava/utils/tokenizer.py:72-115
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch

from .configuration import LlavaConfig

DEFAULT_IMAGE_TOKEN = "<image>"
SIGLIP_MEAN = 0.5  # SiglipImageProcessor: rescale 1/255, normalise mean=std=0.5, bicubic resize
SIGLIP_STD = 0.5


def extract_media(prompt: Union[str, list], config: LlavaConfig):
    """Flatten a prompt (str | list of str / images) into text with <image> tokens + image list."""
    if isinstance(prompt, str):
        return prompt, []
    text, images = "", []
    for part in prompt:
        if isinstance(part, str):
            text += part
        else:
            images.append(part)
            text += DEFAULT_IMAGE_TOKEN + "\n"
    return text, images


def _to_tensor(img, size: int) -> torch.Tensor:
    """PIL image -> normalised [3, size, size] float tensor (SiglipImageProcessor semantics)."""
    import numpy as np
    from PIL import Image

    img = img.convert("RGB").resize((size, size), Image.BICUBIC)
    arr = torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).float() / 255.0
    return (arr - SIGLIP_MEAN) / SIGLIP_STD


def find_closest_aspect_ratio(aspect_ratio, target_ratios, width, height, image_size):
    best_diff, best = float("inf"), (1, 1)
    area = width * height
    for ratio in target_ratios:
        diff = abs(aspect_ratio - ratio[0] / ratio[1])
        if diff < best_diff:
            best_diff, best = diff, ratio
        elif diff == best_diff and area > 0.5 * image_size * image_size * ratio[0] * ratio[1]:
            best = ratio
    return best


def dynamic_s2_preprocess(image, s2_scales, max_num: int, image_size: int):
    """Tiles for every scale but the last on a square grid, then the last scale on the closest
    aspect-ratio grid; returns (tiles, (rows, cols)) like the reference."""
    w0, h0 = image.size
    min_num = (s2_scales[-1] // s2_scales[0]) ** 2
    tiles = []

    def split(resized, tw, th):
        per_row = tw // image_size
        for i in range((tw // image_size) * (th // image_size)):
            box = ((i % per_row) * image_size, (i // per_row) * image_size,
                   ((i % per_row) + 1) * image_size, ((i // per_row) + 1) * image_size)
            tiles.append(resized.crop(box))

    for scale in s2_scales[:-1]:
        k = scale // s2_scales[0]
        split(image.resize((image_size * k, image_size * k)), image_size * k, image_size * k)
    ratios = sorted({(i, j) for n in range(min_num, max_num + 1) for i in range(1, n + 1)
                     for j in range(1, n + 1) if min_num <= i * j <= max_num},
                    key=lambda x: x[0] * x[1])
    best = find_closest_aspect_ratio(w0 / h0, ratios, w0, h0, image_size)
    tw, th = image_size * best[0], image_size * best[1]
    split(image.resize((tw, th)), tw, th)
    return tiles, (best[1], best[0])


def process_images(images: list, config: LlavaConfig, max_tiles: int = 12
                   ) -> Tuple[List[torch.Tensor], Optional[list]]:
    """-> (list of [3,S,S] tensors, block_sizes or None)."""
    size = config.vision_tower_cfg.image_size
    if all(isinstance(im, torch.Tensor) for im in images):
        return [im for im in images], None
    if config.dynamic_s2 and len(images) == 1 and not isinstance(images[0], torch.Tensor):
        tiles, bs = dynamic_s2_preprocess(images[0], list(config.s2_scales), max_tiles, size)
        return [_to_tensor(t, size) for t in tiles], [bs]
    out = [im if isinstance(im, torch.Tensor) else _to_tensor(im, size) for im in images]
    return out, ([None] * len(out) if config.dynamic_s2 else None)


def tokenize_conversation(text: str, tokenizer) -> List[int]:
    """Single human turn + generation prompt. With a real HF tokenizer the chat template is used and
    media tokens are mapped to their ids; the synthetic tokenizer handles both directly."""
    if hasattr(tokenizer, "apply_chat_template"):
        rendered = tokenizer.apply_chat_template([{"role": "user", "content": text}],
                                                 add_generation_prompt=True, tokenize=False)
        return tokenizer(rendered).input_ids
    return tokenizer(text).input_ids
