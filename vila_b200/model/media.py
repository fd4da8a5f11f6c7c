"""Host-side glue around generate_content (CPU, PIL): media extraction, SigLIP preprocessing,
dynamic / dynamic-S2 tiling, padding and conversation tokenisation.  These are CALLERS of the hot path
(SURVEY §8f.2); the geometry stays on the host exactly like the reference, the per-pixel work
(resize + rescale + normalise of every tile) can run in the preprocessing kernel (vila_b200.ops.
preprocess_tiles) instead of PIL + numpy:
  extract_media               llava/utils/media.py:93-123
  process_image(s)            llava/mm_utils.py:442-541
  find_closest_aspect_ratio   llava/mm_utils.py:283-296
  dynamic_preprocess          llava/mm_utils.py:299-338
  dynamic_s2_preprocess       llava/mm_utils.py:341-405
  expand2square ("pad")       llava/mm_utils.py:482-495
  tokenize_conversation       llava/utils/tokenizer.py:72-115
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch

from .configuration import LlavaConfig

DEFAULT_IMAGE_TOKEN = "<image>"
SIGLIP_MEAN = 0.5  # SiglipImageProcessor: rescale 1/255, normalise mean=std=0.5, bicubic resize
SIGLIP_STD = 0.5


class Media:
    """llava/media.py: path-carrying prompt parts (`llava.Image("a.png")`, `llava.Video("b.mp4")`)."""


class File(Media):
    def __init__(self, path: str) -> None:
        self.path = path


class Image(File):
    pass


class Video(File):
    pass


def load_video_frames(path: str, num_frames: int) -> list:
    """llava/utils/media.py:_load_video: a directory of frame images or a video file (cv2), sampled
    uniformly to num_frames PIL images."""
    import glob
    import os

    import numpy as np
    import PIL.Image
    if os.path.isdir(path):
        paths = sorted(glob.glob(os.path.join(path, "*")))
        idx = np.round(np.linspace(0, len(paths) - 1, num_frames)).astype(int)
        return [PIL.Image.open(paths[i]) for i in idx]
    import cv2
    cap = cv2.VideoCapture(path)
    count = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    if count <= 0:
        raise ValueError(f"Video '{path}' has no frames.")
    frames = {}
    for i in np.round(np.linspace(0, count - 1, num_frames)).astype(int):
        if int(i) in frames:
            continue
        cap.set(cv2.CAP_PROP_POS_FRAMES, int(i))
        ok, frame = cap.read()
        if ok:
            frames[int(i)] = PIL.Image.fromarray(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
    return [frames[k] for k in sorted(frames)]


def extract_media(prompt: Union[str, list], config: LlavaConfig):
    """Flatten a prompt (str | list of str / PIL images / image tensors / llava.Image / llava.Video)
    into text with <image> tokens + the image list (llava/utils/media.py:93-123: a video becomes
    `num_video_frames` images)."""
    if isinstance(prompt, str):
        return prompt, []
    text, images = "", []
    for part in prompt:
        if isinstance(part, str):
            text += part
        elif isinstance(part, Video):
            frames = load_video_frames(part.path, config.num_video_frames)
            images.extend(frames)
            text += (DEFAULT_IMAGE_TOKEN + "\n") * len(frames)
        else:
            if isinstance(part, Image):
                import PIL.Image
                part = PIL.Image.open(part.path)
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


def _split_grid(resized, tw: int, th: int, image_size: int) -> list:
    per_row = tw // image_size
    return [resized.crop(((i % per_row) * image_size, (i // per_row) * image_size,
                          ((i % per_row) + 1) * image_size, ((i // per_row) + 1) * image_size))
            for i in range(per_row * (th // image_size))]


def target_ratios(min_num: int, max_num: int):
    """all (cols, rows) grids with min_num <= cols*rows <= max_num, sorted by area (stable)"""
    return sorted({(i, j) for n in range(min_num, max_num + 1) for i in range(1, n + 1)
                   for j in range(1, n + 1) if min_num <= i * j <= max_num}, key=lambda x: x[0] * x[1])


def dynamic_preprocess(image, min_num: int = 1, max_num: int = 12, image_size: int = 384,
                       use_thumbnail: bool = True) -> list:
    """NVILA-Lite `image_aspect_ratio == "dynamic"` tiling (mm_utils.py:299-338): resize to the
    closest-aspect grid of image_size tiles, split row-major, append a thumbnail when > 1 tile."""
    w0, h0 = image.size
    best = find_closest_aspect_ratio(w0 / h0, target_ratios(min_num, max_num), w0, h0, image_size)
    tw, th = image_size * best[0], image_size * best[1]
    tiles = _split_grid(image.resize((tw, th)), tw, th, image_size)
    assert len(tiles) == best[0] * best[1]
    if use_thumbnail and len(tiles) != 1:
        tiles.append(image.resize((image_size, image_size)))
    return tiles


def expand2square(pil_img, background_color):
    """`image_aspect_ratio == "pad"` (mm_utils.py:482-495): centre the image on a square canvas."""
    from PIL import Image
    width, height = pil_img.size
    if width == height:
        return pil_img
    side = max(width, height)
    result = Image.new(pil_img.mode, (side, side), background_color)
    result.paste(pil_img, ((side - width) // 2, (side - height) // 2))
    return result


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


def process_image(image, config: LlavaConfig, enable_dynamic_res: bool = False,
                  enable_dynamic_s2: bool = False, max_tiles: Optional[int] = None):
    """mm_utils.process_image (:442-522) for a PIL image -> [3,S,S] tensor, or stacked tiles [n,3,S,S]
    (dynamic), or (tiles, block_size) (dynamic_s2)."""
    size = config.vision_tower_cfg.image_size
    image = image.convert("RGB")
    ar = config.image_aspect_ratio
    if "dynamic_s2" in ar and enable_dynamic_s2:
        tiles, bs = dynamic_s2_preprocess(image, sorted(config.s2_scales), getattr(config, "max_tiles", 12), size)
        return torch.stack([_to_tensor(t, size) for t in tiles]), bs
    if "dynamic" in ar and enable_dynamic_res:
        tiles = dynamic_preprocess(image, min_num=getattr(config, "min_tiles", 1),
                                   max_num=max_tiles if max_tiles is not None else getattr(config, "max_tiles", 12),
                                   image_size=size)
        return torch.stack([_to_tensor(t, size) for t in tiles])
    if ar == "resize":
        image = image.resize((size, size))  # PIL default filter (bicubic), then the processor's own resize is a no-op
    if ar == "pad":
        image = expand2square(image, tuple(int(x * 255) for x in (SIGLIP_MEAN,) * 3))
    return _to_tensor(image, size)


def process_images(images: list, config: LlavaConfig, max_tiles: int = 12
                   ) -> Tuple[List[torch.Tensor], Optional[list]]:
    """generate_content's media processing (llava_arch.py:858-880): -> (list of [3,S,S] tensors,
    block_sizes or None).  One PIL image under `dynamic` / `dynamic_s2` is tiled; the caller repeats
    the image token once per `dynamic` tile (media.dynamic_prompt)."""
    if all(isinstance(im, torch.Tensor) for im in images):
        return [im for im in images], None
    ar = config.image_aspect_ratio
    if len(images) == 1 and not isinstance(images[0], torch.Tensor) and ar in ("dynamic", "dynamic_s2"):
        if ar == "dynamic":
            tiles = process_image(images[0], config, enable_dynamic_res=True, max_tiles=max_tiles)
            return [t for t in tiles], None
        tiles, bs = process_image(images[0], config, enable_dynamic_s2=True)
        return [t for t in tiles], [bs]
    out = [im if isinstance(im, torch.Tensor) else process_image(im, config) for im in images]
    return out, ([None] * len(out) if config.dynamic_s2 else None)


def dynamic_prompt(text: str, n_tiles: int) -> str:
    """`dynamic`: every tile is an image of its own in the prompt (llava_arch.py:866-868)."""
    return text.replace(DEFAULT_IMAGE_TOKEN, (DEFAULT_IMAGE_TOKEN + "\n") * n_tiles)


def tokenize_conversation(text: str, tokenizer) -> List[int]:
    """Single human turn + generation prompt. With a real HF tokenizer the chat template is used and
    media tokens are mapped to their ids; the synthetic tokenizer handles both directly."""
    if hasattr(tokenizer, "apply_chat_template"):
        rendered = tokenizer.apply_chat_template([{"role": "user", "content": text}],
                                                 add_generation_prompt=True, tokenize=False)
        return tokenizer(rendered).input_ids
    return tokenizer(text).input_ids
