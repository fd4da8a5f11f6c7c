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
MEDIA_TOKENS = {"image": "<image>", "video": "<vila/video>"}   # llava/constants.py:32-35
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


def load_video_frames(path: str, num_frames: int, fps: float = 0.0) -> list:
    """llava/utils/media.py:40-85 `_load_video`: a directory of frame images or a video file (cv2).
    fps <= 0: num_frames indices spread uniformly over the clip; fps > 0: one frame every 1/fps
    seconds, at most num_frames.  Frames that fail to decode are skipped, duplicates decoded once."""
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
    video_fps = cap.get(cv2.CAP_PROP_FPS)
    # the container's frame count can overshoot: walk back to the last frame that can be grabbed
    count = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    while count > 0:
        cap.set(cv2.CAP_PROP_POS_FRAMES, count - 1)
        if cap.grab():
            break
        count -= 1
    if count <= 0:
        raise ValueError(f"Video '{path}' has no frames.")
    if fps > 0:
        duration = count / video_fps if video_fps > 0 else 0
        stamps = np.arange(0, duration, 1.0 / fps)[:num_frames]
        indices = [int(t * video_fps) for t in stamps]
    else:
        indices = [int(i) for i in np.round(np.linspace(0, count - 1, num_frames)).astype(int)]
    frames = {}
    for i in indices:
        if i in frames:
            continue
        cap.set(cv2.CAP_PROP_POS_FRAMES, i)
        ok, frame = cap.read()
        if ok:
            frames[i] = PIL.Image.fromarray(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
    return [frames[i] for i in indices if i in frames]


def extract_media(prompt: Union[str, list], config: LlavaConfig):
    """Flatten a prompt (str | list of str / PIL images / image tensors / llava.Image / llava.Video)
    into (text, images) the way llava/utils/media.py:93-123 does: media tokens typed into a text part
    are removed (and that part stripped), every image contributes one bare `<image>`, a video
    contributes `num_video_frames` of them and its frames join the image list.  (The "\n" after an
    image is not text: it is the image encoder's end token, encoders/image/basic.py.)"""
    import PIL.Image
    parts = list(prompt) if isinstance(prompt, (list, tuple)) else [prompt]
    text, images = "", []
    for part in parts:
        if isinstance(part, str):
            for token in MEDIA_TOKENS.values():
                if token in part:
                    part = part.replace(token, "").strip()
            text += part
        elif isinstance(part, Video):
            images.extend(load_video_frames(part.path, config.num_video_frames, getattr(config, "fps", 0.0)))
            text += DEFAULT_IMAGE_TOKEN * config.num_video_frames
        elif isinstance(part, (Image, PIL.Image.Image, torch.Tensor)):
            if isinstance(part, Image):
                if part.path.startswith(("http://", "https://")):
                    import requests
                    part = PIL.Image.open(requests.get(part.path, stream=True).raw)
                else:
                    part = PIL.Image.open(part.path)
            images.append(part)
            text += DEFAULT_IMAGE_TOKEN
        else:
            raise ValueError(f"Unsupported prompt part type: {type(part)}")
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
    text = text.strip()  # llava/utils/tokenizer.py:80-82 normalises every message before the template
    if hasattr(tokenizer, "apply_chat_template"):
        rendered = tokenizer.apply_chat_template([{"role": "user", "content": text}],
                                                 add_generation_prompt=True, tokenize=False)
        return tokenizer(rendered).input_ids
    return tokenizer(text).input_ids


# ------------------------------------------------------------------------------------------------
# GPU preprocessing: PIL's bicubic resize + rescale + normalise + tiling as kernels.
# The geometry (which grids, which tiles) stays above; what moves to the device is the per-pixel work.
# Pillow's 8-bit resampler (src/libImaging/Resample.c) is a separable fixed-point convolution: double
# precision filter weights, normalised, rounded to 22-bit integers, horizontal pass -> uint8 ->
# vertical pass -> uint8.  The tables below are computed on the host exactly like precompute_coeffs /
# normalize_coeffs_8bpc do, the two passes run in vila_resize_bicubic_tiles, so the tiles are
# bit-identical to PIL + SiglipImageProcessor (tests/test_preprocess_*.py).
# ------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_coeffs(in_size: int, out_size: int):
    """-> (ksize, bounds [out,2] (xmin, count), coeffs [out, ksize] int32) of Pillow's
    precompute_coeffs + normalize_coeffs_8bpc for the full-image box and the bicubic filter."""
    import math
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, coeffs = [], []
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        row = [int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS)) for w in k]
        coeffs.append(row + [0] * (ksize - xmax))
        bounds.append((xmin, xmax))
    return ksize, torch.tensor(bounds, dtype=torch.int32), torch.tensor(coeffs, dtype=torch.int32)


def tiling_plan(width: int, height: int, config: LlavaConfig, max_tiles: Optional[int] = None):
    """The resize jobs of process_image for one image of (width, height): a list of
    (out_w, out_h) grids in tile order, plus block_size (dynamic_s2) — the same decisions as
    dynamic_preprocess / dynamic_s2_preprocess / the plain `resize` path, without touching pixels."""
    size = config.vision_tower_cfg.image_size
    ar = config.image_aspect_ratio
    if ar == "dynamic_s2":
        scales = sorted(config.s2_scales)
        jobs = [(size * (s // scales[0]), size * (s // scales[0])) for s in scales[:-1]]
        min_num = (scales[-1] // scales[0]) ** 2
        best = find_closest_aspect_ratio(width / height, target_ratios(min_num, max_tiles or config.max_tiles),
                                         width, height, size)
        jobs.append((size * best[0], size * best[1]))
        return jobs, (best[1], best[0])
    if ar == "dynamic":
        best = find_closest_aspect_ratio(width / height, target_ratios(config.min_tiles, max_tiles or config.max_tiles),
                                         width, height, size)
        jobs = [(size * best[0], size * best[1])]
        if best[0] * best[1] != 1:
            jobs.append((size, size))  # thumbnail
        return jobs, None
    return [(size, size)], None


def process_image_gpu(image, config: LlavaConfig, device="cuda", max_tiles: Optional[int] = None):
    """process_image on the device: PIL image (or uint8 HWC array / tensor) -> bf16 tiles
    [n, 3, S, S] (+ block_size for dynamic_s2), bit-identical to the PIL + SiglipImageProcessor path.
    One H2D copy of the raw uint8 image, two kernels per resize grid.  `pad` keeps its host path
    (a paste on a mean-coloured canvas) and feeds the padded image to the same kernels."""
    import numpy as np

    from .. import ops
    if hasattr(image, "convert"):
        image = image.convert("RGB")
        if config.image_aspect_ratio == "pad":
            image = expand2square(image, tuple(int(x * 255) for x in (SIGLIP_MEAN,) * 3))
        arr = torch.from_numpy(np.asarray(image).copy())
    else:
        arr = torch.as_tensor(image)
    assert arr.dtype == torch.uint8 and arr.dim() == 3 and arr.shape[2] == 3
    H, W = int(arr.shape[0]), int(arr.shape[1])
    src = arr.to(device, non_blocking=True)
    size = config.vision_tower_cfg.image_size
    jobs, block_size = tiling_plan(W, H, config, max_tiles)
    n_tiles = sum((w // size) * (h // size) for w, h in jobs)
    out = torch.empty((n_tiles, 3, size, size), dtype=torch.bfloat16, device=device)
    t0 = 0
    for (ow, oh) in jobs:
        ops.resize_bicubic_tiles(src, ow, oh, out, size, t0, SIGLIP_MEAN, SIGLIP_STD)
        t0 += (ow // size) * (oh // size)
    return (out, block_size) if config.image_aspect_ratio == "dynamic_s2" else out


def process_images_gpu(images: list, config: LlavaConfig, device="cuda"):
    """process_images (generate_content's media processing) with the per-pixel work on the device:
    -> (list of bf16 [3,S,S] device tensors, block_sizes or None), same results as process_images."""
    ar = config.image_aspect_ratio
    if len(images) == 1 and ar in ("dynamic", "dynamic_s2"):
        if ar == "dynamic":
            tiles = process_image_gpu(images[0], config, device)
            return [t for t in tiles], None
        tiles, bs = process_image_gpu(images[0], config, device)
        return [t for t in tiles], [bs]
    plain = LlavaConfig(**{**config.__dict__, "image_aspect_ratio": "pad" if ar == "pad" else "resize"})
    out = [process_image_gpu(im, plain, device)[0] for im in images]
    return out, ([None] * len(out) if config.dynamic_s2 else None)
