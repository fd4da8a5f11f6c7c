"""Oracle (test infrastructure only) for the preprocessing kernel: numpy restatement of Pillow's 8-bit
bicubic resampler (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc) followed by SiglipImageProcessor's
rescale (1/255) and normalise ((x-0.5)/0.5), which is what `processor.preprocess` does to every tile
in the reference (llava/mm_utils.py:476,480,505,518).  Pillow is a third-party dependency of the
reference (not vendored under /root/reference); this restatement is pinned against the installed PIL
itself in tests/test_preprocess_cpu.py (bit-exact on uint8), so parity here is pinned."""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs(in_size: int, out_size: int):
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 2.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) / fs * 1.0) if False else _bicubic((x + xmin - center + 0.5) * (1.0 / fs))
             for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """one separable pass along `axis` of a uint8 [H, W, C] image"""
    in_size = img.shape[axis]
    bounds, kk = coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)           # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n], src[xmin:xmin + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), BICUBIC) on a uint8 [H, W, 3] array: horizontal pass, then
    vertical pass on the uint8 intermediate (Resample.c ImagingResampleInner)."""
    tmp = _pass(img, out_w, axis=1)
    return _pass(tmp, out_h, axis=0)


def siglip_normalise(u8_hwc: np.ndarray) -> np.ndarray:
    """uint8 [H, W, 3] -> float32 [3, H, W]: x/255 then (x - 0.5) / 0.5 in fp32"""
    x = u8_hwc.astype(np.float32) / np.float32(255.0)
    return ((x - np.float32(0.5)) / np.float32(0.5)).transpose(2, 0, 1)
