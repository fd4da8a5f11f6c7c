"""CPU: the preprocessing oracle (oracle/pil_resample.py: Pillow's fixed-point bicubic resampler) is
bit-exact against the installed PIL, and the product's host-side coefficient tables / tiling plan
(vila_b200.model.media) agree with it — the GPU kernel is then checked against both (-m gpu)."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import pil_resample as R


@pytest.mark.parametrize("w,h,ow,oh", [(640, 480, 448, 448), (1600, 800, 1344, 896), (97, 131, 448, 448),
                                       (3000, 500, 1792, 448), (448, 448, 448, 448), (333, 1000, 896, 896),
                                       (500, 448, 448, 448)])
def test_oracle_resampler_is_bit_exact_vs_pil(w, h, ow, oh):
    img = np.random.RandomState(w + h).randint(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    got = R.resize_bicubic_u8(img, ow, oh)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_product_coefficient_tables_match_oracle():
    from vila_b200.model import media
    for in_size, out_size in [(640, 448), (3000, 1792), (97, 448), (448, 448), (131, 448), (1080, 896)]:
        ksize, bounds, coeffs = media.bicubic_coeffs(in_size, out_size)
        b, kk = R.coeffs(in_size, out_size)
        assert ksize == kk.shape[1]
        assert np.array_equal(bounds.numpy(), b) and np.array_equal(coeffs.numpy(), kk)
        assert coeffs.dtype == torch.int32 and int(coeffs.abs().max()) < 2 ** 31


def test_tiling_plan_matches_the_pil_path():
    """same grids / tile counts / block sizes as dynamic_preprocess / dynamic_s2_preprocess"""
    from vila_b200.model import media, nvila_8b_dynamic_s2, nvila_lite_3b, nvila_8b
    rng = np.random.RandomState(5)
    for (w, h) in [(448, 448), (1600, 800), (800, 1600), (333, 1000), (1920, 1080), (97, 131), (3000, 500)]:
        img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8))
        cfg = nvila_8b_dynamic_s2()
        tiles, bs = media.dynamic_s2_preprocess(img, sorted(cfg.s2_scales), cfg.max_tiles, 448)
        jobs, bs2 = media.tiling_plan(w, h, cfg)
        assert tuple(bs) == tuple(bs2) and sum((a // 448) * (b // 448) for a, b in jobs) == len(tiles)
        cfg = nvila_lite_3b()
        tiles = media.dynamic_preprocess(img, cfg.min_tiles, cfg.max_tiles, 448)
        jobs, _ = media.tiling_plan(w, h, cfg)
        assert sum((a // 448) * (b // 448) for a, b in jobs) == len(tiles)
        assert media.tiling_plan(w, h, nvila_8b())[0] == [(448, 448)]


def test_siglip_normalise_matches_to_tensor():
    from vila_b200.model import media
    img = Image.fromarray(np.random.RandomState(2).randint(0, 256, (448, 448, 3), dtype=np.uint8))
    a = media._to_tensor(img, 448)
    b = torch.from_numpy(R.siglip_normalise(np.asarray(img)))
    assert torch.equal(a, b)
