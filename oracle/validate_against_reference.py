"""Pin the oracle against the reference's own code, executed in the authoring container.

Runs ONLY where /root/reference exists (never on the GPU box, never from tests/ or bench.py).
  * vendored SigLIP (modeling_siglip.py) and MultimodalProjector (base_projector.py) are loaded by
    file path and run on random-init weights;
  * the llava_arch.py glue (merge_chessboard / split_chessboard / merge_features_for_dynamic_s2 /
    encode_images / _embed / __batchify_sequence) and the TSP encoder cannot be imported as a package
    here (deepspeed / hydra / accelerate missing), so their function bodies are extracted from the
    reference source with `ast` and executed unmodified against stub objects;
  * the Qwen2 arithmetic (third-party transformers==4.46.0 in the reference) is checked against the
    installed transformers Qwen2ForCausalLM (sdpa + eager).
Prints max-abs differences; exits non-zero if any exceeds its tolerance.
"""
from __future__ import annotations

import ast
import importlib.util
import sys
import textwrap
import types
from collections import defaultdict, deque
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import vila_oracle as O  # noqa: E402

REF = Path("/root/reference")
FAILED = []


def report(name, diff, tol):
    ok = diff <= tol
    print(f"[{'ok' if ok else 'FAIL'}] {name:58s} max|diff| = {diff:.3e} (tol {tol:.1e})")
    if not ok:
        FAILED.append(name)


def load_by_path(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def extract_functions(path: Path, names):
    """Return {name: source} for (possibly nested-in-class) function defs in a reference file."""
    src = path.read_text()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in out:
            seg = ast.get_source_segment(src, node)
            # drop decorators such as @staticmethod
            out[node.name] = textwrap.dedent(seg)
    return out


# -------------------------------------------------------------------------------------------------
def check_siglip():
    import transformers  # noqa: F401  (import real names before stubbing)
    siglip_dir = REF / "llava/model/multimodal_encoder/siglip"
    ms = load_by_path("ref_modeling_siglip", siglip_dir / "modeling_siglip.py")
    import transformers.models.siglip.configuration_siglip as cfgmod
    torch.manual_seed(0)
    for attn in ("eager", "sdpa"):
        cfg = cfgmod.SiglipVisionConfig(hidden_size=144, intermediate_size=272, num_hidden_layers=4,
                                        num_attention_heads=2, image_size=56, patch_size=14)
        cfg._attn_implementation = attn
        model = ms.SiglipVisionModel(cfg).eval()
        # give biases / LN params non-trivial values
        with torch.no_grad():
            for n, p in model.named_parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.1)
                    if "layer_norm" in n and n.endswith("weight"):
                        p.add_(1.0)
        px = torch.randn(2, 3, 56, 56)
        with torch.no_grad():
            ref = model(px, output_hidden_states=True).hidden_states[-2]
        sd = {k: v for k, v in model.state_dict().items()}
        ocfg = O.SiglipCfg(hidden_size=144, intermediate_size=272, num_hidden_layers=4,
                           num_attention_heads=2, image_size=56, patch_size=14)
        mine = O.siglip_tower(px, sd, ocfg, -2)
        report(f"siglip_tower hidden_states[-2] vs reference ({attn})", (ref - mine).abs().max().item(), 2e-5)
    return ms


def check_projector():
    import transformers  # noqa: F401
    # base_projector imports timm.models.layers.Mlp only for the PS3 head -> 3-line stub
    timm = types.ModuleType("timm"); timm_m = types.ModuleType("timm.models")
    timm_l = types.ModuleType("timm.models.layers")
    timm_l.Mlp = type("Mlp", (torch.nn.Module,), {})
    sys.modules.setdefault("timm", timm); sys.modules.setdefault("timm.models", timm_m)
    sys.modules.setdefault("timm.models.layers", timm_l)
    bp = load_by_path("refprojector", REF / "llava/model/multimodal_projector/base_projector.py")
    torch.manual_seed(1)
    for kind, n_tok in (("mlp_downsample", 64), ("mlp_downsample", 49), ("mlp_downsample_2x2_fix", 49),
                        ("mlp_downsample_3x3_fix", 64), ("mlp_downsample_3x3_fix", 81)):
        cfg = types.SimpleNamespace(mm_hidden_size=48, hidden_size=96)
        pc = bp.MultimodalProjectorConfig(kind)
        model = bp.MultimodalProjector(pc, cfg).eval()
        with torch.no_grad():
            for p in model.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.2)
        x = torch.randn(3, n_tok, 48)
        with torch.no_grad():
            ref = model(x)
        mine = O.projector(x, dict(model.state_dict()), kind)
        report(f"projector {kind} N={n_tok}", (ref - mine).abs().max().item(), 1e-5)
    # flat_square on non-square grids (w != h) directly
    x = torch.randn(2, 5, 7, 8)
    report("flat_square_2x2 5x7", (bp.flat_square_2x2(x) - O.flat_square(x, 2)).abs().max().item(), 0)
    report("flat_square_3x3 5x7", (bp.flat_square_3x3(x) - O.flat_square(x, 3)).abs().max().item(), 0)
    report("DownSampleBlock.flat_square 5x7",
           (bp.DownSampleBlock().flat_square(x) - O.flat_square(x, 2)).abs().max().item(), 0)


def check_arch_glue():
    from einops import rearrange
    names = ["merge_chessboard", "split_chessboard", "merge_features_for_dynamic_s2", "encode_images",
             "_embed", "__embed_media_tokens", "__truncate_sequence", "__batchify_sequence"]
    srcs = extract_functions(REF / "llava/model/llava_arch.py", names)
    ns = {"torch": torch, "F": F, "rearrange": rearrange, "Optional": object, "Tuple": object,
          "Dict": dict, "List": list, "Any": object, "deque": deque, "defaultdict": defaultdict,
          "IGNORE_INDEX": -100, "get_pg_manager": lambda: None, "warnings": __import__("warnings"),
          "chain": __import__("itertools").chain, "distributed": None}
    body = "class RefArch:\n"
    for n in names:
        s = srcs[n].replace("@staticmethod\n", "")
        body += textwrap.indent(("@staticmethod\n" if n in ("merge_chessboard", "split_chessboard") else "") + s, "    ") + "\n"
    exec(compile(body, "<reference llava_arch excerpts>", "exec"), ns)
    RefArch = ns["RefArch"]

    torch.manual_seed(2)
    x = torch.randn(6, 16, 8)
    report("merge_chessboard 2x3", (RefArch.merge_chessboard(x, 2, 3) - O.merge_chessboard(x, 2, 3)).abs().max().item(), 0)
    y = torch.randn(1, 8, 8, 12)
    report("split_chessboard 2x3", (RefArch.split_chessboard(y, 2, 3) - O.split_chessboard(y, 2, 3)).abs().max().item(), 0)

    # dynamic-S2 encode_images with a fake tower / projector
    C_, side = 8, 4
    lin = torch.nn.Linear(3 * 4 * C_, 16)

    class Tower:
        scales = [4, 8, 12]
        resize_output_to_scale_idx = -1

        def __call__(self, images):
            return images  # tests feed features directly

    for idx in (-1, 0, 1):
        tower = Tower(); tower.resize_output_to_scale_idx = idx
        proj = lambda f: lin(O.downsample(f, 2))
        self_ = RefArch()
        self_.config = types.SimpleNamespace(dynamic_s2=True)
        self_.get_vision_tower = lambda: tower
        self_.get_mm_projector = lambda: proj
        block_sizes = [(2, 3), None, (1, 2)]
        n_tiles = (1 + 4 + 6) + 1 + (1 + 4 + 2)
        feats = torch.randn(n_tiles, side * side, C_)
        with torch.no_grad():
            ref = self_.encode_images(feats, block_sizes)
            mine = O.encode_images(feats, tower, proj, dynamic_s2=True, block_sizes=block_sizes,
                                   scales=tower.scales, resize_output_to_scale_idx=idx)
        assert type(ref) is type(mine) and len(ref) == len(mine)
        d = max((a - b).abs().max().item() for a, b in zip(ref, mine))
        report(f"encode_images dynamic_s2 resize_idx={idx}", d, 1e-6)

    # _embed splice
    V, H = 50, 8
    table = torch.randn(V, H)
    self_ = RefArch()
    self_.training = False
    self_.llm = types.SimpleNamespace(model=types.SimpleNamespace(embed_tokens=lambda ids: F.embedding(ids, table)))
    IMG, VID = 40, 41
    self_.tokenizer = types.SimpleNamespace(media_token_ids={"image": IMG, "video": VID},
                                            padding_side="right", model_max_length=4096)
    m_img = [torch.randn(5, H), torch.randn(3, H)]
    m_vid = [torch.randn(7, H)]
    self_.encoders = {"image": lambda media, cfg: list(media), "video": lambda media, cfg: list(media)}
    for side_ in ("right", "left"):
        self_.tokenizer.padding_side = side_
        ids = torch.tensor([[1, 2, IMG, 3, VID, 4, 0, 0], [IMG, 9, 8, 7, 6, 5, 4, 3]])
        am = torch.tensor([[1, 1, 1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1, 1, 1]], dtype=torch.bool)
        ref_in, ref_lab, ref_mask = getattr(self_, "_embed")(ids, {"image": list(m_img), "video": list(m_vid)},
                                                            {"image": {}, "video": {}}, None, am)
        my_in, my_lab, my_mask = O.embed_splice(ids, table, {"image": list(m_img), "video": list(m_vid)},
                                                {"image": IMG, "video": VID}, None, am, side_)
        report(f"_embed splice inputs ({side_})", (ref_in - my_in).abs().max().item(), 0)
        report(f"_embed splice labels/mask ({side_})",
               float((ref_lab != my_lab).sum() + (ref_mask != my_mask).sum()), 0)

    # TSP encoder
    tsp_src = extract_functions(REF / "llava/model/encoders/video/tsp.py", ["pool"])
    ns2 = {"torch": torch}
    exec(tsp_src["pool"], ns2)
    z = torch.randn(8, 4, 4, 6)
    r = z
    for dim, pp in enumerate((4, 2, 2)):
        r = ns2["pool"](r, pp, dim)
    m = z
    for dim, pp in enumerate((4, 2, 2)):
        m = O.tsp_pool(m, pp, dim)
    report("TSP pool (4,2,2)", (r - m).abs().max().item(), 0)


def check_media_preprocess():
    """mm_utils.dynamic_s2_preprocess / find_closest_aspect_ratio (reference, executed from source) vs
    the host-side tiling of vila_b200.model.media (product code, host logic only) on a sweep of image
    sizes: identical block sizes and bit-identical tile pixels."""
    import numpy as np
    from PIL import Image
    from vila_b200.model import media
    srcs = extract_functions(REF / "llava/mm_utils.py", ["find_closest_aspect_ratio", "dynamic_s2_preprocess"])
    ns = {}
    exec(srcs["find_closest_aspect_ratio"], ns)
    exec(srcs["dynamic_s2_preprocess"], ns)
    rng = np.random.RandomState(7)
    worst, n_bad = 0.0, 0
    sizes = [(448, 448), (1600, 800), (800, 1600), (333, 1000), (1920, 1080), (640, 480), (97, 131), (3000, 500)]
    for (w, h) in sizes:
        img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8))
        ref_tiles, ref_bs = ns["dynamic_s2_preprocess"](img, s2_scales=[448, 896, 1344], max_num=12, image_size=448)
        my_tiles, my_bs = media.dynamic_s2_preprocess(img, [448, 896, 1344], 12, 448)
        if tuple(ref_bs) != tuple(my_bs) or len(ref_tiles) != len(my_tiles):
            n_bad += 1
            continue
        for a, b in zip(ref_tiles, my_tiles):
            worst = max(worst, float(np.abs(np.asarray(a, dtype=np.int32) - np.asarray(b, dtype=np.int32)).max()))
    report(f"dynamic_s2_preprocess block sizes ({len(sizes)} image sizes)", float(n_bad), 0)
    report("dynamic_s2_preprocess tile pixels", worst, 0)


def check_pixel_preprocess():
    """processor.preprocess(...) as called by mm_utils.process_image (:476,505) — the reference pins
    transformers==4.46.0 whose SiglipImageProcessor resizes through PIL (bicubic) exactly like
    vila_b200.model.media._to_tensor; the installed 5.x processor interpolates in torch, so agreement is
    exact for 448x448 inputs and within one 8-bit level (2/255 after normalisation) otherwise."""
    import numpy as np
    from PIL import Image
    from transformers import SiglipImageProcessor
    from vila_b200.model import media
    proc = SiglipImageProcessor(size={"height": 448, "width": 448})
    rng = np.random.RandomState(3)
    worst_same, worst_resized = 0.0, 0.0
    for (w, h) in [(448, 448), (640, 480), (97, 131), (1600, 800)]:
        img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8))
        ref = proc.preprocess(img, return_tensors="pt")["pixel_values"][0]
        d = float((ref - media._to_tensor(img, 448)).abs().max())
        if (w, h) == (448, 448):
            worst_same = max(worst_same, d)
        else:
            worst_resized = max(worst_resized, d)
    report("pixel preprocess 448x448 vs SiglipImageProcessor", worst_same, 1e-6)
    report("pixel preprocess resized vs SiglipImageProcessor (1 level)", worst_resized, 2 / 255 + 1e-6)


def check_encoders():
    """BasicImageEncoder._process_features / TSPVideoEncoder._process_features (reference source) vs the
    oracle's image_encoder / tsp_video_encoder."""
    img_src = extract_functions(REF / "llava/model/encoders/image/basic.py", ["_process_features"])["_process_features"]
    tsp_srcs = extract_functions(REF / "llava/model/encoders/video/tsp.py", ["pool", "_process_features"])
    ns = {"torch": torch, "Optional": object}
    exec(tsp_srcs["pool"], ns)
    vid_src = extract_functions(REF / "llava/model/encoders/video/basic.py", ["_process_features"])["_process_features"]
    body = "class Base:\n" + textwrap.indent(img_src, "    ") + "\n"
    body += "class VideoBase:\n" + textwrap.indent(vid_src, "    ") + "\n"
    body += "class TSP(VideoBase):\n" + textwrap.indent(tsp_srcs["_process_features"], "    ") + "\n"
    exec(compile(body, "<reference encoder excerpts>", "exec"), ns)
    torch.manual_seed(11)
    H = 8
    start, end, sep = torch.randn(2, H), torch.randn(1, H), torch.randn(1, H)
    feats = torch.randn(3, 16, H)
    base = ns["Base"]()
    ref = [base._process_features(f, start, end) for f in feats]
    mine = O.image_encoder(list(feats), end, start)
    report("BasicImageEncoder._process_features", max((a - b).abs().max().item() for a, b in zip(ref, mine)), 0)
    vbase = ns["VideoBase"]()
    vfeat = torch.randn(5, 16, H)
    report("BasicVideoEncoder._process_features",
           (vbase._process_features(vfeat, start, end) - O.video_encoder(vfeat, end, start)).abs().max().item(), 0)
    tsp = ns["TSP"]()
    tsp.pool_sizes = [(4, 1, 1), (2, 2, 2)]
    vid = torch.randn(8, 16, H)
    ref_v = tsp._process_features(vid, start, end, sep)
    mine_v = O.tsp_video_encoder(vid, tsp.pool_sizes, end, start, sep)
    report("TSPVideoEncoder._process_features", (ref_v - mine_v).abs().max().item(), 0)


def check_qwen2():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(3)
    for attn in ("eager", "sdpa"):
        hcfg = Qwen2Config(hidden_size=128, intermediate_size=320, num_hidden_layers=3,
                           num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                           rms_norm_eps=1e-6, rope_theta=1000000.0, max_position_embeddings=4096,
                           tie_word_embeddings=False, attn_implementation=attn)
        model = Qwen2ForCausalLM(hcfg).eval()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "norm" in n:
                    p.normal_(1.0, 0.1)
                elif p.dim() == 1:
                    p.normal_(0, 0.1)
        sd = dict(model.state_dict())
        ocfg = O.Qwen2Cfg(hidden_size=128, intermediate_size=320, num_hidden_layers=3,
                          num_attention_heads=4, num_key_value_heads=2, vocab_size=512, head_dim=32)
        emb = torch.randn(1, 37, 128)
        with torch.no_grad():
            ref = model(inputs_embeds=emb).logits[0]
            mine, _ = O.qwen2_forward(emb[0], sd, ocfg)
        report(f"qwen2 forward logits ({attn})", (ref - mine).abs().max().item(), 5e-5)
        with torch.no_grad():
            gen = model.generate(inputs_embeds=emb, attention_mask=torch.ones(1, 37, dtype=torch.long),
                                 max_new_tokens=12, do_sample=False, eos_token_id=None, pad_token_id=0)
            ids, _ = O.greedy_generate(emb[0], sd, ocfg, 12)
        report(f"qwen2 greedy ids ({attn})", float(sum(int(a) != int(b) for a, b in zip(gen[0].tolist(), ids))), 0)
        # large positions (RoPE precision)
        pos = torch.arange(60000, 60037)
        with torch.no_grad():
            ref = model(inputs_embeds=emb, position_ids=pos[None]).logits[0]
            mine, _ = O.qwen2_forward(emb[0], sd, ocfg, position_ids=pos)
        report(f"qwen2 forward logits, positions 60000+ ({attn})", (ref - mine).abs().max().item(), 5e-5)


def ref_packing_namespace():
    """repack_multimodal_data (llava_arch.py:557-800, non-SP branch taken when get_pg_manager() is
    None) and packing._get_unpad_data, executed from the reference source."""
    srcs = extract_functions(REF / "llava/model/llava_arch.py", ["repack_multimodal_data"])
    ns = {"torch": torch, "IGNORE_INDEX": -100, "get_pg_manager": lambda: None, "dist": None}
    exec(srcs["repack_multimodal_data"], ns)
    psrc = extract_functions(REF / "llava/model/utils/packing.py", ["_get_unpad_data"])
    ns2 = {"torch": torch, "F": F, "Tuple": tuple}
    exec("from typing import Tuple\n" + psrc["_get_unpad_data"], ns2)
    return ns["repack_multimodal_data"], ns2["_get_unpad_data"]


def packing_cases():
    g = torch.Generator().manual_seed(11)
    cases = []
    for lens, L, pad_mult in (([5, 3, 7], 8, None), ([1, 1], 4, None), ([6], 6, None), ([4, 9, 2, 9], 9, 8)):
        B, H = len(lens), 6
        emb = torch.randn(B, L, H, generator=g)
        mask = torch.zeros(B, L, dtype=torch.bool)
        for k, n in enumerate(lens):
            mask[k, :n] = True
        labels = torch.randint(0, 50, (B, L), generator=g)
        cases.append((emb, mask, labels, pad_mult))
    return cases


def check_packing():
    from vila_b200.model import packing
    ref_repack, ref_unpad = ref_packing_namespace()
    worst = 0.0
    for emb, mask, labels, pad_mult in packing_cases():
        holder = types.SimpleNamespace(llm=types.SimpleNamespace(pad_token_id=0))
        if pad_mult:
            holder.pad_to_multiple_of = pad_mult
        ref = ref_repack(holder, emb, mask, None, labels.clone())
        mine = packing.repack_multimodal_data(emb, mask, None, labels.clone(), pad_mult, 0)
        for a, b in zip(ref, mine):
            if a.shape != b.shape or a.dtype != b.dtype:
                worst = float("inf")
            else:
                worst = max(worst, (a.double() - b.double()).abs().max().item())
        am = ref[1]
        r_idx, r_cu, r_max = ref_unpad(am)
        m_idx, m_cu, m_max = packing.get_unpad_data(am)
        worst = max(worst, float((r_idx - m_idx).abs().max()), float((r_cu - m_cu).abs().max()), abs(r_max - m_max))
    report("repack_multimodal_data (non-SP) + _get_unpad_data", worst, 0)


def check_dynamic_preprocess():
    """mm_utils.dynamic_preprocess (NVILA-Lite `dynamic`) and expand2square (`pad`) executed from the
    reference source vs vila_b200.model.media."""
    import numpy as np
    from PIL import Image
    from vila_b200.model import media
    srcs = extract_functions(REF / "llava/mm_utils.py", ["find_closest_aspect_ratio", "dynamic_preprocess", "expand2square"])
    ns = {"Image": Image}
    for k in ("find_closest_aspect_ratio", "dynamic_preprocess", "expand2square"):
        exec(srcs[k], ns)
    rng = np.random.RandomState(9)
    worst, n_bad = 0.0, 0
    sizes = [(336, 336), (448, 448), (1600, 800), (800, 1600), (333, 1000), (1920, 1080), (97, 131), (3000, 500)]
    for (w, h) in sizes:
        img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8))
        for max_num in (12, 6):
            ref_tiles = ns["dynamic_preprocess"](img, min_num=1, max_num=max_num, image_size=448)
            my_tiles = media.dynamic_preprocess(img, min_num=1, max_num=max_num, image_size=448)
            if len(ref_tiles) != len(my_tiles):
                n_bad += 1
                continue
            for a, b in zip(ref_tiles, my_tiles):
                worst = max(worst, float(np.abs(np.asarray(a, dtype=np.int32) - np.asarray(b, dtype=np.int32)).max()))
        a = ns["expand2square"](img, (127, 127, 127))
        b = media.expand2square(img, (127, 127, 127))
        worst = max(worst, float(np.abs(np.asarray(a, dtype=np.int32) - np.asarray(b, dtype=np.int32)).max()))
    report(f"dynamic_preprocess tile counts ({len(sizes)} sizes x 2 budgets)", float(n_bad), 0)
    report("dynamic_preprocess / expand2square pixels", worst, 0)


def check_process_image_branches():
    """mm_utils.process_image / process_images (:442-541) executed from the reference source, with a
    PIL-backed stand-in for the 4.46 SiglipImageProcessor (its pixel step is pinned separately in
    check_pixel_preprocess), vs vila_b200.model.media.process_image(s): every aspect-ratio branch
    (`resize`, `pad`, `dynamic`, `dynamic_s2`, default) must produce the same tensors / block sizes."""
    import os
    import numpy as np
    from PIL import Image
    from vila_b200.model import LlavaConfig, media
    names = ["find_closest_aspect_ratio", "dynamic_preprocess", "dynamic_s2_preprocess", "process_image", "process_images"]
    srcs = extract_functions(REF / "llava/mm_utils.py", names)
    ns = {"Image": Image, "os": os, "torch": torch}
    for k in names:
        exec(srcs[k], ns)

    class Processor:  # SigLIP flavour: `size`, no `crop_size`
        size = {"height": 448, "width": 448}
        image_mean = [0.5, 0.5, 0.5]

        def preprocess(self, image, return_tensors="pt"):
            return {"pixel_values": [media._to_tensor(image, 448)]}

    rng = np.random.RandomState(21)
    sizes = [(448, 448), (640, 480), (333, 1000), (1600, 800), (97, 131)]
    worst, bad = 0.0, 0
    for mode in ("resize", "pad", "dynamic", "dynamic_s2", "default"):
        ref_args = types.SimpleNamespace(image_processor=Processor(), image_aspect_ratio=mode, s2_scales=[448, 896, 1344],
                                         max_tiles=12, min_tiles=1)
        cfg = LlavaConfig(image_aspect_ratio=mode, dynamic_s2=(mode == "dynamic_s2"))
        for (w, h) in sizes:
            img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8))
            if mode == "dynamic_s2":
                ref, ref_bs = ns["process_image"](img, ref_args, None, enable_dynamic_s2=True)
                mine, my_bs = media.process_image(img, cfg, enable_dynamic_s2=True)
                bad += int(tuple(ref_bs) != tuple(my_bs))
            elif mode == "dynamic":
                for mt in (None, 6):
                    ref = ns["process_image"](img, ref_args, None, enable_dynamic_res=True, max_tiles=mt)
                    mine = media.process_image(img, cfg, enable_dynamic_res=True, max_tiles=mt)
                    if ref.shape != mine.shape:
                        bad += 1
                    else:
                        worst = max(worst, float((ref - mine).abs().max()))
                continue
            else:
                ref = ns["process_image"](img, ref_args, None)
                mine = media.process_image(img, cfg)
            if ref.shape != mine.shape:
                bad += 1
            else:
                worst = max(worst, float((ref - mine).abs().max()))
        # process_images: a list of images -> one stacked batch (the multi-image / video-frame path)
        imgs = [Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)) for (w, h) in sizes[:3]]
        if mode in ("resize", "pad", "default"):
            ref_b = ns["process_images"](imgs, Processor(), ref_args)
            mine_b, _ = media.process_images(imgs, cfg)
            if tuple(ref_b.shape) != (len(mine_b),) + tuple(mine_b[0].shape):
                bad += 1
            else:
                worst = max(worst, float((ref_b - torch.stack(mine_b)).abs().max()))
    report("process_image(s): shapes / block sizes over 5 aspect-ratio branches", float(bad), 0)
    report("process_image(s): tensors over 5 aspect-ratio branches", worst, 0)


def toy_chat_tokenizer():
    """A genuine HF fast tokenizer (word-level vocabulary, ChatML-style template like Qwen2's) built in
    memory: no tokenizer files ship with the image, but apply_chat_template / add_tokens / decode are
    the real transformers code paths."""
    from tokenizers import Regex, Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    words = ["<unk>", "<|im_start|>", "<|im_end|>", "<|endoftext|>", "\n", "system", "user", "assistant", "question",
             "answer", "You", "are", "a", "helpful", "Describe", "the", "image", ".", "what", "?", "look", ":", "compare"]
    t = Tokenizer(models.WordLevel({w: i for i, w in enumerate(words)}, unk_token="<unk>"))
    t.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split("\n", "isolated"),
                                               pre_tokenizers.Split(Regex(" +"), "removed")])
    tok = PreTrainedTokenizerFast(tokenizer_object=t, unk_token="<unk>", eos_token="<|endoftext|>",
                                  pad_token="<|endoftext|>", additional_special_tokens=["<|im_start|>", "<|im_end|>"])
    tok.chat_template = ("{% for message in messages %}{{'<|im_start|>' + message['role'] + '\n' + message['content'] + "
                         "'<|im_end|>' + '\n'}}{% endfor %}{% if add_generation_prompt %}{{ '<|im_start|>assistant\n' }}{% endif %}")
    return tok


def ref_tokenizer_namespace():
    """llava/utils/tokenizer.py functions lifted out of the reference source (the module imports
    llava.conversation / llava.mm_utils at the top): tokenize_conversation, infer_stop_tokens."""
    from typing import Any, Dict, List, Optional, Sequence
    import transformers
    names = ["tokenize_conversation", "_maybe_add_sentinel_token", "infer_stop_tokens"]
    srcs = extract_functions(REF / "llava/utils/tokenizer.py", names)
    mm = extract_functions(REF / "llava/mm_utils.py", ["tokenizer_image_token"])
    auto = types.SimpleNamespace(AUTO="auto")
    ns = {"torch": torch, "transformers": transformers, "Any": Any, "Dict": Dict, "List": List, "Optional": Optional,
          "Sequence": Sequence, "SENTINEL_TOKEN": "<vila/sentinel>", "IGNORE_INDEX": -100,
          "conversation_lib": types.SimpleNamespace(SeparatorStyle=auto, default_conversation=types.SimpleNamespace(sep_style="auto")),
          "DUMMY_CONVERSATION": [{"from": "human", "value": "question"}, {"from": "gpt", "value": "answer"}] * 10}
    exec(mm["tokenizer_image_token"], ns)
    for k in names:
        exec(srcs[k], ns)
    return ns


def check_tokenizer_setup():
    """build_llm_and_tokenizer's tokenizer preparation (builder.py:187-211) + utils/tokenizer.py
    (infer_stop_tokens, tokenize_conversation with add_generation_prompt) executed from the reference
    source on a real HF tokenizer, vs vila_b200.model.loading.prepare_tokenizer and
    media.tokenize_conversation: stop tokens, media token ids, prompt ids."""
    import copy
    from vila_b200.model import loading, media
    ns = ref_tokenizer_namespace()
    ref_tok, my_tok = toy_chat_tokenizer(), toy_chat_tokenizer()
    # reference order: stop tokens (adds the sentinel), then the media tokens
    ref_tok.stop_tokens = ns["infer_stop_tokens"](ref_tok)
    ref_tok.stop_token_ids = ref_tok.convert_tokens_to_ids(ref_tok.stop_tokens)
    ref_media = {}
    for name, token in {"image": "<image>", "video": "<vila/video>"}.items():
        ref_tok.add_tokens([token], special_tokens=True)
        ref_media[name] = ref_tok.convert_tokens_to_ids(token)
    loading.prepare_tokenizer(my_tok, model_max_length=4096)
    bad = int(sorted(ref_tok.stop_tokens) != sorted(my_tok.stop_tokens)) + int(ref_media != my_tok.media_token_ids) \
        + int(sorted(ref_tok.stop_token_ids) != sorted(my_tok.stop_token_ids))
    report("tokenizer setup: stop tokens + media token ids", float(bad), 0)
    bad = 0
    for text in ("Describe the image .", "<image>Describe the image .", "look : <image>what ?", "  <image><image>compare  "):
        ref_ids = ns["tokenize_conversation"]([{"from": "human", "value": text}], ref_tok, add_generation_prompt=True)
        mine = media.tokenize_conversation(text, my_tok)
        bad += int(ref_ids.tolist() != list(mine))
    report("tokenize_conversation ids (4 prompts, generation prompt appended)", float(bad), 0)


def extract_media_cases(tmp_dir):
    """Seeded prompts for the prompt-flattening check: (label, spec) with spec a list of
    ("text", str) | ("pil", w, h, seed) | ("image_file", w, h, seed) | ("video_dir", n_frames).
    `build_prompt_parts` turns a spec into real prompt parts for either implementation."""
    return [
        ("plain string", [("text", "What is the capital of France?")]),
        ("text image text", [("text", "look: "), ("pil", 64, 48, 1), ("text", "what?")]),
        ("two images", [("pil", 32, 32, 2), ("pil", 40, 24, 3), ("text", "compare")]),
        ("stray media token in the text", [("text", "<image> a token typed by hand "), ("pil", 32, 32, 4)]),
        ("stray video token, whitespace", [("text", "  before <vila/video> after  "), ("pil", 16, 16, 5), ("text", "\nnew line")]),
        ("image file", [("image_file", 48, 32, 6), ("text", "from a path")]),
        ("video directory (5 frames -> 8)", [("video_dir", 5), ("text", "describe the video")]),
        ("video then image", [("video_dir", 3), ("pil", 20, 20, 7), ("text", "q")]),
    ]


def build_prompt_parts(spec, tmp_dir, ImageCls, VideoCls):
    import numpy as np
    from PIL import Image as PILImage
    from pathlib import Path as _P
    parts = []
    for i, item in enumerate(spec):
        if item[0] == "text":
            parts.append(item[1])
        elif item[0] in ("pil", "image_file"):
            _, w, h, seed = item
            img = PILImage.fromarray(np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8))
            if item[0] == "pil":
                parts.append(img)
            else:
                f = _P(tmp_dir) / f"img_{seed}.png"
                img.save(f)
                parts.append(ImageCls(str(f)))
        else:
            d = _P(tmp_dir) / f"video_{item[1]}_{i}"
            d.mkdir(parents=True, exist_ok=True)
            for k in range(item[1]):
                PILImage.fromarray(np.full((24, 32, 3), 10 * k, dtype=np.uint8)).save(d / f"frame_{k:03d}.png")
            parts.append(VideoCls(str(d)))
    return parts


def ref_extract_media_namespace():
    """llava/utils/media.py functions executed from the reference source (its module imports cv2 /
    requests / llava.* at the top, so the functions are lifted out with ast)."""
    import glob
    import os
    from collections import defaultdict
    from typing import Any, Dict, List, Optional, Union
    import numpy as np
    import PIL.Image
    media_mod = load_by_path("ref_llava_media", REF / "llava/media.py")
    srcs = extract_functions(REF / "llava/utils/media.py", ["_extract_image", "_load_video", "_extract_video", "extract_media"])
    ns = {"glob": glob, "os": os, "np": np, "PIL": PIL, "defaultdict": defaultdict, "Any": Any, "Dict": Dict,
          "List": List, "Optional": Optional, "Union": Union, "PretrainedConfig": object,
          "Image": media_mod.Image, "Video": media_mod.Video,
          "MEDIA_TOKENS": {"image": "<image>", "video": "<vila/video>"},
          "make_list": lambda x: x if isinstance(x, list) else [x],
          "logger": types.SimpleNamespace(warning=lambda *a, **k: None, info=lambda *a, **k: None)}
    for k in ("_extract_image", "_load_video", "_extract_video", "extract_media"):
        exec(srcs[k], ns)
    return ns, media_mod


def check_extract_media():
    """llava/utils/media.py extract_media (+ the .strip() of tokenizer.py:80-82) vs
    vila_b200.model.media.extract_media: flattened text and the image list."""
    import tempfile
    import numpy as np
    from vila_b200.model import media
    ns, media_mod = ref_extract_media_namespace()
    # the constants the namespace hard-codes are the reference's
    const_src = (REF / "llava/constants.py").read_text()
    assert '"image": "<image>"' in const_src and '"video": "<vila/video>"' in const_src
    cfg = types.SimpleNamespace(num_video_frames=8, fps=0.0)
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        for label, spec in extract_media_cases(tmp):
            ref_parts = build_prompt_parts(spec, tmp, media_mod.Image, media_mod.Video)
            my_parts = build_prompt_parts(spec, tmp, media.Image, media.Video)
            conv = [{"from": "human", "value": ref_parts if len(ref_parts) > 1 or not isinstance(ref_parts[0], str) else ref_parts[0]}]
            ref_media = ns["extract_media"](conv, cfg)
            ref_text, ref_imgs = conv[0]["value"], ref_media["image"]
            my_text, my_imgs = media.extract_media(my_parts if len(my_parts) > 1 or not isinstance(my_parts[0], str) else my_parts[0], cfg)
            same = (ref_text == my_text and len(ref_imgs) == len(my_imgs) and all(
                np.array_equal(np.asarray(a.convert("RGB")), np.asarray(b.convert("RGB"))) for a, b in zip(ref_imgs, my_imgs)))
            if not same:
                print("   mismatch:", label, repr(ref_text), repr(my_text), len(ref_imgs), len(my_imgs))
                bad += 1
    report("extract_media text + image list (8 prompt shapes)", float(bad), 0)


def check_server_video_sampling():
    """serving/server.py sample_frames_from_video (int(total / n * i), its own rule) executed from the
    reference source vs vila_b200.server on a synthetic 23-frame mp4 sent as a base64 data URL."""
    import base64
    import os
    import tempfile
    import cv2
    import numpy as np
    from PIL import Image as PILImage
    from vila_b200 import server
    srcs = extract_functions(REF / "serving/server.py", ["sample_frames_from_video"])
    ns = {"cv2": cv2, "PILImage": PILImage}
    exec(srcs["sample_frames_from_video"], ns)
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "clip.mp4")
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 10, (64, 48))
        for k in range(23):
            wr.write(np.full((48, 64, 3), k * 10, dtype=np.uint8))
        wr.release()
        url = "data:video/mp4;base64," + base64.b64encode(open(path, "rb").read()).decode()
        for n in (8, 5, 16):
            ref = ns["sample_frames_from_video"](path, n)
            mine = server.sample_frames_from_video(server.load_video(url), n)
            if len(ref) != len(mine) or not all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(ref, mine)):
                bad += 1
    report("server video_url frame sampling (3 frame counts)", float(bad), 0)


if __name__ == "__main__":
    if not REF.exists():
        print("reference tree not present: nothing to validate against")
        sys.exit(0)
    torch.set_num_threads(8)
    check_siglip()
    check_projector()
    check_arch_glue()
    check_media_preprocess()
    check_pixel_preprocess()
    check_encoders()
    check_packing()
    check_dynamic_preprocess()
    check_process_image_branches()
    check_extract_media()
    check_tokenizer_setup()
    check_server_video_sampling()
    check_qwen2()
    print("FAILED:" if FAILED else "ALL OK", FAILED)
    sys.exit(1 if FAILED else 0)
