"""Generate tests/golden/*.pt from the REFERENCE's own modules (run once in the authoring container,
where /root/reference exists; the fixtures are committed so tests never need the reference tree).

Each fixture holds seeded weights (state dict with the reference's names), the synthetic input and the
output produced by the reference implementation:
  siglip_tiny.pt     vendored SiglipVisionModel (modeling_siglip.py), hidden_states[-2], sdpa + eager
  projector_*.pt     reference MultimodalProjector for the three NVILA projector types
  arch_glue.pt       reference llava_arch.py excerpts (dynamic-S2 encode_images, _embed splice), run
                     unmodified via ast extraction
  media_preprocess.pt  reference mm_utils.dynamic_s2_preprocess on seeded images: block sizes + per-tile
                     pixel sums and a 16x16 thumbnail of every tile
  qwen2_tiny.pt      transformers Qwen2ForCausalLM logits + greedy ids (the reference's LLM is this
                     third-party class; pinned transformers==4.46.0, here the installed version)
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import validate_against_reference as V  # noqa: E402
from oracle import vila_oracle as O  # noqa: E402

OUT = ROOT / "tests" / "golden"
REF = V.REF


def half(sd):
    return {k: v.detach().clone() for k, v in sd.items()}


def gen_siglip():
    import transformers.models.siglip.configuration_siglip as cfgmod
    ms = V.load_by_path("ref_modeling_siglip", REF / "llava/model/multimodal_encoder/siglip/modeling_siglip.py")
    torch.manual_seed(100)
    kw = dict(hidden_size=72, intermediate_size=136, num_hidden_layers=3, num_attention_heads=1,
              image_size=42, patch_size=14)
    cfg = cfgmod.SiglipVisionConfig(**kw)
    cfg._attn_implementation = "sdpa"
    model = ms.SiglipVisionModel(cfg).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
                if "layer_norm" in n and n.endswith("weight"):
                    p.add_(1.0)
    px = torch.randn(2, 3, 42, 42)
    with torch.no_grad():
        hs = model(px, output_hidden_states=True).hidden_states
    sd = {k: v for k, v in model.state_dict().items() if ".head." not in k}
    torch.save({"cfg": kw, "weights": half(sd), "pixels": px, "hidden_m2": hs[-2], "hidden_m1": hs[-1],
                "n_hidden_states": len(hs)}, OUT / "siglip_tiny.pt")


def gen_projector():
    import transformers  # noqa: F401
    timm = types.ModuleType("timm"); timm_m = types.ModuleType("timm.models")
    timm_l = types.ModuleType("timm.models.layers"); timm_l.Mlp = type("Mlp", (torch.nn.Module,), {})
    sys.modules.setdefault("timm", timm); sys.modules.setdefault("timm.models", timm_m)
    sys.modules.setdefault("timm.models.layers", timm_l)
    bp = V.load_by_path("refprojector", REF / "llava/model/multimodal_projector/base_projector.py")
    torch.manual_seed(101)
    out = {}
    for kind, n_tok in (("mlp_downsample", 49), ("mlp_downsample_2x2_fix", 64), ("mlp_downsample_3x3_fix", 64)):
        cfg = types.SimpleNamespace(mm_hidden_size=24, hidden_size=40)
        model = bp.MultimodalProjector(bp.MultimodalProjectorConfig(kind), cfg).eval()
        with torch.no_grad():
            for p in model.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.2)
        x = torch.randn(2, n_tok, 24)
        with torch.no_grad():
            y = model(x)
        out[kind] = {"weights": half(model.state_dict()), "x": x, "y": y}
    torch.save(out, OUT / "projector.pt")


def gen_arch_glue():
    from collections import defaultdict, deque
    from einops import rearrange
    import textwrap
    names = ["merge_chessboard", "split_chessboard", "merge_features_for_dynamic_s2", "encode_images",
             "_embed", "__embed_media_tokens", "__truncate_sequence", "__batchify_sequence"]
    srcs = V.extract_functions(REF / "llava/model/llava_arch.py", names)
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
    torch.manual_seed(102)
    C_, side = 8, 4
    lin_w, lin_b = torch.randn(16, 3 * 4 * C_) * 0.1, torch.randn(16) * 0.1
    fixtures = {"lin_w": lin_w, "lin_b": lin_b, "s2": []}
    for idx in (-1, 0, 1):
        tower = types.SimpleNamespace(scales=[4, 8, 12], resize_output_to_scale_idx=idx)
        tower_fn = lambda images: images
        proj = lambda f: F.linear(O.downsample(f, 2), lin_w, lin_b)
        self_ = RefArch()
        self_.config = types.SimpleNamespace(dynamic_s2=True)
        tower_obj = type("T", (), {"scales": [4, 8, 12], "resize_output_to_scale_idx": idx,
                                   "__call__": lambda s, im: im})()
        self_.get_vision_tower = lambda t=tower_obj: t
        self_.get_mm_projector = lambda: proj
        block_sizes = [(2, 3), None, (1, 2)]
        n_tiles = (1 + 4 + 6) + 1 + (1 + 4 + 2)
        feats = torch.randn(n_tiles, side * side, C_)
        with torch.no_grad():
            ref = self_.encode_images(feats, block_sizes)
        fixtures["s2"].append({"idx": idx, "block_sizes": block_sizes, "feats": feats,
                               "out": [r.clone() for r in ref]})
    # _embed
    V_, H = 50, 8
    table = torch.randn(V_, H)
    self_ = RefArch()
    self_.training = False
    self_.llm = types.SimpleNamespace(model=types.SimpleNamespace(embed_tokens=lambda ids: F.embedding(ids, table)))
    IMG, VID = 40, 41
    self_.tokenizer = types.SimpleNamespace(media_token_ids={"image": IMG, "video": VID},
                                            padding_side="right", model_max_length=4096)
    m_img = [torch.randn(5, H), torch.randn(3, H)]
    m_vid = [torch.randn(7, H)]
    self_.encoders = {"image": lambda media, cfg: list(media), "video": lambda media, cfg: list(media)}
    ids = torch.tensor([[1, 2, IMG, 3, VID, 4, 0, 0], [IMG, 9, 8, 7, 6, 5, 4, 3]])
    am = torch.tensor([[1, 1, 1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1, 1, 1]], dtype=torch.bool)
    emb = {}
    for side_ in ("right", "left"):
        self_.tokenizer.padding_side = side_
        a, b, c = getattr(self_, "_embed")(ids, {"image": list(m_img), "video": list(m_vid)},
                                           {"image": {}, "video": {}}, None, am)
        emb[side_] = {"inputs": a, "labels": b, "mask": c}
    fixtures["embed"] = {"table": table, "ids": ids, "mask": am, "m_img": m_img, "m_vid": m_vid,
                         "IMG": IMG, "VID": VID, "out": emb}
    torch.save(fixtures, OUT / "arch_glue.pt")


def gen_qwen2():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(103)
    kw = dict(hidden_size=64, intermediate_size=160, num_hidden_layers=2, num_attention_heads=4,
              num_key_value_heads=2, vocab_size=256, rms_norm_eps=1e-6, rope_theta=1000000.0)
    hcfg = Qwen2Config(max_position_embeddings=4096, tie_word_embeddings=False,
                       attn_implementation="sdpa", **kw)
    model = Qwen2ForCausalLM(hcfg).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.normal_(1.0, 0.1)
            elif p.dim() == 1:
                p.normal_(0, 0.1)
    emb = torch.randn(1, 21, 64)
    pos = torch.arange(30000, 30021)
    with torch.no_grad():
        logits = model(inputs_embeds=emb).logits[0]
        logits_far = model(inputs_embeds=emb, position_ids=pos[None]).logits[0]
        gen = model.generate(inputs_embeds=emb, attention_mask=torch.ones(1, 21, dtype=torch.long),
                             max_new_tokens=10, do_sample=False, eos_token_id=None, pad_token_id=0)
    torch.save({"cfg": kw, "weights": half(model.state_dict()), "emb": emb, "logits": logits,
                "pos_far": pos, "logits_far": logits_far, "greedy": gen[0].tolist(),
                "transformers_version": __import__("transformers").__version__},
               OUT / "qwen2_tiny.pt")


MEDIA_SIZES = [(448, 448), (1600, 800), (800, 1600), (333, 1000), (1920, 1080), (640, 480), (97, 131), (3000, 500)]


def media_test_image(w: int, h: int, seed: int):
    """Seeded synthetic RGB image (numpy RandomState: identical in the test)."""
    import numpy as np
    from PIL import Image
    return Image.fromarray(np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8))


def gen_media():
    import numpy as np
    srcs = V.extract_functions(REF / "llava/mm_utils.py", ["find_closest_aspect_ratio", "dynamic_s2_preprocess"])
    ns = {}
    exec(srcs["find_closest_aspect_ratio"], ns)
    exec(srcs["dynamic_s2_preprocess"], ns)
    out = []
    for i, (w, h) in enumerate(MEDIA_SIZES):
        tiles, bs = ns["dynamic_s2_preprocess"](media_test_image(w, h, 100 + i), s2_scales=[448, 896, 1344],
                                                max_num=12, image_size=448)
        arrs = [np.asarray(t, dtype=np.int64) for t in tiles]
        out.append({"size": (w, h), "seed": 100 + i, "block_size": tuple(bs), "n_tiles": len(tiles),
                    "tile_sums": torch.tensor([int(a.sum()) for a in arrs]),
                    "tile_thumbs": torch.tensor(np.stack([a[::28, ::28, :] for a in arrs]), dtype=torch.uint8)})
    torch.save(out, OUT / "media_preprocess.pt")


def gen_packing():
    """reference repack_multimodal_data / _get_unpad_data outputs on the seeded cases of
    validate_against_reference.packing_cases()."""
    import types
    ref_repack, ref_unpad = V.ref_packing_namespace()
    out = []
    for emb, mask, labels, pad_mult in V.packing_cases():
        holder = types.SimpleNamespace(llm=types.SimpleNamespace(pad_token_id=0))
        if pad_mult:
            holder.pad_to_multiple_of = pad_mult
        e, am, pos, lab = ref_repack(holder, emb, mask, None, labels.clone())
        idx, cu, mx = ref_unpad(am)
        out.append({"emb": emb, "mask": mask, "labels": labels, "pad_mult": pad_mult,
                    "out": (e, am, pos, lab), "unpad": (idx, cu, mx)})
    torch.save(out, OUT / "packing.pt")


def gen_dynamic_preprocess():
    import numpy as np
    from PIL import Image
    srcs = V.extract_functions(REF / "llava/mm_utils.py", ["find_closest_aspect_ratio", "dynamic_preprocess", "expand2square"])
    ns = {"Image": Image}
    for k in ("find_closest_aspect_ratio", "dynamic_preprocess", "expand2square"):
        exec(srcs[k], ns)
    out = []
    for i, (w, h) in enumerate(MEDIA_SIZES + [(336, 336)]):
        img = media_test_image(w, h, 200 + i)
        tiles = ns["dynamic_preprocess"](img, min_num=1, max_num=12, image_size=448)
        arrs = [np.asarray(t, dtype=np.int64) for t in tiles]
        sq = np.asarray(ns["expand2square"](img, (127, 127, 127)), dtype=np.int64)
        out.append({"size": (w, h), "seed": 200 + i, "n_tiles": len(tiles),
                    "tile_sums": torch.tensor([int(a.sum()) for a in arrs]),
                    "square_shape": tuple(sq.shape), "square_sum": int(sq.sum())})
    torch.save(out, OUT / "dynamic_preprocess.pt")


def gen_extract_media():
    """reference extract_media (llava/utils/media.py) on the seeded prompt shapes of
    validate_against_reference.extract_media_cases() -> tests/golden/extract_media.json"""
    import json
    import tempfile
    ns, media_mod = V.ref_extract_media_namespace()
    cfg = types.SimpleNamespace(num_video_frames=8, fps=0.0)
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for label, spec in V.extract_media_cases(tmp):
            parts = V.build_prompt_parts(spec, tmp, media_mod.Image, media_mod.Video)
            conv = [{"from": "human", "value": parts if len(parts) > 1 or not isinstance(parts[0], str) else parts[0]}]
            media = ns["extract_media"](conv, cfg)
            out.append({"label": label, "spec": spec, "text": conv[0]["value"], "stripped": conv[0]["value"].strip(),
                        "image_sizes": [list(im.size) for im in media["image"]]})
    (OUT / "extract_media.json").write_text(json.dumps(out, indent=1))


def gen_tokenizer_setup():
    """reference infer_stop_tokens / media-token registration / tokenize_conversation on the in-memory
    HF tokenizer of validate_against_reference.toy_chat_tokenizer() -> tests/golden/tokenizer_setup.json"""
    import json
    ns = V.ref_tokenizer_namespace()
    tok = V.toy_chat_tokenizer()
    stop = ns["infer_stop_tokens"](tok)
    media = {}
    for name, token in {"image": "<image>", "video": "<vila/video>"}.items():
        tok.add_tokens([token], special_tokens=True)
        media[name] = tok.convert_tokens_to_ids(token)
    prompts = ["Describe the image .", "<image>Describe the image .", "look : <image>what ?", "  <image><image>compare  "]
    ids = [ns["tokenize_conversation"]([{"from": "human", "value": t}], tok, add_generation_prompt=True).tolist() for t in prompts]
    import copy
    convs = [[{"from": "human", "value": " <image>Describe the image . "}, {"from": "gpt", "value": "answer"}],
             [{"from": "human", "value": "question"}, {"from": "gpt", "value": "answer"}, {"from": "human", "value": "what ?"}]]
    kwargs = [{}, {"add_generation_prompt": True}, {"overrides": {"gpt": "<vila/sentinel>"}}, {"no_system_prompt": True}]
    multi = [{"conversation": c, "kwargs": kw,
              "input_ids": ns["tokenize_conversation"](copy.deepcopy(c), tok, **kw).tolist()} for c in convs for kw in kwargs]
    (OUT / "tokenizer_setup.json").write_text(json.dumps({"conversations": multi,
        "stop_tokens": sorted(stop), "stop_token_ids": sorted(tok.convert_tokens_to_ids(stop)), "media_token_ids": media,
        "sentinel_token_id": tok.sentinel_token_id, "prompts": prompts, "input_ids": ids}, indent=1))


def gen_api_signatures():
    """Public-method signatures of the reference's model classes, extracted from the source with ast
    (names, order, literal defaults, *args / **kwargs) -> tests/golden/api_signatures.json."""
    import ast
    import json

    def sigs(path, cls_names, methods):
        src = (REF / path).read_text()
        tree = ast.parse(src)
        out = {}
        for node in ast.walk(tree):
            if isinstance(node, ast.ClassDef) and node.name in cls_names:
                for fn in node.body:
                    if isinstance(fn, ast.FunctionDef) and fn.name in methods:
                        a = fn.args
                        names = [x.arg for x in a.args]
                        defaults = [None] * (len(names) - len(a.defaults)) + [
                            ast.literal_eval(d) if isinstance(d, ast.Constant) else "<expr>" for d in a.defaults]
                        out.setdefault(node.name, {})[fn.name] = {
                            "args": names, "defaults": ["<required>" if i < len(names) - len(a.defaults) else defaults[i]
                                                        for i in range(len(names))],
                            "vararg": a.vararg.arg if a.vararg else None,
                            "kwarg": a.kwarg.arg if a.kwarg else None,
                            "kwonly": [x.arg for x in a.kwonlyargs]}
        return out

    methods = {"forward", "generate", "generate_content", "encode_images", "_embed", "get_llm", "get_lm_head",
               "get_vision_tower", "get_mm_projector", "repack_multimodal_data", "get_xgr_logits_processor",
               "merge_features_for_dynamic_s2"}
    table = {
        "llava/model/language_model/llava_llama.py": sigs("llava/model/language_model/llava_llama.py", {"LlavaLlamaModel"}, methods),
        "llava/model/llava_arch.py": sigs("llava/model/llava_arch.py", {"LlavaMetaModel", "LlavaMetaForCausalLM"}, methods),
        "llava/remote_code/modeling_vila.py": sigs("llava/remote_code/modeling_vila.py", {"VILAForCausalLM", "VILAPretrainedModel"}, methods),
        "llava/model/multimodal_projector/base_projector.py": sigs("llava/model/multimodal_projector/base_projector.py", {"MultimodalProjector"}, {"forward"}),
        "llava/model/encoders/image/basic.py": sigs("llava/model/encoders/image/basic.py", {"BasicImageEncoder"}, {"forward", "embed_tokens", "_process_features"}),
        "llava/model/encoders/video/basic.py": sigs("llava/model/encoders/video/basic.py", {"BasicVideoEncoder"}, {"forward", "_process_features"}),
        "llava/model/encoders/video/tsp.py": sigs("llava/model/encoders/video/tsp.py", {"TSPVideoEncoder"}, {"forward", "_process_features"}),
    }
    (OUT / "api_signatures.json").write_text(json.dumps(table, indent=1, sort_keys=True))


if __name__ == "__main__":
    assert REF.exists(), "needs /root/reference"
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(4)
    if "--new-only" not in sys.argv:
        gen_siglip(); gen_projector(); gen_arch_glue(); gen_qwen2(); gen_media()
    gen_packing(); gen_dynamic_preprocess(); gen_extract_media(); gen_tokenizer_setup(); gen_api_signatures()
    for f in sorted(OUT.glob("*.pt")):
        print(f.name, f.stat().st_size)
