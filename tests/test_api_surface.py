"""CPU: the drop-in surface — signatures of every public method against the
signatures extracted from the reference source (tests/golden/api_signatures.json, made by
oracle/gen_golden.py with `ast`), the `llava` namespace shim, the CLI plumbing of BASELINE configs[0]
(`llava.cli.infer` flow up to the model call), packing / dynamic tiling against reference-generated
fixtures, and the remote-code `generate` semantics."""
import inspect
import json
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"

# reference class -> our class
def _classes():
    from vila_b200 import model as M
    return {"LlavaLlamaModel": M.LlavaLlamaModel, "LlavaMetaModel": M.LlavaLlamaModel,
            "LlavaMetaForCausalLM": M.LlavaLlamaModel, "VILAForCausalLM": M.VILAForCausalLM,
            "VILAPretrainedModel": M.VILAForCausalLM, "MultimodalProjector": M.MultimodalProjector,
            "BasicImageEncoder": M.BasicImageEncoder, "BasicVideoEncoder": M.BasicVideoEncoder,
            "TSPVideoEncoder": M.TSPVideoEncoder}


def test_public_signatures_accept_every_reference_call():
    """For every method the reference defines on the path: same name; every reference parameter is
    either a parameter of ours at the same position with the same literal default, or (for the PS3 /
    top-down arguments that are out of scope) absorbed by our **kwargs; ours may only ADD defaulted
    parameters at the end."""
    table = json.loads((GOLDEN / "api_signatures.json").read_text())
    classes = _classes()
    checked = 0
    for fname, per_class in table.items():
        for cname, methods in per_class.items():
            ours_cls = classes[cname]
            for mname, ref in methods.items():
                fn = getattr(ours_cls, mname, None)
                assert fn is not None, f"{cname}.{mname} ({fname}) is missing"
                if isinstance(inspect.getattr_static(ours_cls, mname), property):
                    continue
                sig = inspect.signature(fn)
                params = list(sig.parameters.values())
                names = [p.name for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
                has_kw = any(p.kind == p.VAR_KEYWORD for p in params)
                for i, (arg, default) in enumerate(zip(ref["args"], ref["defaults"])):
                    if i < len(names) and names[i] == arg:
                        ours_default = sig.parameters[arg].default
                        if default == "<required>":
                            continue
                        if default != "<expr>":
                            assert ours_default == default or ours_default is inspect.Parameter.empty and arg == "self", \
                                f"{cname}.{mname}({arg}): default {ours_default!r} != reference {default!r}"
                    else:
                        assert has_kw and default != "<required>", \
                            f"{cname}.{mname}: reference parameter #{i} `{arg}` not accepted (ours: {names})"
                        break  # the remaining reference parameters go to **kwargs as well
                if ref["kwarg"]:
                    assert has_kw, f"{cname}.{mname} must accept **{ref['kwarg']}"
                for p in params[len(ref["args"]):]:
                    if p.kind == p.POSITIONAL_OR_KEYWORD:
                        assert p.default is not inspect.Parameter.empty, f"{cname}.{mname}: extra required `{p.name}`"
                checked += 1
    assert checked >= 30


def test_llava_namespace_shim():
    import llava
    import llava.cli.infer
    import llava.conversation as clib
    import llava.mm_utils
    import llava.model
    import llava.model.builder
    import llava.model.configuration_llava as cl
    import llava.remote_code.modeling_vila as rv
    from vila_b200 import model as M
    assert llava.model.LlavaLlamaModel is M.LlavaLlamaModel and rv.VILAForCausalLM is M.VILAForCausalLM
    assert list(inspect.signature(llava.load).parameters) == ["model_path", "model_base", "devices", "kwargs"]
    assert list(inspect.signature(llava.model.builder.load_pretrained_model).parameters)[:7] == [
        "model_path", "model_name", "model_base", "load_8bit", "load_4bit", "device_map", "device"]
    assert llava.Image("a.png").path == "a.png" and issubclass(llava.Video, llava.Media)
    assert llava.mm_utils.get_model_name_from_path("x/run/checkpoint-7/") == "run_checkpoint-7"
    rf = cl.ResponseFormat(type="json_schema", json_schema=cl.JsonSchemaResponseFormat(schema="{}"))
    assert rf.json_schema.schema_ == "{}"
    assert "auto" in clib.conv_templates and clib.default_conversation.copy().name == "auto"


def test_cli_infer_plumbing(monkeypatch, tmp_path):
    """BASELINE configs[0] plumbing: `llava.cli.infer` with an image + text builds the prompt, maps the
    flags and calls model.generate_content exactly once (the model itself needs the GPU)."""
    import llava
    import llava.cli.infer as infer
    from PIL import Image as PILImage
    img = tmp_path / "synthetic.png"
    PILImage.fromarray(np.random.RandomState(0).randint(0, 256, (336, 336, 3), dtype=np.uint8)).save(img)
    calls = {}

    class Stub:
        config = SimpleNamespace(num_video_frames=8, video_max_tiles=1)

        def generate_content(self, prompt, response_format=None):
            calls["prompt"], calls["rf"] = prompt, response_format
            return "ok"

    monkeypatch.setattr(llava, "load", lambda path, model_base=None: Stub())
    out = infer.main(["--model-path", "NVILA-Lite-3B", "--media", str(img), "--text", "Describe.",
                      "--num_video_frames", "16", "--json-mode"])
    assert out == "ok" and calls["rf"].type == "json_object"
    assert isinstance(calls["prompt"][0], llava.Image) and calls["prompt"][1] == "Describe."
    assert Stub.config.num_video_frames == 16
    with pytest.raises(ValueError):
        infer.main(["--model-path", "x", "--media", "file.xyz"])


def test_prepare_content_dynamic_tiles_on_host():
    """NVILA-Lite (`image_aspect_ratio == "dynamic"`): one PIL image -> tiles + thumbnail, the image
    token repeated once per tile, ids tokenised — all before any GPU work (configs[0]: 336^2 input)."""
    from PIL import Image as PILImage
    from vila_b200.model import LlavaLlamaModel, SyntheticTokenizer, nvila_lite_3b
    cfg = nvila_lite_3b()
    m = object.__new__(LlavaLlamaModel)
    torch.nn.Module.__init__(m)
    m.config, m.tokenizer = cfg, SyntheticTokenizer(cfg)
    m.preprocess_on_device = False  # the host (PIL) path; the kernel path is compared with it under -m gpu
    img = PILImage.fromarray(np.random.RandomState(1).randint(0, 256, (336, 336, 3), dtype=np.uint8))
    ids, media, media_config = m._prepare_content([img, "What is this?"])
    assert len(media["image"]) == 1 and media["image"][0].shape == (3, 448, 448)   # 1x1 grid: no thumbnail
    assert int((ids == cfg.image_token_id).sum()) == 1
    wide = PILImage.fromarray(np.random.RandomState(2).randint(0, 256, (400, 1200, 3), dtype=np.uint8))
    ids, media, _ = m._prepare_content([wide, "And this?"])
    assert len(media["image"]) == 4 and int((ids == cfg.image_token_id).sum()) == 4  # 3x1 grid + thumbnail
    assert all(t.dtype == torch.float32 and float(t.abs().max()) <= 1.0 for t in media["image"])


def test_packing_matches_reference_fixture():
    from vila_b200.model import packing
    cases = torch.load(GOLDEN / "packing.pt")
    assert len(cases) == 4
    for c in cases:
        got = packing.repack_multimodal_data(c["emb"], c["mask"], None, c["labels"].clone(), c["pad_mult"], 0)
        for a, b in zip(c["out"], got):
            assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b)
        idx, cu, mx = packing.get_unpad_data(got[1])
        assert torch.equal(idx, c["unpad"][0]) and torch.equal(cu, c["unpad"][1]) and mx == c["unpad"][2]


def test_dynamic_preprocess_and_pad_match_reference_fixture():
    from PIL import Image as PILImage
    from vila_b200.model import media
    cases = torch.load(GOLDEN / "dynamic_preprocess.pt")
    for c in cases:
        w, h = c["size"]
        img = PILImage.fromarray(np.random.RandomState(c["seed"]).randint(0, 256, (h, w, 3), dtype=np.uint8))
        tiles = media.dynamic_preprocess(img, min_num=1, max_num=12, image_size=448)
        assert len(tiles) == c["n_tiles"]
        sums = torch.tensor([int(np.asarray(t, dtype=np.int64).sum()) for t in tiles])
        assert torch.equal(sums, c["tile_sums"])
        sq = np.asarray(media.expand2square(img, (127, 127, 127)), dtype=np.int64)
        assert tuple(sq.shape) == c["square_shape"] and int(sq.sum()) == c["square_sum"]


def test_remote_code_generate_prepends_prompt_ids(monkeypatch):
    """modeling_vila.py:1112-1125."""
    from vila_b200.model import LlavaLlamaModel, VILAForCausalLM
    new = torch.tensor([[7, 8, 9]])
    monkeypatch.setattr(LlavaLlamaModel, "generate", lambda self, **kw: new)
    m = object.__new__(VILAForCausalLM)
    ids = torch.tensor([[1, 2, 3, 4]])
    assert VILAForCausalLM.generate.__wrapped__(m, input_ids=ids).tolist() == [[1, 2, 3, 4, 7, 8, 9]] \
        if hasattr(VILAForCausalLM.generate, "__wrapped__") else True
    out = VILAForCausalLM.generate(m, input_ids=ids)
    assert out.tolist() == [[1, 2, 3, 4, 7, 8, 9]]
    assert VILAForCausalLM.generate(m, input_ids=ids, return_output_ids_only=True).tolist() == [[7, 8, 9]]
    gc = SimpleNamespace(num_return_sequences=2)
    monkeypatch.setattr(LlavaLlamaModel, "generate", lambda self, **kw: torch.tensor([[7], [8]]))
    assert VILAForCausalLM.generate(m, input_ids=ids, generation_config=gc).tolist() == [[1, 2, 3, 4, 7], [1, 2, 3, 4, 8]]


def test_chat_completions_server_plumbing():
    """f3: the OpenAI-style endpoint (reference serving/server.py:209-300) on a stub model — request
    schema, prompt assembly from base64 image parts, streaming SSE chunks, and grouping of concurrent
    non-streaming requests into one generate_batch call."""
    import asyncio
    import base64
    import io
    from PIL import Image as PILImage
    from vila_b200 import server as S
    buf = io.BytesIO()
    PILImage.fromarray(np.random.RandomState(0).randint(0, 256, (32, 48, 3), dtype=np.uint8)).save(buf, format="PNG")
    url = "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode()
    calls = {"batch": [], "single": 0}

    class Tok:
        def decode(self, ids, skip_special_tokens=True):
            return " ".join(str(i) for i in ids)

    class Stub:
        tokenizer = Tok()
        default_generation_config = SimpleNamespace(max_new_tokens=None)

        def generate_content(self, prompt, generation_config=None, response_format=None, stream=False):
            assert any(hasattr(p, "size") for p in prompt)  # the PIL image made it into the prompt
            if stream:
                return iter(["he", "", "llo"])
            calls["single"] += 1
            return "single"

        def _prepare_content(self, prompt):
            return torch.tensor([[1, 2]]), {"image": [torch.zeros(3, 4, 4)]}, {}

        def generate_batch(self, requests, max_new_tokens=128, slots=8):
            calls["batch"].append(len(requests))
            return [[7, 8, 9, 10][:max_new_tokens] for _ in requests]

    def req(stream=False, max_tokens=3):
        return S.ChatCompletionRequest(model="m", stream=stream, max_tokens=max_tokens, messages=[S.ChatMessage(
            role="user", content=[S.TextContent(type="text", text="describe"),
                                  S.ImageContent(type="image_url", image_url=S.MediaURL(url=url))])])

    async def scenario():
        eng = S.Engine(Stub(), "m", slots=4)
        one = await eng.complete(req())
        three = await asyncio.gather(*[eng.complete(req()) for _ in range(3)])
        chunks = [c async for c in eng.stream(req(stream=True))]
        return one, three, chunks

    one, three, chunks = asyncio.run(scenario())
    assert one["object"] == "chat.completion" and one["choices"][0]["message"]["content"] == "single"
    # the first of three concurrent requests finds the engine idle and runs alone; the two that pile up
    # behind the lock are decoded together by ONE generate_batch call
    assert [t["choices"][0]["message"]["content"] for t in three] == ["single", "7 8 9", "7 8 9"]
    assert calls["batch"] == [2] and calls["single"] == 2
    assert chunks[-1] == "data: [DONE]\n\n" and len(chunks) == 3
    assert json.loads(chunks[0][6:])["choices"][0]["delta"]["content"] == "he"
    with pytest.raises(ValueError):
        asyncio.run(S.Engine(Stub(), "other").complete(req()))

    # a batch that fails answers EVERY request it took with the error (nobody is left waiting on a future)
    class Broken(Stub):
        def generate_content(self, prompt, generation_config=None, response_format=None, stream=False):
            raise RuntimeError("boom-single")

        def generate_batch(self, requests, max_new_tokens=128, slots=8):
            raise RuntimeError("boom-batch")

    async def failing():
        eng = S.Engine(Broken(), "m", slots=4)
        return await asyncio.wait_for(asyncio.gather(*[eng.complete(req()) for _ in range(3)], return_exceptions=True), 10)

    errs = asyncio.run(failing())
    assert [type(e) for e in errs] == [RuntimeError] * 3
    assert [str(e) for e in errs] == ["boom-single", "boom-batch", "boom-batch"]


def test_extract_media_matches_reference_fixture(tmp_path):
    """Prompt flattening (llava/utils/media.py:93-123 + the strip of utils/tokenizer.py:80-82): the text
    the tokenizer sees and the image list, against outputs of the reference's own function
    (tests/golden/extract_media.json, oracle/gen_golden.py)."""
    import json
    from types import SimpleNamespace
    from oracle.validate_against_reference import build_prompt_parts
    from vila_b200.model import media
    cases = json.loads((GOLDEN / "extract_media.json").read_text())
    assert len(cases) >= 8
    cfg = SimpleNamespace(num_video_frames=8, fps=0.0)
    for c in cases:
        spec = [tuple(x) for x in c["spec"]]
        parts = build_prompt_parts(spec, tmp_path, media.Image, media.Video)
        prompt = parts if len(parts) > 1 or not isinstance(parts[0], str) else parts[0]
        text, images = media.extract_media(prompt, cfg)
        assert text == c["text"], c["label"]
        assert [list(im.size) for im in images] == c["image_sizes"], c["label"]

        class Tok:  # records what reaches the tokenizer
            def __call__(self, t):
                self.seen = t
                return SimpleNamespace(input_ids=[1])
        tok = Tok()
        media.tokenize_conversation(text, tok)
        assert tok.seen == c["stripped"], c["label"]
    with pytest.raises(ValueError):
        media.extract_media(["x", 3.14], cfg)


def test_server_video_url_sampling(tmp_path):
    """`video_url` parts (serving/server.py:106-143,246-252): a base64 mp4 becomes `frames` PIL images taken at
    index int(total / frames * i) — the server's rule, not extract_media's linspace (validated against the
    reference function in oracle/validate_against_reference.py)."""
    import base64
    cv2 = pytest.importorskip("cv2")
    from vila_b200 import server
    path = str(tmp_path / "clip.mp4")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 10, (64, 48))
    for k in range(23):
        wr.write(np.full((48, 64, 3), k * 10, dtype=np.uint8))
    wr.release()
    url = "data:video/mp4;base64," + base64.b64encode(open(path, "rb").read()).decode()
    msg = server.ChatMessage(role="user", content=[
        {"type": "video_url", "video_url": {"url": url}, "frames": 8}, {"type": "text", "text": "what happens?"}])
    prompt = server.build_prompt([msg])
    assert len(prompt) == 9 and prompt[-1] == "what happens?"
    got = [int(np.asarray(f)[0, 0, 0]) for f in prompt[:8]]
    want = [10 * int(23 / 8 * i) for i in range(8)]
    assert all(abs(g - w) <= 8 for g, w in zip(got, want)), (got, want)   # mp4v is lossy by a few levels
    with pytest.raises(ValueError):
        server.load_video("data:video/avi;base64,AAAA")


def test_tokenizer_setup_matches_reference_fixture():
    """What the reference does to the HF tokenizer at load (builder.py:187-211: stop tokens inferred from the
    chat template, media tokens registered) and `tokenize_conversation(..., add_generation_prompt=True)`,
    on a real (in-memory) HF tokenizer — against values produced by the reference's own functions."""
    pytest.importorskip("tokenizers")
    from oracle.validate_against_reference import toy_chat_tokenizer
    from vila_b200.model import loading, media
    want = json.loads((GOLDEN / "tokenizer_setup.json").read_text())
    tok = loading.prepare_tokenizer(toy_chat_tokenizer(), model_max_length=4096)
    assert tok.padding_side == "right" and tok.model_max_length == 4096
    assert sorted(tok.stop_tokens) == want["stop_tokens"] and sorted(tok.stop_token_ids) == want["stop_token_ids"]
    assert tok.media_token_ids == want["media_token_ids"] and tok.sentinel_token_id == want["sentinel_token_id"]
    for text, ids in zip(want["prompts"], want["input_ids"]):
        assert list(media.tokenize_conversation(text, tok)) == ids, text
    # idempotent: a tokenizer that already carries the tokens (a saved NVILA checkpoint) keeps its ids
    again = loading.prepare_tokenizer(tok)
    assert again.media_token_ids == want["media_token_ids"]


def test_llava_utils_shim_matches_reference_fixture(tmp_path):
    """`llava.constants`, `llava.utils.media.extract_media(messages, config, draft)` and
    `llava.utils.tokenizer.tokenize_conversation(messages, tokenizer, add_generation_prompt, overrides,
    no_system_prompt)` in the reference's call forms, against reference-generated fixtures."""
    import copy
    pytest.importorskip("tokenizers")
    import llava
    from llava.constants import DEFAULT_IMAGE_TOKEN, IGNORE_INDEX, MEDIA_TOKENS, SENTINEL_TOKEN
    from llava.utils import make_list
    from llava.utils.media import extract_media
    from llava.utils.tokenizer import infer_stop_tokens, tokenize_conversation
    from oracle.validate_against_reference import build_prompt_parts, toy_chat_tokenizer
    from vila_b200.model import loading
    assert (IGNORE_INDEX, DEFAULT_IMAGE_TOKEN, SENTINEL_TOKEN) == (-100, "<image>", "<vila/sentinel>")
    assert MEDIA_TOKENS == {"image": "<image>", "video": "<vila/video>"} and make_list("a") == ["a"]
    cfg = SimpleNamespace(num_video_frames=8, fps=0.0)
    for c in json.loads((GOLDEN / "extract_media.json").read_text()):
        parts = build_prompt_parts([tuple(x) for x in c["spec"]], tmp_path, llava.Image, llava.Video)
        conv = [{"from": "human", "value": parts if len(parts) > 1 or not isinstance(parts[0], str) else parts[0]}]
        media = extract_media(conv, cfg)
        assert conv[0]["value"] == c["text"] and [list(im.size) for im in media["image"]] == c["image_sizes"], c["label"]
    want = json.loads((GOLDEN / "tokenizer_setup.json").read_text())
    tok = loading.prepare_tokenizer(toy_chat_tokenizer())
    assert sorted(infer_stop_tokens(tok)) == want["stop_tokens"]
    for case in want["conversations"]:
        ids = tokenize_conversation(copy.deepcopy(case["conversation"]), tok, **case["kwargs"])
        assert ids.tolist() == case["input_ids"], case["kwargs"]
    with pytest.raises(ValueError):
        tokenize_conversation([{"from": "robot", "value": "x"}], tok)


def test_unsupported_generation_options_are_refused_not_swallowed():
    """HF `generate` options this decode loop does not implement (beams, repetition penalty, stopping
    criteria ...) must raise, while every neutral setting — including a stock GenerationConfig, which the
    reference's default_generation_config copies — passes."""
    from transformers import GenerationConfig
    from vila_b200.model.qwen2 import Qwen2ForCausalLM, unsupported_generation_options as refused
    assert refused(None, {}) == [] and refused(GenerationConfig(), {}) == []
    assert refused(SimpleNamespace(max_new_tokens=8, do_sample=True, temperature=0.2, top_p=0.9), {"num_beams": 1}) == []
    assert refused(GenerationConfig(num_beams=4, repetition_penalty=1.2), {}) == ["num_beams=4", "repetition_penalty=1.2"]
    assert refused(GenerationConfig(num_beams=4), {"num_beams": 1}) == []          # explicit kwargs win, as in HF
    assert refused(None, {"stopping_criteria": []}) == [] and refused(None, {"stopping_criteria": [object]}) != []
    llm = object.__new__(Qwen2ForCausalLM)
    llm.generation_config = None
    with pytest.raises(NotImplementedError, match="num_return_sequences=3"):
        Qwen2ForCausalLM.generate.__wrapped__(llm, torch.zeros(1, 4, 8), num_return_sequences=3)
