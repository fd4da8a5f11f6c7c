"""GPU: INTEGRATION.md option B executed — the ctypes stub a VILA maintainer would add
(vila_b200/integration/b200_ops.py) behind the reference's two operator seams: `flash_attn_func` and
the `attn_implementation=` kwarg (through transformers' attention registry, on stock HF SigLIP / Qwen2
modules)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flash_attn_func_stub(cuda):
    from tests.helpers import report_rel
    from tests.test_kernels_gpu import ref_attention
    from vila_b200.integration import b200_ops
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, S, H, Hkv, D, causal) in [(2, 1024, 16, 16, 72, False), (1, 300, 28, 4, 128, True)]:
        q = torch.randn(B, S, H, D, device="cuda", generator=g).to(torch.bfloat16)
        k = torch.randn(B, S, Hkv, D, device="cuda", generator=g).to(torch.bfloat16)
        v = torch.randn(B, S, Hkv, D, device="cuda", generator=g).to(torch.bfloat16)
        out = b200_ops.flash_attn_func(q, k, v, softmax_scale=D ** -0.5, causal=causal)
        report_rel(f"b200_ops.flash_attn_func d={D}", out, ref_attention(q, k, v, causal, D ** -0.5), 1.5e-2)


def test_attn_implementation_seam_on_stock_hf_modules(cuda):
    """attn_implementation="vila_b200" on transformers' own SiglipVisionModel and Qwen2ForCausalLM
    (the classes the reference instantiates) vs attn_implementation="sdpa", same weights."""
    from transformers import Qwen2Config, Qwen2ForCausalLM, SiglipVisionConfig, SiglipVisionModel

    from vila_b200.integration import b200_ops
    name = b200_ops.register_hf_attention()
    torch.manual_seed(0)
    vc = SiglipVisionConfig(hidden_size=288, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                            image_size=224, patch_size=14)
    vc._attn_implementation = "sdpa"
    with torch.device("cuda"):
        vit = SiglipVisionModel(vc).to(torch.bfloat16).eval()
    px = torch.randn(2, 3, 224, 224, device="cuda").to(torch.bfloat16)
    with torch.inference_mode():
        ref = vit(pixel_values=px, output_hidden_states=True).hidden_states[-1].float()
        vit.config._attn_implementation = name
        for m in vit.modules():
            if hasattr(m, "config") and hasattr(m.config, "_attn_implementation"):
                m.config._attn_implementation = name
        got = vit(pixel_values=px, output_hidden_states=True).hidden_states[-1].float()
    assert (got - ref).abs().max().item() <= 2 ** -5 * ref.abs().max().item()
    lc = Qwen2Config(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                     num_key_value_heads=2, vocab_size=1024, max_position_embeddings=4096)
    lc._attn_implementation = "sdpa"
    with torch.device("cuda"):
        llm = Qwen2ForCausalLM(lc).to(torch.bfloat16).eval()
    ids = torch.randint(0, 1024, (1, 300), device="cuda")
    with torch.inference_mode():
        ref = llm(input_ids=ids).logits.float()
        for m in llm.modules():
            if hasattr(m, "config") and hasattr(m.config, "_attn_implementation"):
                m.config._attn_implementation = name
        got = llm(input_ids=ids).logits.float()
    assert (got - ref).abs().max().item() <= 2 ** -5 * max(1.0, ref.abs().max().item())
