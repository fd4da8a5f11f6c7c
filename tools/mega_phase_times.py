"""GPU profiling aid: per-phase clock64 stamps of one CTA of the decode mega-kernel (one token)."""
import sys, os, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200.model import LlavaLlamaModel, nvila_8b
from vila_b200.model.qwen2 import MegaDecoder
torch.cuda.set_device(0)
cfg = nvila_8b()
model = LlavaLlamaModel(cfg, device="cuda").init_random(0, device_rng=True)
llm = model.llm
emb = (torch.randn(279, cfg.hidden_size, device="cuda") * 0.05).to(torch.bfloat16)
nph = 4 * cfg.llm_cfg.num_hidden_layers + 1
out = {}
for cta in (0, 40, 147):
    dbg = torch.zeros(nph * 6, dtype=torch.int64, device="cuda")
    os.environ["VILA_B200_MEGA_DEBUG"] = "%x,%d" % (dbg.data_ptr(), cta)
    dec = MegaDecoder(llm, 128, num_splits=int(os.environ.get("SPLITS", "8")))
    cache = dec.cache_for(279 + 128)
    hid = llm.prefill_hidden(emb, cache)
    dec.start(hid[-1], cache)
    dec.run(3)      # first call produces tokens 2 and 3: two launches? no: one launch of 2 tokens
    torch.cuda.synchronize()
    dbg.zero_()
    dec.run(1)      # one token, stamps recorded
    torch.cuda.synchronize()
    t = dbg.view(nph, 6).cpu().double() / 1.965e3  # us at 1965 MHz
    names = ["qkv", "o", "gu", "down"]
    agg = {}
    for g in range(nph):
        kind = "lm_head" if g == nph - 1 else names[g & 3]
        a = agg.setdefault(kind, [0.0] * 5 + [0])
        for k in range(5):
            a[k] += float(t[g, k + 1] - t[g, k])
        a[5] += 1
    rows = {k: {"attn+barrier": round(v[0] / v[5], 2), "stage_x": round(v[1] / v[5], 2), "consume": round(v[2] / v[5], 2),
                "reduce+epilogue": round(v[3] / v[5], 2), "grid_barrier": round(v[4] / v[5], 2), "n": v[5]} for k, v in agg.items()}
    total = float(t[nph - 1, 5] - t[0, 0])
    out[cta] = {"per_phase_us": rows, "token_us": round(total, 1)}
    print("cta", cta, json.dumps(out[cta]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/mega_phase_times.json").write_text(json.dumps(out, indent=1))
