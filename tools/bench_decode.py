"""GPU: decode-step timing of the two decode engines on NVILA-8B (random init).
Usage: python tools/bench_decode.py [tiny]"""
import sys, os, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vila_b200.model import LlavaLlamaModel, nvila_8b, tiny_test_config
from vila_b200.model.qwen2 import GraphDecoder, MegaDecoder

torch.cuda.set_device(0)
tiny = len(sys.argv) > 1 and sys.argv[1] == "tiny"
cfg = tiny_test_config(llm_layers=3) if tiny else nvila_8b()
model = LlavaLlamaModel(cfg, device="cuda").init_random(0, device_rng=True)
llm = model.llm
emb = (torch.randn(279, cfg.hidden_size, device="cuda") * 0.05).to(torch.bfloat16)
res = {}
for name, cls, kw in (("graph", GraphDecoder, {}), ("mega_s8", MegaDecoder, {"num_splits": 8}),
                      ("mega_s4", MegaDecoder, {"num_splits": 4}), ("mega_s16", MegaDecoder, {"num_splits": 16})):
    dec = cls(llm, 128, **kw)
    toks = None
    for rep in range(3):
        cache = dec.cache_for(279 + 128)
        hid = llm.prefill_hidden(emb, cache)
        dec.start(hid[-1], cache)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dec.run(64)
        dec.run(64)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        toks = dec.tokens(128)
    res[name] = {"ms_per_token": round(ms / 127, 4), "tok_s": round(127 / ms * 1e3, 1), "wall_s": round(wall, 3), "first_tokens": toks[:6]}
    print(name, res[name], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/bench_decode%s.json" % ("_tiny" if tiny else "")).write_text(json.dumps(res, indent=1))
