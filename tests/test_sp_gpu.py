"""GPU, world_size 2 (NCCL): sequence-parallel prefill (zigzag chunks + in-place KV all-gather into the
paged pool + per-chunk causal FMHA) must reproduce the single-GPU prefill.  Skipped with < 2 GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from vila_b200 import sp
        from vila_b200.model import LlavaLlamaModel, tiny_test_config
        cfg = tiny_test_config(llm_layers=3)
        model = LlavaLlamaModel(cfg, device=f"cuda:{rank}").init_random(5)  # same seed -> same weights
        llm = model.llm
        g = torch.Generator().manual_seed(6)
        emb = (torch.randn(S, cfg.hidden_size, generator=g) * 0.05).to(torch.bfloat16)
        plan = sp.make_plan(S, world, rank)
        padded = torch.zeros(plan.padded_len, cfg.hidden_size, dtype=torch.bfloat16)
        padded[:S] = emb
        runner = sp.SequenceParallelPrefill(llm)
        local = plan.extract_local(padded).cuda()
        hid_local, pool = runner.prefill_hidden(local, plan)
        logits = runner.last_token_logits(hid_local, plan)
        gathered = [torch.empty_like(hid_local) for _ in range(world)]
        dist.all_gather(gathered, hid_local)
        full = plan.undo_extract_local(torch.stack(gathered))[:S]
        # single-GPU reference on this rank
        cache = llm.new_cache(plan.padded_len)
        ref = llm.prefill_hidden(emb.cuda(), cache)
        ref_logits = llm.logits_from_hidden(ref[-1:])
        err = (full.float() - ref.float()).abs().max().item()
        scale = ref.float().abs().max().item()
        lerr = (logits.float() - ref_logits.float()).abs().max().item()
        # every rank ends with the complete KV (all-gather landed in the paged pool)
        pt = plan.page_table().cuda()
        kv_ok = True
        for j in range(S // 128):
            a = pool[1, 0, pt[j]]
            b = cache.pool[1, 0, cache.page_table[j]]
            kv_ok = kv_ok and bool((a.float() - b.float()).abs().max().item() <= 2 ** -7 * b.float().abs().max().item())
        ret[rank] = (err, scale, lerr, kv_ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S", [1000, 2048])
def test_sp_prefill_matches_single_gpu(S):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), S, ret), nprocs=world, join=True)
    for r in range(world):
        err, scale, lerr, kv_ok = ret[r]
        # same kernels, different tiling of the sequence: bf16-level agreement
        assert err <= 2 ** -6 * scale, (r, err, scale)
        assert lerr <= 2 ** -5 * max(1.0, scale), (r, lerr)
        assert kv_ok
