"""CPU / gloo (world_size 2 and 4): host logic of the sequence-parallel prefill — zigzag partition,
page-table permutation and the in-place all-gather layout — without any GPU kernel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vila_b200 import sp


def test_plan_matches_reference_zigzag():
    """prepare_zigzag_ring_attn_inputs / extract_local (zigzag_ring_attn/prepare_inputs.py:20-23):
    value.chunk(2*world)[rank] ++ value.chunk(2*world)[2*world-1-rank]."""
    for world in (1, 2, 4, 8):
        S = 2 * world * 128 * 3
        x = torch.arange(S)
        for rank in range(world):
            plan = sp.make_plan(S, world, rank)
            chunks = x.chunk(2 * world)
            want = torch.cat([chunks[rank], chunks[2 * world - 1 - rank]])
            assert torch.equal(plan.extract_local(x), want)
            assert torch.equal(plan.local_positions().long(), want)


def test_padding_and_ownership():
    plan = sp.make_plan(65800, 8, 3)
    assert plan.padded_len % (16 * 128) == 0 and plan.padded_len >= 65800
    assert plan.padded_len - 65800 < 16 * 128
    assert plan.chunk % 128 == 0
    last_owner = plan.owner_of(plan.padded_len - 1)
    assert last_owner == 0  # the last chunk (2P-1) belongs to rank 0
    # causal work is balanced: sum over a rank's two chunks of the KV length they attend to
    work = []
    for r in range(8):
        work.append(sum(e for _, e in plan.local_chunks(r)))
    assert max(work) == min(work)
    # frame sharding covers all frames exactly once
    spans = [sp.shard_frames(250, 8, r) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 250
    assert all(spans[i][1] == spans[i + 1][0] for i in range(7))


def test_page_table_is_a_permutation_and_undo():
    for world in (2, 4):
        S = 2 * world * 128 * 2
        tables = sp.make_plan(S, world, 0).page_table()
        assert sorted(tables.tolist()) == list(range(S // 128))
        plan = sp.make_plan(S, world, 1)
        x = torch.randn(S, 3)
        gathered = torch.stack([plan.extract_local(x, r) for r in range(world)])
        assert torch.equal(plan.undo_extract_local(gathered), x)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Hkv, D = 2, 4
        plan = sp.make_plan(S, world, rank)
        g = torch.Generator().manual_seed(0)
        kv_global = torch.randn(plan.padded_len, Hkv, D, generator=g)  # same on every rank
        pool = torch.zeros(plan.padded_len // 128, 128, Hkv, D)
        pt = plan.page_table()
        # each rank writes ONLY its own tokens through the page table (what rope_kv_append does)
        for pos in plan.local_positions().tolist():
            pool[pt[pos // 128], pos % 128] = kv_global[pos]
        n_local = 2 * plan.chunk_pages
        lo = rank * n_local
        # pages written so far must all lie inside this rank's all-gather region
        touched = pool.abs().sum(dim=(1, 2, 3)) > 0
        assert touched[lo:lo + n_local].all() and int(touched.sum()) == n_local
        dist.all_gather_into_tensor(pool.view(-1), pool[lo:lo + n_local].clone().view(-1))
        # reading block j through the page table now yields the global order on every rank
        rebuilt = torch.cat([pool[pt[j]] for j in range(plan.padded_len // 128)])
        ok = torch.equal(rebuilt, kv_global)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_inplace_allgather_into_paged_pool_gloo(world):
    S = 2 * world * 128 * 2 - 77  # not a multiple: exercises padding
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, S, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _runner_worker(rank, world, port, n_frames, ret):
    """host logic of SequenceParallelPrefill that needs a process group but no GPU: ragged frame
    gather, last-token ownership / broadcast, decode-capable page order."""
    from types import SimpleNamespace
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        llm = SimpleNamespace(device=torch.device("cpu"), dtype=torch.float32, vocab_size=8)
        runner = sp.SequenceParallelPrefill(llm)
        assert runner.world == world and runner.rank == rank
        g = torch.Generator().manual_seed(1)
        feats = torch.randn(n_frames, 5, 3, generator=g)            # same on every rank
        f0, f1 = sp.shard_frames(n_frames, world, rank)
        got = runner.gather_frame_features(feats[f0:f1].clone(), n_frames)
        ok = torch.equal(got, feats)
        S = 2 * world * 128 * 2 - 77
        plan = sp.make_plan(S, world, rank)
        hid = torch.zeros(2 * plan.chunk, 4)
        row = runner.local_row_of(plan, S - 1)
        if row is not None:
            hid[row] = torch.tensor([1.0, 2.0, 3.0, 4.0])
        last = runner.last_token_hidden(hid, plan)
        ok = ok and torch.equal(last, torch.tensor([1.0, 2.0, 3.0, 4.0]))
        ok = ok and ((row is not None) == (plan.owner_of(S - 1) == rank))
        order = sp.sp_cache_page_order(plan, plan.padded_len // 128 + 3)
        ok = ok and sorted(order) == list(range(len(order))) and order[-3:] == list(range(len(order) - 3, len(order)))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 7), (4, 9), (2, 1)])
def test_runner_host_logic_gloo(world, n_frames):
    ret = mp.Manager().dict()
    mp.spawn(_runner_worker, args=(world, _free_port(), n_frames, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_sp_registry_defaults_off():
    assert not sp.sequence_parallel_enabled()
    sp.set_sequence_parallel_group(None)
    assert not sp.sequence_parallel_enabled()  # no process group initialised in this process
    sp.set_sequence_parallel_group(None, enabled=False)


def test_serving_page_allocator():
    """host bookkeeping of the continuous-batching engine (vila_b200/serving.py)"""
    from vila_b200.serving import PageAllocator
    a = PageAllocator(6)
    p1 = a.alloc(2)
    p2 = a.alloc(3)
    assert p1 == [0, 1] and p2 == [2, 3, 4] and a.available == 1
    with pytest.raises(MemoryError):
        a.alloc(2)
    a.release(p1)
    assert a.available == 3 and sorted(a.alloc(3)) == [0, 1, 5]


class _FakeDecoder:
    """Stands in for serving.BatchedDecoder on a machine without a GPU: same host-visible surface
    (slots, allocator, admit / run / release / generated), a deterministic token rule instead of kernels:
    request with prompt length S emits S, S+1, S+2, ... so EOS can be placed at a chosen step."""

    def __init__(self, slots, pages_per_slot, total_pages):
        from vila_b200.serving import PageAllocator
        self.slots, self.pages_per_slot = slots, pages_per_slot
        self.allocator = PageAllocator(total_pages)
        self.slot_pages = [[] for _ in range(slots)]
        self.state = [None] * slots
        self.max_in_flight = 0
        self.admitted = []

    def capture(self):
        pass

    def _ensure(self, s, n_tokens):
        need = min(self.pages_per_slot, -(-n_tokens // 128)) - len(self.slot_pages[s])
        if need > 0:
            self.slot_pages[s].extend(self.allocator.alloc(need))

    def admit(self, s, emb):
        assert self.state[s] is None
        S = emb.shape[0]
        self._ensure(s, S + 1)
        self.state[s] = [S, [S]]  # (tokens cached, generated ids)
        self.admitted.append((s, S))
        self.max_in_flight = max(self.max_in_flight, sum(x is not None for x in self.state))

    def run(self, n):
        for s, st in enumerate(self.state):
            if st is not None:
                self._ensure(s, st[0] + n + 1)
                for _ in range(n):
                    st[1].append(st[1][-1] + 1)
                st[0] += n

    def generated(self, s):
        return list(self.state[s][1])

    def release(self, s):
        self.allocator.release(self.slot_pages[s])
        self.slot_pages[s], self.state[s] = [], None


def test_continuous_batching_scheduler_host_logic():
    """generate_batch (vila_b200/serving.py): admission order, slot reuse, EOS cut (EOS included),
    max_new_tokens cut, admission held back while the pool cannot hold a request's whole budget, every
    page returned at the end."""
    import torch
    from vila_b200.serving import generate_batch
    lens = [100, 300, 50, 700, 120, 260, 900]
    prompts = [torch.zeros(n, 4) for n in lens]
    dec = _FakeDecoder(slots=3, pages_per_slot=8, total_pages=12)
    # request r emits lens[r], lens[r]+1, ...: EOS ids hit request 0 at its 6th token and request 3 at its 1st
    out = generate_batch(None, prompts, max_new_tokens=20, eos_token_ids=(105, 700), slots=3, check_every=4,
                         decoder=dec)
    assert out[0] == [100, 101, 102, 103, 104, 105]
    assert out[3] == [700]
    for r in (1, 2, 4, 5, 6):
        assert out[r] == list(range(lens[r], lens[r] + 20))
    assert [s for _, s in dec.admitted] == lens              # FIFO admission
    assert dec.max_in_flight <= 3 and len({s for s, _ in dec.admitted}) == 3   # slots are reused
    assert dec.allocator.available == 12 and all(x is None for x in dec.state)
    # a pool too small for two budgets at once serialises the requests instead of overflowing
    dec2 = _FakeDecoder(slots=3, pages_per_slot=8, total_pages=8)
    out2 = generate_batch(None, prompts[:4], max_new_tokens=20, eos_token_ids=(), slots=3, check_every=4, decoder=dec2)
    assert [len(o) for o in out2] == [20] * 4 and dec2.allocator.available == 8
    # a request that can never fit a slot is refused up front
    with pytest.raises(ValueError):
        generate_batch(None, [torch.zeros(1020, 4)], max_new_tokens=20, slots=1, check_every=4,
                       decoder=_FakeDecoder(1, 8, 8))
