"""The DEFAULT multi-GPU paths - Ulysses with both all-to-alls fused into the kernels' stores, and cp_ring_kv with the
K/V all-gather fused into the prologue's stores (FullyParallelAttention._forward_ulysses_fused / _forward_ring_fused) -
emulated on the CPU: the ranks of one context-parallel group are THREADS of this process, "peer pointers" are handles
into a registry of the other threads' buffers (tests/cpu_functional_double.py), the device-side barrier is a
threading.Barrier, symmetric-memory allocation is replaced by plain tensors. What runs unmodified is the module's own
host logic: routing plans (row offsets, row bounds, head slots), workspace use and buffer alternation, the
direct-to-cache-slot pointer arithmetic, barrier placement, per-rank ranges. Checked against the single-rank forward
through a KV-cache sequence, like tools/cp_check.py does with the real kernels on real GPUs (where these paths are
bit-identical to one GPU; here the two sides differ by BLAS blocking only)."""
import dataclasses
import threading

import numpy as np
import pytest
import torch

import cpu_functional_double as D

TLS = threading.local()


class _Shared:
    """Per-test state shared by the rank threads."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.lock = threading.Lock()
        self.slots = {}

    def collective(self, key, make):
        """Every rank calls this with the same key; `make()` runs once and returns the list of per-rank objects."""
        with self.lock:
            if key not in self.slots:
                self.slots[key] = make()
        return self.slots[key]


class _UlyssesWS:
    def __init__(self, shared, rank, bufs):
        self.q, self.kv, self.out = bufs["q"][rank], bufs["kv"][rank], bufs["out"][rank]
        self.q_ptrs, self.kv_ptrs, self.out_ptrs = bufs["q_ptrs"], bufs["kv_ptrs"], bufs["out_ptrs"]
        self._b = shared.barrier
        self._proj, self._proj_ptrs, self.proj_calls, self.rank = bufs["proj"], bufs["proj_ptrs"], 0, rank

    def next_proj(self):
        b = self.proj_calls % 2
        self.proj_calls += 1
        return self._proj[b][self.rank], self._proj_ptrs[b]

    def barrier(self):
        self._b.wait()


class _RingWS:
    def __init__(self, shared, rank, bufs):
        self.kv, self.ptrs, self.calls, self._b, self.rank = bufs["kv"], bufs["ptrs"], 0, shared.barrier, rank
        self.mc_ptrs = [0, 0]

    def next(self):
        b = self.calls % 2
        self.calls += 1
        return self.kv[b][self.rank], self.ptrs[b], 0

    def barrier(self):
        self._b.wait()


@pytest.fixture()
def emulation(monkeypatch):
    import magi_b200  # noqa: F401
    import magi_b200.attention as A
    from magi_b200 import context_parallel as cpm
    from magi_b200 import parallel_state as mpu
    monkeypatch.setattr(A, "F", D)
    monkeypatch.setattr(mpu, "get_cp_rank", lambda: TLS.rank)
    monkeypatch.setattr(mpu, "get_cp_world_size", lambda: TLS.shared.world)
    monkeypatch.setattr(mpu, "get_cp_group", lambda: TLS.shared)

    def zeros(*shape):
        return torch.zeros(shape, dtype=torch.bfloat16)

    class FakeUlysses:
        _instances = {}
        available = staticmethod(lambda group, device: True)

        @staticmethod
        def get(group, total_rows, max_local_rows, hq_loc, hkv_loc, heads_q, device):
            def make():
                w = group.world
                b = dict(q=[zeros(total_rows, hq_loc, 128) for _ in range(w)],
                         kv=[zeros(total_rows, hkv_loc, 256) for _ in range(w)],
                         out=[zeros(max_local_rows, heads_q, 128) for _ in range(w)])
                for n in ("q", "kv", "out"):
                    b[n + "_ptrs"] = [D.fake_ptr(t) for t in b[n]]
                b["proj"] = [[zeros(max_local_rows, 2 * heads_q * 128) for _ in range(w)] for _ in range(2)]
                b["proj_ptrs"] = [[D.fake_ptr(t) for t in b["proj"][i]] for i in range(2)]
                return b
            bufs = group.collective(("uly", total_rows, max_local_rows, hq_loc, hkv_loc, heads_q), make)
            key = ("uly_ws", TLS.rank, total_rows, max_local_rows)
            if key not in group.slots:
                group.slots[key] = _UlyssesWS(group, TLS.rank, bufs)
            return group.slots[key]

    class FakeRing:
        _instances = {}
        available = staticmethod(lambda group, device: True)

        @staticmethod
        def get(group, rows, heads_kv, device):
            def make():
                kv = [[zeros(rows, heads_kv, 256) for _ in range(group.world)] for _ in range(2)]
                return dict(kv=kv, ptrs=[[D.fake_ptr(t) for t in kv[b]] for b in range(2)])
            key = ("ring_ws", TLS.rank)
            ws = group.slots.get(key)
            if ws is None or ws.rows < rows:     # like the real class: a larger forward re-creates the workspace
                ws = group.slots[key] = _RingWS(group, TLS.rank, group.collective(("ring", rows, heads_kv), make))
                ws.rows = rows
            return ws

    monkeypatch.setattr(cpm, "UlyssesPeerWorkspace", FakeUlysses)
    monkeypatch.setattr(cpm, "RingGatherWorkspace", FakeRing)

    def alloc(self, sequence_length, batch_size, dtype, device):
        """Stands in for the symmetric-memory allocation + rendezvous of _allocate_key_and_value_memory."""
        shared = TLS.shared
        shape = (sequence_length * batch_size, self.num_query_groups_per_partition, 2 * self.hidden_size_per_attention_head)
        mine = torch.zeros(shape, dtype=dtype)
        with shared.lock:
            shared.slots.setdefault(("cache", self.layer_number), {})[TLS.rank] = (mine, D.fake_ptr(mine))
        shared.barrier.wait()
        peers = shared.slots[("cache", self.layer_number)]
        self._cache_peer_ptrs = [peers[r][1] for r in range(shared.world)]
        self._cache_mc_ptr = 0
        self._cache_peer_of = mine.data_ptr()
        return mine

    monkeypatch.setattr(A.FullyParallelAttention, "_allocate_key_and_value_memory", alloc)
    return A, cpm


def i32(a):
    return torch.as_tensor(np.asarray(a, dtype=np.int32))


@pytest.mark.parametrize("strategy,world,Hq,Hkv,fused_proj", [
    ("cp_ulysses", 2, 4, 2, False), ("cp_ulysses", 4, 4, 2, False), ("cp_ring_kv", 2, 4, 2, False),
    ("cp_ring_kv", 3, 6, 2, False),
    # fused_proj_input under the fused Ulysses path: every rank's epilogue stores its heads into slot 0 of the OWNER's
    # out-projection input "(hn n hd)" (hn = 8 groups), the local cross-attention fills slot 1; double-buffered
    ("cp_ulysses", 2, 8, 2, True), ("cp_ulysses", 4, 16, 4, True)])
def test_fused_paths_match_single_rank_through_a_cache_sequence(emulation, strategy, world, Hq, Hkv, fused_proj):
    A, cpm = emulation
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    hidden, clip = 128, 37

    def build(strat, cp):
        mc = ModelConfig(num_layers=4, hidden_size=hidden, num_attention_heads=Hq, num_query_groups=Hkv)
        m = A.FullyParallelAttention(mc, EngineConfig(cp_size=cp, cp_strategy=strat), layer_number=1)
        for n_ in ("q_layernorm", "k_layernorm"):
            getattr(m, n_).float()
        return m.eval()

    shared = _Shared(world)
    TLS.rank, TLS.shared = 0, _Shared(1)          # the main thread is its own single-rank "group"
    torch.manual_seed(4)
    ref_m = build("none", 1)
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for n, p in ref_m.named_parameters():
            if "layernorm" in n:
                p.copy_((torch.randn(p.shape, generator=g) * 0.2).to(p.dtype))
            elif n.endswith("layer_norm.weight"):
                p.fill_(1.0)
            elif n.endswith("layer_norm.bias"):
                p.zero_()
            else:
                p.copy_((torch.randn(p.shape, generator=g) / p.shape[-1] ** 0.5).to(p.dtype))
    steps = [  # slice_point, dn, k ranges in chunks, caption tokens, update, fwd_extra, distill_ncc
        (0, 2, [(0, 1), (0, 2)], [20, 31], True, False, False),          # no cache involved
        (0, 3, [(0, 1), (0, 2), (1, 3)], [0, 17, 40], True, True, False),  # every row cached: stores go to the slot
        (1, 2, [(0, 2), (1, 3)], [13, 29], False, False, False),          # reads the cache, workspace buffers
        (1, 3, [(0, 2), (1, 3), (3, 4)], [13, 29, 13], True, False, True),  # distill: last chunk not cached -> append
        (1, 2, [(0, 2), (0, 3)], [5, 9], True, False, False),             # second use of the alternate ring buffer
    ]
    inputs, refs = [], []
    inf_ref = InferenceParams(1, 5 * clip)
    inf_ref.key_value_memory_dict[1] = torch.zeros(5 * clip, Hkv, 256, dtype=torch.bfloat16)
    for (sp, dn, kc, yi, upd, extra, ncc) in steps:
        S = dn * clip
        hs = torch.randn(S, 1, hidden, generator=g).to(torch.bfloat16)
        kvs = torch.randn(sum(yi), hidden, generator=g).to(torch.bfloat16)
        ang = torch.rand(S, 48, generator=g) * 6.28
        rope = torch.cat([torch.sin(ang), torch.cos(ang)], -1)
        q_range = np.array([[i * clip, (i + 1) * clip] for i in range(dn)], np.int32)
        k_range = np.array([[s * clip, e * clip] for s, e in kc], np.int32)
        cu_q, cu_k = np.arange(dn + 1) * clip, np.concatenate([[0], np.cumsum(yi)])
        cross = PackedCrossAttnParams(i32(np.stack([cu_q[:-1], cu_q[1:]], 1)), i32(np.stack([cu_k[:-1], cu_k[1:]], 1)),
                                      i32(cu_q), i32(cu_k), clip, 800)
        core = PackedCoreAttnParams(i32(q_range), i32(k_range), q_range, k_range, clip, int(k_range.max() - k_range.min()))
        meta = ModelMetaArgs(H=4, W=4, cp_pad_size=0, cp_split_sizes=None, slice_point=sp, denoising_range_num=dn,
                             range_num=sp + dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=extra,
                             distill_nearly_clean_chunk=ncc, clip_token_nums=clip, enable_cuda_graph=False,
                             core_attn_params=core, cross_attn_params=cross)
        inf_ref.update_kv_cache = upd
        refs.append(ref_m(hs, kvs, inf_ref, rope, meta))
        inputs.append((hs, kvs, rope, meta, upd, dn))

    # the batched, cache-less uncond pass of the 3-CFG models (dit_model.py:413-494): chunks on the batch dim, one range
    # shared by the batch rows - the fused path keeps the "(sq b)" row order through both all-to-alls
    bs = 3
    hs_u = torch.randn(clip, bs, hidden, generator=g).to(torch.bfloat16)
    yi_u = [11, 0, 23]
    kvs_u = torch.randn(sum(yi_u), hidden, generator=g).to(torch.bfloat16)
    ang = torch.rand(clip, 48, generator=g) * 6.28
    rope_u = torch.cat([torch.sin(ang), torch.cos(ang)], -1)
    edges = np.arange(bs + 1, dtype=np.int32) * clip
    rng_u = np.stack([edges[:-1], edges[1:]], 1)
    cu_k = np.concatenate([[0], np.cumsum(yi_u)]).astype(np.int32)
    core_u = PackedCoreAttnParams(i32(rng_u), i32(rng_u), rng_u, rng_u, clip, bs * clip)
    cross_u = PackedCrossAttnParams(i32(rng_u), i32(np.stack([cu_k[:-1], cu_k[1:]], 1)), i32(edges), i32(cu_k), clip, 800)

    def meta_u(split, cross):
        return ModelMetaArgs(H=4, W=4, cp_pad_size=0, cp_split_sizes=split, slice_point=0, denoising_range_num=1,
                             range_num=1, extract_prefix_video_feature=False, fwd_extra_1st_chunk=False,
                             distill_nearly_clean_chunk=False, clip_token_nums=clip, enable_cuda_graph=False,
                             core_attn_params=core_u, cross_attn_params=cross)

    # (the uncond pass shards heads like Ulysses under both strategies: it needs Hkv | cp or cp | Hkv)
    uncond = not fused_proj and (Hkv % world == 0 or world % Hkv == 0)
    ref_u = ref_m(hs_u, kvs_u, None, rope_u, meta_u(None, cross_u)) if uncond else None

    results, errors = {}, []

    def rank_main(rank):
        try:
            TLS.rank, TLS.shared = rank, shared
            m = build(strategy, world)
            m.load_state_dict(ref_m.state_dict())
            inf = InferenceParams(1, 5 * clip)
            outs = []
            for (hs, kvs, rope, meta, upd, dn) in inputs:
                cpp = meta.core_attn_params
                ardf = dict(denoising_range_num=dn, q_range=cpp.q_range, k_range=cpp.k_range, max_seqlen_q=clip,
                            max_seqlen_k=cpp.max_seqlen_k)
                x, _, rope_l, pad, split, core_l, cross_l = cpm.cp_pre_process(
                    world, strategy, hs, torch.zeros(hs.shape[0], 1), rope, None, ardf, cpp, meta.cross_attn_params)
                m_cp = dataclasses.replace(meta, cp_split_sizes=split, cross_attn_params=cross_l)
                inf.update_kv_cache = upd
                if fused_proj:
                    assert m.supports_fused_proj_input(1, x.device)
                    pin = m(x, kvs, inf, rope_l, m_cp, fused_proj_input=True)
                    outs.append((pin.float().clone(), None, split))
                    continue
                core, xo = m(x, kvs, inf, rope_l, m_cp)
                outs.append((core.float().clone(), xo.float().clone(), split))
            unc = None
            if uncond:
                ardf = dict(denoising_range_num=1, q_range=core_u.q_range, k_range=core_u.k_range, max_seqlen_q=clip,
                            max_seqlen_k=bs * clip)
                x, _, rope_l, pad, split, _, cross_l = cpm.cp_pre_process(
                    world, strategy, hs_u, torch.zeros(clip, bs), rope_u, None, ardf, core_u, cross_u)
                core, xo = m(x, kvs_u, None, rope_l, meta_u(split, cross_l))
                unc = (core.float().clone(), xo.float().clone())
            results[rank] = (outs, inf.key_value_memory_dict[1].float().clone(), unc)
        except Exception:                       # pragma: no cover - reported by the main thread
            import traceback
            errors.append(f"rank {rank}:\n{traceback.format_exc()}")
            shared.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, "\n".join(errors)
    assert len(results) == world

    for i, ((hs, kvs, rope, meta, upd, dn), (ref_core, ref_x)) in enumerate(zip(inputs, refs)):
        parts_c = [results[r][0][i][0] for r in range(world)]
        parts_x = [results[r][0][i][1] for r in range(world)]
        split = results[0][0][i][2]
        if fused_proj:    # the reference's cat + rearrange (attn_linear_proj, dit_module.py:1332-1338) of the two outputs
            want = A.FullyParallelAttention.attn_linear_proj_input(ref_core, ref_x).float()
            got = torch.cat(parts_c, 0)
            assert got.shape == want.shape
            assert float((got - want).abs().max()) <= 1e-2, (strategy, world, i)
            continue
        if strategy == "cp_ulysses":             # contiguous shards
            core, xo = torch.cat(parts_c, 0), torch.cat(parts_x, 0)
        else:                                    # every chunk striped across the ranks: local rows are (chunk, stripe)
            sizes = [s_ // dn for s_ in split]
            def unstripe(parts):
                return torch.cat([torch.cat([p_[j * n_:(j + 1) * n_] for p_, n_ in zip(parts, sizes)], 0)
                                  for j in range(dn)], 0)
            core, xo = unstripe(parts_c), unstripe(parts_x)
        e_core = float((core - ref_core.float()).abs().max())
        e_x = float((xo - ref_x.float()).abs().max())
        assert e_core <= 1e-2 and e_x <= 1e-2, (strategy, world, i, e_core, e_x)
    if uncond:   # uncond pass: contiguous sequence shards under both strategies (cp_ring_kv_process)
        core = torch.cat([results[r][2][0] for r in range(world)], 0)
        xo = torch.cat([results[r][2][1] for r in range(world)], 0)
        assert core.shape == ref_u[0].shape
        assert float((core - ref_u[0].float()).abs().max()) <= 1e-2 and float((xo - ref_u[1].float()).abs().max()) <= 1e-2
    # KV caches: Ulysses keeps the rank's head slice, the ring strategy a full replica - all as the single rank left them
    ref_cache = inf_ref.key_value_memory_dict[1].float()
    for r in range(world):
        cache = results[r][1]
        if strategy == "cp_ulysses":
            hn = cache.shape[1]
            lo = r * Hkv // world if world <= Hkv else r // (world // Hkv)
            want = ref_cache[:, lo:lo + hn]
        else:
            want = ref_cache
        assert float((cache[:4 * clip] - want[:4 * clip]).abs().max()) <= 1e-2, (strategy, world, r)
