"""The drop-in FullyParallelAttention MODULE under context parallelism on CPU: world size 2 over gloo, its kernels
replaced by the oracle-backed double of tests/cpu_functional_double.py. What runs for real is everything the module
does on the host between the kernels - projections, cache plans, send layouts, the all-to-alls / ring exchange through
torch.distributed, head groups, per-rank range handling, the strategy switch of dit_module.py:1163-1244 - i.e. the code
that tools/cp_check.py exercises on real GPUs, here as part of the CPU suite.

(1) cp_ulysses and cp_shuffle_overlap against the reference's OWN per-rank outputs (tests/golden/ulysses_cp2_*,
    cso_cp2_*: core_attn_out and xattn_out of the reference module run on 2 gloo ranks);
(2) all three strategies (incl. cp_ring_kv) through a KV-cache sequence (cache fill with the extra clean chunk, window
    advance, update on / off, distill self-only range, an empty caption) against the single-rank forward of the same
    module (oracle O3 of SURVEY.md section 8c)."""
import dataclasses
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist

from test_cp_gloo import _setup, _spawn


def _module(g, strategy, world, Hq, Hkv, hidden, fp8=False):
    import cpu_functional_double as D
    import magi_b200.attention as A
    import magi_b200.fp8_linear as L8
    from magi_b200.common import EngineConfig, ModelConfig
    A.F = D                                     # the module's kernels -> CPU doubles
    L8.F = D
    mc = ModelConfig(num_layers=4, hidden_size=hidden, num_attention_heads=Hq, num_query_groups=Hkv)
    m = A.FullyParallelAttention(mc, EngineConfig(cp_size=world if strategy != "none" else 1, cp_strategy=strategy,
                                                  fp8_quant=fp8), layer_number=1)
    for n_ in ("q_layernorm", "k_layernorm"):   # dit_model.py:621-633
        getattr(m, n_).float()
    if g is not None:
        m.load_state_dict({name: g.torch(key) for name, key in g.weights().items()}, strict=True)
    return m.eval()


def _i32(a):
    return torch.as_tensor(np.asarray(a, dtype=np.int32))


def _meta(clip, sp, dn, q_range, k_range, cross, split=None, pad=0, extra=False, ncc=False, max_q=None):
    from magi_b200.common import ModelMetaArgs, PackedCoreAttnParams
    q_range, k_range = np.asarray(q_range, np.int32), np.asarray(k_range, np.int32)
    core = PackedCoreAttnParams(_i32(q_range), _i32(k_range), q_range, k_range, max_q or clip,
                                int(k_range.max() - k_range.min()))
    return ModelMetaArgs(H=4, W=4, cp_pad_size=pad, cp_split_sizes=split, slice_point=sp, denoising_range_num=dn,
                         range_num=sp + dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=extra,
                         distill_nearly_clean_chunk=ncc, clip_token_nums=clip, enable_cuda_graph=False,
                         core_attn_params=core, cross_attn_params=cross)


def _close(got, ref, what):
    got, ref = got.float().numpy(), np.asarray(ref, np.float32).reshape(got.shape)
    err = float(np.abs(got - ref).max())
    rel = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-9))
    assert err < 6e-2 and rel < 1.5e-2, (what, err, rel)


def _golden_worker(rank, world, port, strategy, errq):
    try:
        _setup(rank, world, port)
        import magi_oracle as O
        from golden_io import Golden
        from magi_b200.common import PackedCrossAttnParams
        stem = "ulysses_cp2" if strategy == "cp_ulysses" else "cso_cp2"
        g0, gr = Golden(f"{stem}_rank0.npz"), Golden(f"{stem}_rank{rank}.npz")
        Hq, Hkv = [int(v) for v in g0["heads"]]
        clip, dn = int(g0["clip"]), 3
        hidden = g0.raw("w.linear_qkv.layer_norm.weight").shape[0]
        m = _module(g0, strategy, world, Hq, Hkv, hidden)
        cross = PackedCrossAttnParams(q_ranges=_i32(gr.raw("cross.q_ranges")), kv_ranges=_i32(gr.raw("cross.kv_ranges")),
                                      cu_seqlens_q=None, cu_seqlens_kv=None,
                                      max_seqlen_q=int(gr["cross.max_seqlen_q"]), max_seqlen_kv=800)
        split = [int(v) for v in gr["split_sizes"]]
        if strategy == "cp_ulysses":
            meta = _meta(clip, 0, dn, O.core_q_ranges(clip, dn, 1), g0.raw("k_range"), cross, split=split)
        else:
            meta = _meta(clip, 0, dn, gr.raw("core.q_range"), gr.raw("core.k_range"), cross, split=split,
                         pad=int(gr["pad"]), max_q=int(gr["core.max_seqlen_q"]))
        core, xattn = m(gr.torch("x_local"), g0.torch("key_value_states"), None, gr.torch("rope_local"), meta)
        _close(core, gr["core_attn_out"], f"{strategy} core_attn_out rank {rank}")
        _close(xattn, gr["xattn_out"], f"{strategy} xattn_out rank {rank}")
        dist.barrier()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("strategy,port", [("cp_ulysses", 29641), ("cp_shuffle_overlap", 29642)])
def test_module_matches_reference_per_rank_outputs(strategy, port):
    _spawn(_golden_worker, 2, (port, strategy))


class _HostOffload:
    """CPU stand-in for attention.KVOffload (which stages through CUDA streams): same interface, synchronous copies.
    Lets the module's kv_offload branches (pinned-host cache, staged prefix, write-back) run in the CPU suite."""

    def __init__(self, device):
        self.stores = self.prefixes = 0

    @classmethod
    def of(cls, inference_params, device):
        st = getattr(inference_params, "kv_offload_state", None)
        if st is None:
            st = inference_params.kv_offload_state = cls(device)
        return st

    def prefix(self, layer, cache, rows, next_layer_cache):
        self.prefixes += 1
        return cache[:rows].clone()

    def store(self, layer, cache, kv_new, row_offset, rows):
        if rows > 0:
            cache[row_offset:row_offset + rows] = kv_new[:rows]
            self.stores += 1


def _sequence_worker(rank, world, port, strategy, errq, offload=False, fp8=False):
    try:
        _setup(rank, world, port)
        from magi_b200 import context_parallel as cpm
        from magi_b200.common import InferenceParams, PackedCrossAttnParams
        Hq, Hkv, hidden, clip = 4, 2, 128, 37          # 37: shards and chunk stripes are ragged
        torch.manual_seed(4)
        ref_m = _module(None, "none", 1, Hq, Hkv, hidden, fp8=fp8)
        g = torch.Generator().manual_seed(99)
        with torch.no_grad():
            for n, p in ref_m.named_parameters():
                if p.dtype == torch.float8_e4m3fn:       # fp8_quant: weights in e4m3 units of weight_scale
                    p.copy_((torch.randn(p.shape, generator=g) * 120.0).clamp_(-448, 448).to(p.dtype))
                elif n.endswith("weight_scale"):
                    p.fill_(1.0 / (120.0 * hidden ** 0.5))
                elif n.endswith("input_scale"):
                    p.fill_(5.0 / 448.0)
                elif n.endswith("smooth_scale"):
                    p.copy_(1.0 + 0.2 * torch.rand(p.shape, generator=g))
                elif "layernorm" in n:
                    p.copy_((torch.randn(p.shape, generator=g) * 0.2).to(p.dtype))
                elif n.endswith("layer_norm.weight"):
                    p.fill_(1.0)
                elif n.endswith("layer_norm.bias"):
                    p.zero_()
                else:
                    p.copy_((torch.randn(p.shape, generator=g) / p.shape[-1] ** 0.5).to(p.dtype))
        cp_m = _module(None, strategy, world, Hq, Hkv, hidden, fp8=fp8)
        cp_m.load_state_dict(ref_m.state_dict())
        inf_ref, inf_cp = InferenceParams(1, 5 * clip), InferenceParams(1, 5 * clip)
        if offload:
            # kv_offload (dit_module.py:803-811): the rank's cache lives in host memory and is staged through the
            # offload manager; the resident single-rank forward stays the checker
            import magi_b200.attention as A
            A.KVOffload = _HostOffload
            torch.cuda.current_device = lambda: 0
            cp_m.engine_config.kv_offload = True
            inf_cp.key_value_memory_dict[1] = torch.zeros(5 * clip, cp_m.num_query_groups_per_partition, 256,
                                                          dtype=torch.bfloat16)
        steps = [  # slice_point, dn, k ranges in chunks, caption tokens, update, fwd_extra, distill_ncc
            (0, 2, [(0, 1), (0, 2)], [20, 31], True, False, False),
            (0, 3, [(0, 1), (0, 2), (1, 3)], [0, 17, 40], True, True, False),
            (1, 2, [(0, 2), (1, 3)], [13, 29], False, False, False),
            (1, 3, [(0, 2), (1, 3), (3, 4)], [13, 29, 13], True, False, True),
        ]
        def post(t, m_cp):
            if strategy != "cp_ulysses":
                return cpm.cp_post_process(world, strategy, t, m_cp)
            # gloo's all_gather wants equal shapes (NCCL does not): pad the ragged Ulysses shards for the exchange
            sizes = m_cp.cp_split_sizes
            padded = torch.zeros((max(sizes),) + tuple(t.shape[1:]), dtype=t.dtype)
            padded[:t.shape[0]] = t
            parts = [torch.empty_like(padded) for _ in sizes]
            dist.all_gather(parts, padded)
            return torch.cat([p_[:n_] for p_, n_ in zip(parts, sizes)], 0)

        for (sp, dn, kc, yi, upd, extra, ncc) in steps:
            S = dn * clip
            hs = torch.randn(S, 1, hidden, generator=g).to(torch.bfloat16)
            kvs = torch.randn(sum(yi), hidden, generator=g).to(torch.bfloat16)
            ang = torch.rand(S, 48, generator=g) * 6.28
            rope = torch.cat([torch.sin(ang), torch.cos(ang)], -1)
            q_range = [[i * clip, (i + 1) * clip] for i in range(dn)]
            k_range = [[s * clip, e * clip] for s, e in kc]
            cu_q = np.arange(dn + 1) * clip
            cu_k = np.concatenate([[0], np.cumsum(yi)])
            cross = PackedCrossAttnParams(_i32(np.stack([cu_q[:-1], cu_q[1:]], 1)), _i32(np.stack([cu_k[:-1], cu_k[1:]], 1)),
                                          _i32(cu_q), _i32(cu_k), clip, 800)
            m_full = _meta(clip, sp, dn, q_range, k_range, cross, extra=extra, ncc=ncc)
            inf_ref.update_kv_cache = inf_cp.update_kv_cache = upd
            ref_core, ref_x = ref_m(hs, kvs, inf_ref, rope, m_full)
            cpp = m_full.core_attn_params
            ardf = dict(denoising_range_num=dn, q_range=cpp.q_range, k_range=cpp.k_range, max_seqlen_q=clip,
                        max_seqlen_k=cpp.max_seqlen_k)
            x, _, rope_l, pad, split, core_l, cross_l = cpm.cp_pre_process(
                world, strategy, hs, torch.zeros(S, 1), rope, None, ardf, cpp, cross)
            m_cp = dataclasses.replace(m_full, cp_split_sizes=split, cross_attn_params=cross_l)
            if strategy == "cp_shuffle_overlap":
                m_cp = dataclasses.replace(m_cp, cp_pad_size=pad, core_attn_params=core_l)
            core, xo = cp_m(x, kvs, inf_cp, rope_l, m_cp)
            core, xo = post(core.float().contiguous(), m_cp), post(xo.float().contiguous(), m_cp)
            # Ulysses / CSO keep every row's key order (only BLAS blocking differs between the two sides); the ring
            # merges partial results with one extra bf16 rounding per step
            bound = 2.5e-2 if strategy == "cp_ring_kv" else 1e-2
            e_core = float((core - ref_core.float()).abs().max())
            e_x = float((xo - ref_x.float()).abs().max())
            assert e_core <= bound and e_x <= 1e-2, (strategy, (sp, dn, kc), e_core, e_x)
        if offload:
            st = inf_cp.kv_offload_state
            assert st.stores >= 2 and st.prefixes >= 2, (st.stores, st.prefixes)
        dist.barrier()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("strategy,world,port", [("cp_ulysses", 2, 29651), ("cp_shuffle_overlap", 2, 29652),
                                                 ("cp_ring_kv", 2, 29653), ("cp_ring_kv", 3, 29654),
                                                 ("cp_ulysses", 4, 29655)])   # cp 4 > 2 kv heads: KV-head replication
def test_module_kv_cache_sequence_matches_single_rank(strategy, world, port):
    _spawn(_sequence_worker, world, (port, strategy))


@pytest.mark.parametrize("strategy,port", [("cp_ulysses", 29661), ("cp_shuffle_overlap", 29662)])
def test_module_kv_offload_branches_match_single_rank(strategy, port):
    """The same sequence with engine_config.kv_offload: host-resident caches staged through the offload manager."""
    _spawn(_sequence_worker_offload, 2, (port, strategy))


def _sequence_worker_offload(rank, world, port, strategy, errq):
    _sequence_worker(rank, world, port, strategy, errq, offload=True)


@pytest.mark.parametrize("strategy,port", [("cp_ulysses", 29671), ("cp_ring_kv", 29672)])
def test_module_fp8_quant_sequence_matches_single_rank(strategy, port):
    """The same sequence with engine_config.fp8_quant (fp8 q / qx / k / v linears, dit_module.py:405-410): the
    single-rank forward takes the fused fp8 projection GEMM, the context-parallel paths over gloo project through the
    fp8 linear modules one by one - same quantised activations, same results."""
    _spawn(_sequence_worker_fp8, 2, (port, strategy))


def _sequence_worker_fp8(rank, world, port, strategy, errq):
    _sequence_worker(rank, world, port, strategy, errq, fp8=True)
