"""Multi-GPU parity of the context-parallel paths with the real kernels (run on the GPU box):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29700 tools/cp_check.py

Every rank builds the same FullyParallelAttention weights and the same full-sequence inputs, runs the cp_ulysses, cp_shuffle_overlap and
cp_ring_kv forwards on its shard through a KV-cache sequence (cache fill, window advance, update on/off, distill extra
range), all-gathers the outputs and compares them with the single-GPU (cp_strategy="none") forward computed locally
(oracle O3 of SURVEY.md section 8c). Ulysses, cp_shuffle_overlap and the fused-gather cp_ring_kv must be bit-identical (same
kernels, same per-row inputs and key order); the NCCL ring is bounded by one bf16 rounding per merge step.
MAGI_RING_MULTICAST=0 keeps the fused gather on per-peer stores when an NVSwitch multicast mapping exists."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import magi_b200  # noqa: F401
    from magi_b200 import context_parallel as cpm
    from magi_b200 import parallel_state as mpu
    from magi_b200.attention import FullyParallelAttention
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    mpu.initialize_model_parallel(cp_size=world)
    FP8 = os.environ.get("CP_CHECK_FP8") == "1"      # engine_config.fp8_quant: fp8 q / qx / k / v (fused e4m3 GEMM)
    GEMM = os.environ.get("CP_CHECK_FUSED_GEMM") == "1" or FP8

    Hq, Hkv, hidden, clip = 24, 8, 512, 1111       # 1111: shards and tiles are ragged everywhere
    mc = ModelConfig(num_layers=4, hidden_size=hidden, num_attention_heads=Hq, num_query_groups=Hkv)

    def build(strategy, cp, offload=False):
        torch.manual_seed(1234)
        m = FullyParallelAttention(mc, EngineConfig(cp_size=cp, cp_strategy=strategy, ulysses_overlap_degree=overlap,
                                                    cp_fused_all_to_all=fused, kv_offload=offload,
                                                    # like with like: the context-parallel paths project through
                                                    # cuBLAS + the prologue kernel, so the single-GPU side does too
                                                    # (the fused projection GEMM differs by single bf16 ulps)
                                                    # CP_CHECK_FUSED_GEMM=1: the fused projection GEMM on both
                                                    # sides instead (under fused Ulysses its epilogue IS the input
                                                    # all-to-all) - only that variant is run then
                                                    fused_qkv_proj=GEMM, fp8_quant=FP8), layer_number=1)
        g = torch.Generator().manual_seed(99)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if p.dtype == torch.float8_e4m3fn:
                    p.copy_((torch.randn(p.shape, generator=g) * 120.0).clamp_(-448, 448).to(p.dtype))
                elif n.endswith("weight_scale"):
                    p.fill_(1.0 / (120.0 * hidden ** 0.5))
                elif n.endswith("input_scale"):
                    p.fill_(5.0 / 448.0)
                elif n.endswith("smooth_scale"):
                    p.fill_(1.0)
                elif "layernorm" in n:
                    p.copy_((torch.randn(p.shape, generator=g) * 0.2).to(p.dtype))
                elif n.endswith("layer_norm.weight"):
                    p.fill_(1.0)
                elif n.endswith("layer_norm.bias"):
                    p.zero_()
                else:
                    p.copy_((torch.randn(p.shape, generator=g) / p.shape[-1] ** 0.5).to(p.dtype))
        for n_ in ("q_layernorm", "k_layernorm"):
            getattr(m, n_).float()
        return m.to(dev).eval()

    i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32), device=dev)

    def meta(sp, dn, k_chunks, y_index, extra=False, ncc=False, split=None, cross=None):
        q_range = np.array([[i * clip, (i + 1) * clip] for i in range(dn)], dtype=np.int32)
        k_range = np.array([[s * clip, e * clip] for s, e in k_chunks], dtype=np.int32)
        cu_k = np.concatenate([[0], np.cumsum(y_index)]).astype(np.int32)
        core = PackedCoreAttnParams(i32(q_range), i32(k_range), q_range, k_range, clip, int(k_range.max() - k_range.min()))
        if cross is None:
            cu_q = np.arange(dn + 1, dtype=np.int32) * clip
            cross = PackedCrossAttnParams(i32(np.stack([cu_q[:-1], cu_q[1:]], 1)), i32(np.stack([cu_k[:-1], cu_k[1:]], 1)),
                                          i32(cu_q), i32(cu_k), clip, 800)
        return ModelMetaArgs(H=1, W=1, cp_pad_size=0, cp_split_sizes=split, slice_point=sp, denoising_range_num=dn,
                             range_num=sp + dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=extra,
                             distill_nearly_clean_chunk=ncc, clip_token_nums=clip, enable_cuda_graph=False,
                             core_attn_params=core, cross_attn_params=cross)

    steps = [  # (slice_point, dn, k ranges in chunks, caption tokens, update, fwd_extra, distill_ncc)
        (0, 2, [(0, 1), (0, 2)], [50, 61], True, False, False),
        (0, 3, [(0, 1), (0, 2), (1, 3)], [0, 37, 300], True, True, False),
        (1, 2, [(0, 2), (1, 3)], [33, 90], False, False, False),
        (1, 3, [(0, 2), (1, 3), (3, 4)], [33, 90, 33], True, False, True),
    ]
    results = {}
    # CP_CHECK_STRATEGIES=cp_ring_kv,... restricts the run (the 8-GPU box time is charged 8x)
    chosen = os.environ.get("CP_CHECK_STRATEGIES", "cp_ulysses,cp_shuffle_overlap,cp_ring_kv").split(",")
    if GEMM:
        chosen = ["cp_ulysses"]
    for strategy in [s_ for s_ in ("cp_ulysses", "cp_shuffle_overlap", "cp_ring_kv") if s_ in chosen]:
        kv_loc = max(1, Hkv // world)
        second = kv_loc if kv_loc > 1 else 3      # whole kv heads per group, or 3 q-head groups in the MQA case
        # cp_ulysses: all-to-alls fused into the kernels (NVLink peer stores), then the NCCL path at two overlap degrees
        # cp_ring_kv: K/V all-gather fused into the prologue's stores (one attention launch), then the NCCL ring
        # CP_CHECK_KV_OFFLOAD=1 adds, for the two reference strategies, the NCCL path with the KV caches in pinned host
        # memory (kv_offload, dit_module.py:803-811) against the resident single-GPU forward
        off_variants = ()
        if os.environ.get("CP_CHECK_KV_OFFLOAD") == "1":   # cp_ring_kv: on the fused gather path (the NCCL ring has none)
            off_variants = ((1, True, True),) if strategy == "cp_ring_kv" else ((1, False, True),)
        for overlap, fused, offload in ((((1, True, False), (1, False, False), (second, False, False))
                                         if strategy == "cp_ulysses" else
                                         ((1, True, False), (1, False, False)) if strategy == "cp_ring_kv" else
                                         ((1, False, False),)) + off_variants):
            if GEMM and not fused:
                continue
            ref_m, cp_m = build("none", 1), build(strategy, world, offload)
            inf_ref = InferenceParams(1, 5 * clip)
            inf_cp = InferenceParams(1, 5 * clip)
            worst = {"core": 0.0, "xattn": 0.0}
            g = torch.Generator().manual_seed(5)
            for (sp, dn, kc, yi, upd, extra, ncc) in steps:
                S = dn * clip
                hs = torch.randn(S, 1, hidden, generator=g).to(torch.bfloat16).to(dev)
                kvs = torch.randn(sum(yi), hidden, generator=g).to(torch.bfloat16).to(dev)
                ang = torch.rand(S, 48, generator=g) * 6.28
                rope = torch.cat([torch.sin(ang), torch.cos(ang)], -1).to(dev)
                m_full = meta(sp, dn, kc, yi, extra, ncc)
                inf_ref.update_kv_cache = inf_cp.update_kv_cache = upd
                ref_core, ref_x = ref_m(hs, kvs, inf_ref, rope, m_full)
                cond = torch.zeros(S, 1, device=dev)
                cp_ = m_full.core_attn_params
                ardf = dict(denoising_range_num=dn, q_range=cp_.q_range, k_range=cp_.k_range, max_seqlen_q=clip,
                            max_seqlen_k=cp_.max_seqlen_k)
                x, _, rope_l, pad, split, core_l, cross_l = cpm.cp_pre_process(world, strategy, hs, cond, rope, None,
                                                                                ardf, cp_, m_full.cross_attn_params)
                m_cp = meta(sp, dn, kc, yi, extra, ncc, split=split, cross=cross_l)
                if strategy == "cp_shuffle_overlap":   # padded q ranges, per-chunk shards (context_parallel.py:283-296)
                    import dataclasses
                    m_cp = dataclasses.replace(m_cp, cp_pad_size=pad, core_attn_params=core_l)
                core, xo = cp_m(x, kvs, inf_cp, rope_l, m_cp)
                if strategy == "cp_ulysses" and fused and cp_m.supports_fused_proj_input(1, dev):
                    # fused_proj_input under the fused path: the same forward again (cache untouched), both attention
                    # kernels storing straight into the owner's out-projection input; must equal cat + rearrange
                    inf_cp.update_kv_cache = False
                    want = type(cp_m).attn_linear_proj_input(core, xo).clone()
                    pin = cp_m(x, kvs, inf_cp, rope_l, m_cp, fused_proj_input=True)
                    worst["proj"] = max(worst.get("proj", 0.0), (pin.float() - want.float()).abs().max().item())
                    inf_cp.update_kv_cache = upd
                core = cpm.cp_post_process(world, strategy, core.contiguous(), m_cp)
                xo = cpm.cp_post_process(world, strategy, xo.contiguous(), m_cp)
                worst["core"] = max(worst["core"], (core.float() - ref_core.float()).abs().max().item())
                worst["xattn"] = max(worst["xattn"], (xo.float() - ref_x.float()).abs().max().item())
            if strategy != "cp_shuffle_overlap":
                # batched uncond pass of the 3-CFG models (dit_model.py:413-494): chunks on the batch dim, one range
                # per batch row, no KV cache (fused variants: the batched fused path, rows stay "(sq b)"; else NCCL Ulysses)
                bs = 3
                hs = torch.randn(clip, bs, hidden, generator=g).to(torch.bfloat16).to(dev)
                yi = [50, 0, 77]
                kvs = torch.randn(sum(yi), hidden, generator=g).to(torch.bfloat16).to(dev)
                ang = torch.rand(clip, 48, generator=g) * 6.28
                rope = torch.cat([torch.sin(ang), torch.cos(ang)], -1).to(dev)
                edges = np.arange(bs + 1, dtype=np.int32) * clip
                rng_ = np.stack([edges[:-1], edges[1:]], 1)
                cu_k = np.concatenate([[0], np.cumsum(yi)]).astype(np.int32)
                core_u = PackedCoreAttnParams(i32(rng_), i32(rng_), rng_, rng_, clip, bs * clip)
                cross_u = PackedCrossAttnParams(i32(rng_), i32(np.stack([cu_k[:-1], cu_k[1:]], 1)), i32(edges), i32(cu_k),
                                                clip, 800)

                def meta_u(split, cross):
                    return ModelMetaArgs(H=1, W=1, cp_pad_size=0, cp_split_sizes=split, slice_point=0,
                                         denoising_range_num=1, range_num=1, extract_prefix_video_feature=False,
                                         fwd_extra_1st_chunk=False, distill_nearly_clean_chunk=False,
                                         clip_token_nums=clip, enable_cuda_graph=False, core_attn_params=core_u,
                                         cross_attn_params=cross)

                ref_core, ref_x = ref_m(hs, kvs, None, rope, meta_u(None, cross_u))
                ardf = dict(denoising_range_num=1, q_range=core_u.q_range, k_range=core_u.k_range, max_seqlen_q=clip,
                            max_seqlen_k=bs * clip)
                x, _, rope_l, pad, split, _, cross_l = cpm.cp_pre_process(
                    world, strategy, hs, torch.zeros(clip, bs, device=dev), rope, None, ardf, core_u, cross_u)
                m_cp = meta_u(split, cross_l)
                core, xo = cp_m(x, kvs, None, rope_l, m_cp)
                core = cpm.cp_post_process(world, strategy, core.contiguous(), m_cp)
                xo = cpm.cp_post_process(world, strategy, xo.contiguous(), m_cp)
                worst["core"] = max(worst["core"], (core.float() - ref_core.float()).abs().max().item())
                worst["xattn"] = max(worst["xattn"], (xo.float() - ref_x.float()).abs().max().item())
            results[f"{strategy}/{'fused-p2p' if fused else 'nccl'}/overlap{overlap}{'/kv_offload' if offload else ''}"
                    f"{'/fused-gemm' if GEMM else ''}{'/fp8_quant' if FP8 else ''}"] = worst
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        ok = True
        for k, w in results.items():
            bound = 2.5e-2 if k.startswith("cp_ring_kv/nccl") else 0.0   # the fused gather keeps every row's key order
            good = w["core"] <= bound and w["xattn"] == 0.0 and w.get("proj", 0.0) == 0.0
            ok &= good
            print(f"[cp_check N={world}] {k}: core max_abs {w['core']:.3e} (bound {bound}) xattn max_abs "
                  f"{w['xattn']:.3e}" + (f" fused_proj_input max_abs {w['proj']:.3e}" if "proj" in w else "")
                  + f" -> {'OK' if good else 'FAIL'}", flush=True)
        mcast = [any(w_.mc_ptrs) for w_ in cpm.RingGatherWorkspace._instances.values()]
        print(f"[cp_check N={world}] fused gather stores: "
              f"{'NVSwitch multicast' if (mcast and all(mcast) and cpm.ring_multicast_enabled()) else 'per-peer NVLink stores'}"
              f" (multicast mapping available: {bool(mcast and all(mcast))})", flush=True)
        print("CP_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    overlap, fused = 1, False
    main()
