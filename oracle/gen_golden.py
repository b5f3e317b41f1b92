"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN PYTHON CODE (imported from /root/reference) on CPU
in the authoring container. Test infrastructure only; /root/reference does not exist on the GPU box, so the
fixtures are committed and this script is only re-run by hand:

    python oracle/gen_golden.py            # writes tests/golden/

The reference's orchestration code runs unmodified (FullyParallelAttention.forward, get_q/get_k/get_kv/get_xqkv,
adjust_key_and_value_for_inference, core_attention's fallback loop, cross_attention, cp_ulysses_process,
cp_update_cross_attn_qkv_range, all_to_all_input/output_split, UlyssesScheduler, generate_sequences, the kv-range
generators). Only its third-party CUDA-only kernels are substituted by their published math so that it can run
without a GPU:
  * flash_attn.layers.rotary.apply_rotary_emb (Triton)  -> flash-attn's own `apply_rotary_emb_torch`
  * flash_attn_func / flash_attn_varlen_func (CUDA)     -> fp32 softmax(QK^T/sqrt(d))V with GQA, output cast to bf16
  * torch.cuda.current_device / Tensor.cuda / get_device_capability -> CPU stand-ins
  * flashinfer.gemm.bmm_fp8 (cuBLASLt fp8 matmul, fp8_quant mode only) -> (A.float() @ B.float()) * (a_scale[0] *
    b_scale[0]) cast to the output dtype: fp32 accumulation, ONE float read behind each scale pointer
  * range_mod_kernel_fwd (the reference's own Triton kernel, dit_module.py:201-241) -> the same row-gather multiply
    in torch; its host wrapper range_mod_triton and bias_modulate_add run unmodified (in the whole-layer scenario the
    wrapper's `assert x.is_cuda` forces restating the wrapper too: range_mod_cpu)
"""
import math
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

warnings.filterwarnings("ignore")
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ------------------------------------------------------------------ third-party stand-ins (published math)
def sdpa_math(q, k, v):
    """flash_attn_func semantics: (b, s, h, d), non-causal, softmax_scale = 1/sqrt(d), GQA. fp32 math, bf16 out."""
    b, sq, hq, d = q.shape
    hk = k.shape[2]
    g = hq // hk
    qf = q.float().permute(0, 2, 1, 3)
    kf = k.float().repeat_interleave(g, dim=2).permute(0, 2, 1, 3)
    vf = v.float().repeat_interleave(g, dim=2).permute(0, 2, 1, 3)
    s = torch.matmul(qf, kf.transpose(-1, -2)) / math.sqrt(d)
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, vf).permute(0, 2, 1, 3).to(q.dtype).contiguous()


def flash_attn_func_cpu(q, k, v, deterministic=False, **kw):
    return sdpa_math(q, k, v)


def flash_attn_varlen_func_cpu(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, deterministic=False,
                               **kw):
    out = torch.zeros_like(q)
    cq, ck = cu_seqlens_q.tolist(), cu_seqlens_k.tolist()
    for i in range(len(cq) - 1):
        if cq[i + 1] > cq[i] and ck[i + 1] > ck[i]:
            out[cq[i]:cq[i + 1]] = sdpa_math(q[None, cq[i]:cq[i + 1]], k[None, ck[i]:ck[i + 1]],
                                             v[None, ck[i]:ck[i + 1]])[0]
    return out


def install_patches():
    sys.path.insert(0, REF)
    torch.cuda.current_device = lambda: "cpu"
    torch.cuda.get_device_capability = lambda *a, **k: (8, 0)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import inference.model.dit.dit_module as dm
    from flash_attn.layers.rotary import apply_rotary_emb_torch
    dm.flash_apply_rotary_emb = lambda x, cos, sin: apply_rotary_emb_torch(x, cos, sin)
    dm.flash_attn_func = flash_attn_func_cpu
    dm.flash_attn_varlen_func = flash_attn_varlen_func_cpu
    dm.flex_attention = None
    return dm


def init_dist(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inference.infra.distributed import parallel_state as mpu
    mpu.initialize_model_parallel(cp_size=world, pp_size=1, nccl_communicator_config_path=None,
                                  distributed_timeout_minutes=10, order="tp-cp-pp-dp")
    return mpu


def make_configs(cp_size=1, cp_strategy="none", hidden=128, hq=4, hkv=2):
    from inference.common.config import EngineConfig, ModelConfig
    mc = ModelConfig(model_name="videodit_ardf", num_layers=4, hidden_size=hidden, ffn_hidden_size=4 * hidden,
                     num_attention_heads=hq, num_query_groups=hkv, kv_channels=128, layernorm_epsilon=1e-6,
                     apply_layernorm_1p=True, params_dtype=torch.bfloat16, xattn_cond_hidden_ratio=1.0)
    ec = EngineConfig(cp_size=cp_size, cp_strategy=cp_strategy, ulysses_overlap_degree=1)
    return mc, ec


def build_attention(dm, mc, ec, seed):
    """Reference module with explicit initialisation (its LayerNorm params are torch.empty, dit_module.py:352-353)
    and the precision promotion of dit_model.py:621-633 applied."""
    torch.manual_seed(seed)
    attn = dm.FullyParallelAttention(mc, ec, layer_number=1)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p_ in attn.named_parameters():
            if "layernorm" in name or "layer_norm" in name:
                p_.copy_((torch.randn(p_.shape, generator=g) * (0.2 if "layernorm" in name else 0.02)).to(p_.dtype))
                if name.endswith("layer_norm.weight"):
                    p_.add_(1.0)
            else:
                p_.copy_((torch.randn(p_.shape, generator=g) * (1.0 / math.sqrt(p_.shape[-1]))).to(p_.dtype))
    for name, sub in attn.named_modules():
        if "_xattn" in name:
            continue
        if "q_layernorm" in name or "k_layernorm" in name:
            sub.float()
    return attn.eval()


def pk(t):
    """bf16 tensors are stored as their raw 16-bit patterns (uint16) to keep the fixtures small; tests/golden_io.py
    turns them back into fp32 values. Everything else is stored as is."""
    if isinstance(t, torch.Tensor):
        if t.dtype == torch.bfloat16:
            return t.contiguous().view(torch.int16).numpy().view(np.uint16)
        return t.detach().numpy()
    return np.asarray(t)


def state_to_np(attn):
    return {"w." + k_: pk(v_) for k_, v_ in attn.state_dict().items()}


def meta(dm_common, clip, slice_point, dn, bs, k_range, y_index, fwd_extra=False, distill_ncc=False,
         cp_split_sizes=None, cross_override=None):
    from inference.common import ModelMetaArgs, PackedCoreAttnParams, PackedCrossAttnParams
    cu_q = torch.tensor([0] + [clip] * (dn * bs), dtype=torch.int64).cumsum(0).to(torch.int32)
    cu_k = torch.tensor([0] + list(y_index), dtype=torch.int64).cumsum(0).to(torch.int32)
    q_range = torch.stack([cu_q[:-1], cu_q[1:]], 1)
    k_range = torch.tensor(k_range, dtype=torch.int32)
    flat = torch.unique(k_range, sorted=True)
    core = PackedCoreAttnParams(q_range=q_range, k_range=k_range, np_q_range=q_range.numpy(),
                                np_k_range=k_range.numpy(), max_seqlen_q=clip,
                                max_seqlen_k=int(flat[-1] - flat[0]))
    cross = cross_override or PackedCrossAttnParams(
        q_ranges=q_range, kv_ranges=torch.stack([cu_k[:-1], cu_k[1:]], 1), cu_seqlens_q=cu_q, cu_seqlens_kv=cu_k,
        max_seqlen_q=clip, max_seqlen_kv=800)
    return ModelMetaArgs(H=0, W=0, cp_pad_size=0, cp_split_sizes=cp_split_sizes, slice_point=slice_point,
                         denoising_range_num=dn, range_num=slice_point + dn, extract_prefix_video_feature=False,
                         fwd_extra_1st_chunk=fwd_extra, distill_nearly_clean_chunk=distill_ncc, clip_token_nums=clip,
                         enable_cuda_graph=False, core_attn_params=core, cross_attn_params=cross)


def rope_for(dm, T_total, H, W, keep_T):
    """The reference's own table builder on CPU (dit_module.py:670-717 via the dit_model.py:163-166 recipe)."""
    bands = dm.freq_bands(16, temperature=10000.0, step=1, device="cpu")
    rescale = math.sqrt((H * W) / (16 * 16))
    sin, cos = dm.build_rotary_pos_embed([T_total, H, W], bands=bands, dim=128, in_pixels=False,
                                         ref_feat_shape=[T_total, H / rescale, W / rescale], device="cpu")
    rope = torch.cat([sin, cos], -1)
    return rope[-(keep_T * H * W):].contiguous()


# ------------------------------------------------------------------ scenario 1: single rank, KV-cache sequence
def gen_single_rank():
    dm = install_patches()
    init_dist(0, 1, 29541)
    from inference.common import InferenceParams
    mc, ec = make_configs()
    attn = build_attention(dm, mc, ec, seed=1234)
    H = W = 4
    frames = 3                    # latent frames per chunk -> clip = 48 tokens
    clip = frames * H * W
    hidden = mc.hidden_size
    g = torch.Generator().manual_seed(7)
    fx = dict(state_to_np(attn))
    fx["clip"] = np.array(clip)
    inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)

    def run(tag, slice_point, dn, k_chunks, y_index, update, fwd_extra=False, distill_ncc=False, bs=1, params=inf):
        S = dn * clip
        hs = torch.randn(S, bs, hidden, generator=g).to(torch.bfloat16)
        ytot = int(sum(y_index))
        kvs = torch.randn(ytot, hidden, generator=g).to(torch.bfloat16)
        rope = rope_for(dm, (slice_point + dn) * frames, H, W, dn * frames)
        k_range = [[s * clip, e * clip] for s, e in k_chunks]
        m = meta(dm, clip, slice_point, dn, bs, k_range, y_index, fwd_extra, distill_ncc)
        if params is not None:
            params.update_kv_cache = update
        with torch.no_grad():
            core, xattn = attn(hs, kvs, params, rope, m)
            mixed = attn.linear_qkv.forward_ln(hs)
            sin_emb, cos_emb = rope.tensor_split(2, -1)
            q = attn.get_q(mixed, cos_emb, sin_emb)
            kv = attn.get_kv(mixed, cos_emb, sin_emb)
            xq, xk, xv = attn.get_xqkv(mixed, kvs)
            lin = {n: getattr(attn.linear_qkv, "forward_" + n)(mixed) for n in ("q", "qx", "k", "v")}
        fx.update({
            f"{tag}.hidden_states": pk(hs), f"{tag}.key_value_states": pk(kvs),
            f"{tag}.rope": pk(rope), f"{tag}.k_range": np.array(k_range, dtype=np.int32),
            f"{tag}.q_range": m.core_attn_params.np_q_range, f"{tag}.y_index": np.array(y_index),
            f"{tag}.meta": np.array([slice_point, dn, bs, int(update), int(fwd_extra), int(distill_ncc),
                                     int(params is not None)]),
            f"{tag}.lin_q": pk(lin["q"]), f"{tag}.lin_qx": pk(lin["qx"]),
            f"{tag}.lin_k": pk(lin["k"]), f"{tag}.lin_v": pk(lin["v"]),
            f"{tag}.q": pk(q), f"{tag}.kv": pk(kv), f"{tag}.xq": pk(xq), f"{tag}.xk": pk(xk), f"{tag}.xv": pk(xv),
            f"{tag}.core_attn_out": pk(core), f"{tag}.xattn_out": pk(xattn),
        })
        if params is not None and 1 in params.key_value_memory_dict:
            fx[f"{tag}.cache_after"] = pk(params.key_value_memory_dict[1]).copy()

    # the reference allocates the cache with torch.empty on first use (dit_module.py:905-912): make it deterministic
    inf.key_value_memory_dict[1] = torch.zeros(4 * clip, 2, 256, dtype=torch.bfloat16)
    # step A: 2 chunks denoising from scratch (no cache involvement: slice_point == 0, no extra chunk)
    run("A", 0, 2, [(0, 1), (0, 2)], [50, 50], update=True)
    # step B: first chunk finished -> re-forward it clean and cache it (fwd_extra_1st_chunk), 2 more in flight
    run("B", 0, 3, [(0, 1), (0, 2), (1, 3)], [0, 37, 61], update=True, fwd_extra=True)
    # step C: window advanced: chunks 1,2 in flight reading cached chunk 0 (3-CFG pass A: update_kv_cache False)
    run("C", 1, 2, [(0, 2), (1, 3)], [33, 90], update=False)
    # step D: same state, update True, distill extra self-only range appended (dit_model.py:544-566)
    run("D", 1, 3, [(0, 2), (1, 3), (3, 4)], [33, 90, 33], update=True, distill_ncc=True)
    # step E: 3-CFG uncond pass: chunks on the batch dim, no inference params (dit_model.py:470-491)
    run("E", 0, 1, [(0, 1)], [50, 50], update=False, bs=2, params=None)
    np.savez_compressed(os.path.join(OUT, "attention_single_rank.npz"), **fx)
    print("single-rank fixture:", len(fx), "arrays")
    dist.destroy_process_group()


# ------------------------------------------------------------------ scenario 2: range / schedule generators
def gen_ranges():
    sys.path.insert(0, REF)
    for name in ("ffmpeg", "diffusers", "timm", "imageio", "cv2", "av"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    fx = {}
    # `inference.pipeline.__init__` pulls in ffmpeg/diffusers/timm (absent here); register the package without
    # running it and stub the one sibling module video_generate.py imports names from (prompt text handling only).
    import importlib.util
    import inference  # noqa: F401
    pkg = types.ModuleType("inference.pipeline")
    pkg.__path__ = [os.path.join(REF, "inference/pipeline")]
    sys.modules["inference.pipeline"] = pkg
    pp = types.ModuleType("inference.pipeline.prompt_process")
    pp.get_negative_special_token_keys = pp.get_special_token_keys = pp.pad_special_token = lambda *a, **k: None
    sys.modules["inference.pipeline.prompt_process"] = pp
    spec = importlib.util.spec_from_file_location("inference.pipeline.video_generate",
                                                  os.path.join(REF, "inference/pipeline/video_generate.py"))
    vg = importlib.util.module_from_spec(spec)
    sys.modules["inference.pipeline.video_generate"] = vg
    spec.loader.exec_module(vg)
    cases = [(8, 8, 0), (8, 4, 0), (8, 4, 2), (8, 1, 0), (4, 4, 0), (5, 4, 1), (1, 4, 0)]
    for i, (cn, ws, off) in enumerate(cases):
        cs, ce, ts, te = vg.generate_sequences(cn, ws, off)
        fx[f"seq{i}.args"] = np.array([cn, ws, off])
        fx[f"seq{i}.out"] = np.array([cs, ce, ts, te])
    # kv-range generators are methods of SampleTransport; call them unbound on a minimal stand-in
    ti = types.SimpleNamespace(num_steps=64)
    fake = types.SimpleNamespace(
        device="cpu", transport_inputs=[ti],
        get_batch_size_and_chunk_token_nums=lambda infer_idx: (fake_bs[0], 9360))
    fake_bs = [1]
    kcases = [(0, 4, [0, 16, 32, 48], 1), (2, 3, [64, 20, 3], 1), (1, 4, [64, 47, 31, 15], 1), (0, 2, [10, 0], 2),
              (3, 1, [63], 1)]
    for i, (sp, dn, steps, bs) in enumerate(kcases):
        fake_bs[0] = bs
        r = vg.SampleTransport.generate_noise2clean_kvrange(fake, 0, sp, dn, [5, 4, 3, 2], 1, steps)
        d = vg.SampleTransport.generate_default_kvrange(fake, 0, sp, dn)
        fx[f"kv{i}.args"] = np.array([sp, dn, bs] + steps)
        fx[f"kv{i}.noise2clean"] = r.numpy()
        fx[f"kv{i}.default"] = d.numpy()
    np.savez_compressed(os.path.join(OUT, "ranges.npz"), **fx)
    print("range fixture:", len(fx), "arrays")


# ------------------------------------------------------------------ scenario 3: 2- and 4-rank Ulysses on gloo
def _cp_worker(rank, world, port, out_path, hq, hkv):
    dm = install_patches()
    init_dist(rank, world, port)
    import inference.infra.parallelism.context_parallel as cpm
    from inference.common import PackedCrossAttnParams
    mc, ec = make_configs(cp_size=world, cp_strategy="cp_ulysses", hq=hq, hkv=hkv)
    attn = build_attention(dm, mc, ec, seed=99)   # same seed on all ranks -> replicated weights
    H = W = 4
    frames = 3
    clip = frames * H * W + 3 * 0
    dn, sp, bs = 3, 0, 1
    S = dn * clip + 0
    g = torch.Generator().manual_seed(5)
    hs_full = torch.randn(S, bs, mc.hidden_size, generator=g).to(torch.bfloat16)
    y_index = [41, 50, 77]
    kvs = torch.randn(sum(y_index), mc.hidden_size, generator=g).to(torch.bfloat16)
    rope_full = rope_for(dm, dn * frames, H, W, dn * frames)
    k_range = [[0, clip], [0, 2 * clip], [clip, 3 * clip]]
    m_full = meta(dm, clip, sp, dn, bs, k_range, y_index)
    cond_map = torch.arange(S).reshape(S, 1) // clip
    x, cmap, rope, split_sizes, cross = cpm.cp_ulysses_process(world, hs_full, cond_map, rope_full, None,
                                                               m_full.cross_attn_params)
    m = meta(dm, clip, sp, dn, bs, k_range, y_index, cp_split_sizes=split_sizes, cross_override=cross)
    with torch.no_grad():
        core, xattn = attn(x, kvs, None, rope, m)
        # raw movement primitives on a recognisable tensor
        probe = (torch.arange(x.shape[0] * hq * 4, dtype=torch.float32).reshape(x.shape[0], hq, 4) + 1000 * rank)
        probe = probe.to(torch.bfloat16).contiguous()
        a2a_in, h = cpm.all_to_all_input_split(probe, split_sizes)
        h.wait()
        a2a_out, h = cpm.all_to_all_output_split(a2a_in.contiguous(), split_sizes)
        h.wait()
        kvprobe = (torch.arange(x.shape[0] * hkv * 4, dtype=torch.float32).reshape(x.shape[0], hkv, 4) + 500 * rank)
        kvprobe = kvprobe.to(torch.bfloat16).contiguous()
        kv_in, h = cpm.all_to_all_input_split(kvprobe, split_sizes)
        h.wait()
        gathered = cpm.gather_from_context_parallel_region(core.float().contiguous(), split_sizes)
    fx = {
        "split_sizes": np.array(split_sizes), "x_local": pk(x), "rope_local": pk(rope),
        "cross.q_ranges": pk(cross.q_ranges), "cross.kv_ranges": pk(cross.kv_ranges),
        "cross.cu_seqlens_q": pk(cross.cu_seqlens_q), "cross.cu_seqlens_kv": pk(cross.cu_seqlens_kv),
        "cross.max_seqlen_q": np.array(cross.max_seqlen_q),
        "core_attn_out": pk(core), "xattn_out": pk(xattn),
        "probe": pk(probe), "a2a_in": pk(a2a_in), "a2a_out": pk(a2a_out),
        "kvprobe": pk(kvprobe), "kv_in": pk(kv_in), "gathered": pk(gathered.to(torch.bfloat16)),
    }
    if rank == 0:
        fx.update(state_to_np(attn))
        fx.update({"hidden_states": pk(hs_full), "key_value_states": pk(kvs),
                   "rope": pk(rope_full), "k_range": np.array(k_range, dtype=np.int32),
                   "y_index": np.array(y_index), "clip": np.array(clip), "heads": np.array([hq, hkv])})
    np.savez_compressed(out_path.format(rank=rank), **fx)
    dist.barrier()
    dist.destroy_process_group()


def gen_cp(world, port, hq, hkv, tag):
    path = os.path.join(OUT, f"ulysses_{tag}_rank{{rank}}.npz")
    mp.spawn(_cp_worker, args=(world, port, path, hq, hkv), nprocs=world, join=True)
    print(f"ulysses cp={world} fixture written ({tag})")


# ------------------------------------------------------------------ scenario 4: cp_shuffle_overlap (CSO) on gloo
def _cso_worker(rank, world, port, out_path):
    dm = install_patches()
    init_dist(rank, world, port)
    import inference.infra.parallelism.context_parallel as cpm
    mc, ec = make_configs(cp_size=world, cp_strategy="cp_shuffle_overlap", hq=4, hkv=2)
    attn = build_attention(dm, mc, ec, seed=77)
    H, W, frames = 3, 5, 3
    clip = frames * H * W                 # 45: odd, so every chunk is padded by one row for cp = 2
    dn, sp, bs = 3, 0, 1
    S = dn * clip
    g = torch.Generator().manual_seed(11)
    hs_full = torch.randn(S, bs, mc.hidden_size, generator=g).to(torch.bfloat16)
    y_index = [41, 50, 77]
    kvs = torch.randn(sum(y_index), mc.hidden_size, generator=g).to(torch.bfloat16)
    rope_full = rope_for(dm, dn * frames, H, W, dn * frames)
    k_range = [[0, clip], [0, 2 * clip], [clip, 3 * clip]]
    m_full = meta(dm, clip, sp, dn, bs, k_range, y_index)
    cond_map = torch.arange(S).reshape(S, 1) // clip
    cp_ = m_full.core_attn_params
    ardf_meta = dict(clip_token_nums=clip, slice_point=sp, range_num=sp + dn, denoising_range_num=dn,
                     q_range=cp_.q_range, k_range=cp_.k_range, max_seqlen_q=clip, max_seqlen_k=cp_.max_seqlen_k)
    (x, cmap, rope, pad, split, core_p, cross) = cpm.cp_pre_process(
        world, "cp_shuffle_overlap", hs_full, cond_map, rope_full, None, ardf_meta, cp_, m_full.cross_attn_params)
    from inference.common import ModelMetaArgs
    m = ModelMetaArgs(H=0, W=0, cp_pad_size=pad, cp_split_sizes=split, slice_point=sp, denoising_range_num=dn,
                      range_num=sp + dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=False,
                      distill_nearly_clean_chunk=False, clip_token_nums=clip, enable_cuda_graph=False,
                      core_attn_params=core_p, cross_attn_params=cross)
    with torch.no_grad():
        core, xattn = attn(x, kvs, None, rope, m)
        gathered = cpm.cp_post_process(world, "cp_shuffle_overlap", core.float().contiguous(), m)
    fx = {
        "split_sizes": np.array(split), "pad": np.array(pad), "x_local": pk(x), "rope_local": pk(rope),
        "cond_local": pk(cmap), "core.q_range": pk(core_p.q_range), "core.k_range": pk(core_p.k_range),
        "core.max_seqlen_q": np.array(core_p.max_seqlen_q),
        "cross.q_ranges": pk(cross.q_ranges), "cross.kv_ranges": pk(cross.kv_ranges),
        "cross.max_seqlen_q": np.array(cross.max_seqlen_q),
        "core_attn_out": pk(core), "xattn_out": pk(xattn), "gathered": pk(gathered.to(torch.bfloat16)),
    }
    if rank == 0:
        fx.update(state_to_np(attn))
        fx.update({"hidden_states": pk(hs_full), "key_value_states": pk(kvs), "rope": pk(rope_full),
                   "k_range": np.array(k_range, dtype=np.int32), "y_index": np.array(y_index), "clip": np.array(clip),
                   "heads": np.array([4, 2])})
    np.savez_compressed(out_path.format(rank=rank), **fx)
    dist.barrier()
    dist.destroy_process_group()


def gen_cso(world, port):
    path = os.path.join(OUT, f"cso_cp{world}_rank{{rank}}.npz")
    mp.spawn(_cso_worker, args=(world, port, path), nprocs=world, join=True)
    print(f"cso cp={world} fixture written")


# ------------------------------------------------------------------ scenario 5: adaLN gating chain (bias_modulate_add)
class _RangeModGrid:
    """Stand-in for `range_mod_kernel_fwd[(M,)](...)`: what the Triton program does for every row
    (dit_module.py:218-241): y[row, :] = x[row, :] * gatings[map[row], :]."""

    def __getitem__(self, grid):
        def run(x, MAP, gatings, y, M, N, sxm, sxn, sgm, sgn, sym, syn, BLOCK_SIZE=None):
            y.copy_(x * gatings[MAP.long()])
        return run


def gen_gating():
    dm = install_patches()
    dm.range_mod_kernel_fwd = _RangeModGrid()
    torch.Tensor.is_cuda = property(lambda self: True)      # the wrapper asserts x.is_cuda (dit_module.py:252-254)
    mc, _ = make_configs(hidden=512)
    fx = {}
    g = torch.Generator().manual_seed(2024)
    for tag, (s_, b_, dn, xdt) in {"attn_b1": (3 * 29, 1, 3, torch.float32), "mlp_b1": (3 * 29, 1, 3, torch.bfloat16),
                                  "attn_b2": (41, 2, 1, torch.float32)}.items():
        h = mc.hidden_size
        norm = dm.FusedLayerNorm(mc, h).float()
        with torch.no_grad():
            norm.weight.copy_(torch.randn(h, generator=g) * 0.1)
            norm.bias.copy_(torch.randn(h, generator=g) * 0.1)
        x = (torch.randn(s_, b_, h, generator=g) * 1.5).to(xdt)
        residual = torch.randn(s_, b_, h, generator=g).to(torch.bfloat16)
        cond = torch.randn(b_, dn, h, generator=g)
        gate = dm.softcap(cond.to(torch.bfloat16), 1.0)                       # dit_module.py:1353-1355
        cmap = torch.arange(b_ * dn).repeat_interleave(s_ // dn).reshape(b_, -1).transpose(0, 1).contiguous()
        with torch.no_grad():
            out = dm.bias_modulate_add(x, residual, cmap, gate, norm)
        assert out.dtype == xdt
        fx.update({f"{tag}.x": pk(x) if xdt == torch.bfloat16 else x.numpy(), f"{tag}.residual": pk(residual),
                   f"{tag}.cmap": cmap.numpy(), f"{tag}.gate": pk(gate), f"{tag}.w": norm.weight.detach().numpy(),
                   f"{tag}.b": norm.bias.detach().numpy(), f"{tag}.cond": cond.numpy(),
                   f"{tag}.out": pk(out.to(torch.bfloat16)), f"{tag}.out_is_fp32": np.array(xdt == torch.float32)})
    del torch.Tensor.is_cuda
    np.savez_compressed(os.path.join(OUT, "gating.npz"), **fx)
    print("gating fixture written")


# ------------------------------------------------------------------ scenario 6: a whole TransformerLayer
def range_mod_cpu(x, c_mapping, gatings):
    """range_mod_triton (dit_module.py:244-289) with its Triton program (:201-241) restated: the host wrapper's
    transposes / flattens kept, the per-row program replaced by the gather-multiply it performs."""
    s, b, h = x.shape
    x2 = x.transpose(0, 1).flatten(0, 1)
    cm = c_mapping.transpose(0, 1).flatten(0, 1)
    y = x2 * gatings.flatten(0, 1)[cm.long()]
    return y.reshape(b, s, h).transpose(0, 1)


def gen_layer():
    dm = install_patches()
    dm.range_mod_triton = range_mod_cpu
    init_dist(0, 1, 29581)
    from inference.common import InferenceParams
    mc, ec = make_configs(hidden=256, hq=4, hkv=2)
    torch.manual_seed(4321)
    layer = dm.TransformerLayer(mc, ec, layer_number=0)
    g = torch.Generator().manual_seed(4321)
    with torch.no_grad():
        for name, p_ in layer.named_parameters():
            if "layernorm" in name or "post_norm" in name:
                p_.copy_((torch.randn(p_.shape, generator=g) * 0.2).to(p_.dtype))
            elif name.endswith("layer_norm.weight"):
                p_.copy_((1.0 + torch.randn(p_.shape, generator=g) * 0.02).to(p_.dtype))
            elif name.endswith("bias"):
                p_.copy_((torch.randn(p_.shape, generator=g) * 0.02).to(p_.dtype))
            else:
                p_.copy_((torch.randn(p_.shape, generator=g) * (1.0 / math.sqrt(p_.shape[-1]))).to(p_.dtype))
    for name, sub in layer.named_modules():       # dit_model.py:621-638
        if "_xattn" in name:
            continue
        if any(k in name for k in ("q_layernorm", "k_layernorm", "self_attn_post_norm", "mlp_post_norm")):
            sub.float()
    layer.eval()
    H = W = 4
    frames = 3
    clip = frames * H * W
    hidden = mc.hidden_size
    fx = {"w." + k_: pk(v_) for k_, v_ in layer.state_dict().items()}
    fx["clip"] = np.array(clip)
    fx["heads"] = np.array([4, 2])
    inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
    inf.key_value_memory_dict[0] = torch.zeros(4 * clip, 2, 256, dtype=torch.bfloat16)
    for tag, sp, dn, k_chunks, y_index, update in (("A", 0, 2, [(0, 1), (0, 2)], [50, 41], True),
                                                   ("B", 1, 2, [(0, 2), (1, 3)], [33, 90], True)):
        S = dn * clip
        hs = torch.randn(S, 1, hidden, generator=g).to(torch.bfloat16)
        kvs = torch.randn(int(sum(y_index)), hidden, generator=g).to(torch.bfloat16)
        cond = torch.randn(1, dn, int(hidden * mc.cond_hidden_ratio), generator=g).to(torch.bfloat16)
        cmap = torch.arange(dn).repeat_interleave(clip).reshape(1, -1).transpose(0, 1).contiguous()
        rope = rope_for(dm, (sp + dn) * frames, H, W, dn * frames)
        m = meta(dm, clip, sp, dn, 1, [[s * clip, e * clip] for s, e in k_chunks], y_index)
        inf.update_kv_cache = update
        with torch.no_grad():
            out = layer(hs, cond, cmap, kvs, rope, inf, m)
        fx.update({f"{tag}.hidden_states": pk(hs), f"{tag}.key_value_states": pk(kvs), f"{tag}.condition": pk(cond),
                   f"{tag}.condition_map": cmap.numpy(), f"{tag}.rope": pk(rope),
                   f"{tag}.k_range": np.array([[s * clip, e * clip] for s, e in k_chunks], dtype=np.int32),
                   f"{tag}.y_index": np.array(y_index), f"{tag}.meta": np.array([sp, dn, int(update)]),
                   f"{tag}.out": pk(out)})
    np.savez_compressed(os.path.join(OUT, "transformer_layer.npz"), **fx)
    print("transformer layer fixture written")


def gen_two_blocks():
    """BASELINE.json configs[0] in miniature: the DiT cut to 2 blocks. Two reference TransformerLayers (layer numbers
    0 and 1, own weights, own KV caches) chained through a three-forward KV-cache sequence; the fixture holds the
    hidden state after both blocks (what the final norm + projection turn into the step's denoised latent)."""
    dm = install_patches()
    dm.range_mod_triton = range_mod_cpu
    init_dist(0, 1, 29591)
    from inference.common import InferenceParams
    mc, ec = make_configs(hidden=256, hq=4, hkv=2)
    layers = []
    for ln_ in (0, 1):
        torch.manual_seed(777 + ln_)
        layer = dm.TransformerLayer(mc, ec, layer_number=ln_)
        g = torch.Generator().manual_seed(777 + ln_)
        with torch.no_grad():
            for name, p_ in layer.named_parameters():
                if "layernorm" in name or "post_norm" in name:
                    p_.copy_((torch.randn(p_.shape, generator=g) * 0.2).to(p_.dtype))
                elif name.endswith("layer_norm.weight"):
                    p_.copy_((1.0 + torch.randn(p_.shape, generator=g) * 0.02).to(p_.dtype))
                elif name.endswith("bias"):
                    p_.copy_((torch.randn(p_.shape, generator=g) * 0.02).to(p_.dtype))
                else:
                    p_.copy_((torch.randn(p_.shape, generator=g) * (1.0 / math.sqrt(p_.shape[-1]))).to(p_.dtype))
        for name, sub in layer.named_modules():       # dit_model.py:621-638
            if "_xattn" in name:
                continue
            if any(k in name for k in ("q_layernorm", "k_layernorm", "self_attn_post_norm", "mlp_post_norm")):
                sub.float()
        layers.append(layer.eval())
    H = W = 4
    frames = 3
    clip = frames * H * W
    hidden = mc.hidden_size
    fx = {}
    for ln_, layer in enumerate(layers):
        fx.update({f"w{ln_}." + k_: pk(v_) for k_, v_ in layer.state_dict().items()})
    fx["clip"] = np.array(clip)
    fx["heads"] = np.array([4, 2])
    inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
    g = torch.Generator().manual_seed(778899)
    # A: first stage step with the extra clean-chunk forward (writes chunks 0, 1 of both caches);
    # B: window advanced, two chunks in flight behind cached chunk 0 (reads what A wrote, overwrites slots 1, 2);
    # C: same state, update_kv_cache False (3-CFG first pass): reads chunk 0, leaves the caches alone
    for tag, sp, dn, k_chunks, y_index, update, extra in (
            ("A", 0, 2, [(0, 1), (0, 2)], [50, 41], True, True),
            ("B", 1, 2, [(0, 2), (0, 3)], [33, 90], True, False),
            ("C", 1, 2, [(0, 2), (1, 3)], [0, 64], False, False)):
        S = dn * clip
        hs = torch.randn(S, 1, hidden, generator=g).to(torch.bfloat16)
        kvs = torch.randn(int(sum(y_index)), hidden, generator=g).to(torch.bfloat16)
        cond = torch.randn(1, dn, int(hidden * mc.cond_hidden_ratio), generator=g).to(torch.bfloat16)
        cmap = torch.arange(dn).repeat_interleave(clip).reshape(1, -1).transpose(0, 1).contiguous()
        rope = rope_for(dm, (sp + dn) * frames, H, W, dn * frames)
        m = meta(dm, clip, sp, dn, 1, [[s_ * clip, e_ * clip] for s_, e_ in k_chunks], y_index, fwd_extra=extra)
        inf.update_kv_cache = update
        x = hs
        with torch.no_grad():
            for layer in layers:
                x = layer(x, cond, cmap, kvs, rope, inf, m)
        fx.update({f"{tag}.hidden_states": pk(hs), f"{tag}.key_value_states": pk(kvs), f"{tag}.condition": pk(cond),
                   f"{tag}.condition_map": cmap.numpy(), f"{tag}.rope": pk(rope),
                   f"{tag}.k_range": np.array([[s_ * clip, e_ * clip] for s_, e_ in k_chunks], dtype=np.int32),
                   f"{tag}.y_index": np.array(y_index), f"{tag}.meta": np.array([sp, dn, int(update), int(extra)]),
                   f"{tag}.out": pk(x)})
    for ln_ in (0, 1):
        fx[f"cache{ln_}"] = pk(inf.key_value_memory_dict[ln_][:3 * clip])   # rows beyond are torch.empty garbage
    np.savez_compressed(os.path.join(OUT, "dit_two_blocks.npz"), **fx)
    print("two-block fixture written")


# ------------------------------------------------------------------ scenario: fp8_quant linears
def bmm_fp8_cpu(A, B, A_scale, B_scale, dtype, out=None, backend="cublas"):
    """flashinfer.gemm.bmm_fp8 -> cuBLASLt: D = dtype((A @ B) * a_scale * b_scale), per-tensor scalars."""
    acc = torch.matmul(A.float(), B.float())
    return (acc * (A_scale.flatten()[0].float() * B_scale.flatten()[0].float())).to(dtype)


def init_fp8_linear(lin, g, input_amax=4.0, vector_jitter=0.0):
    """Fill a reference fp8 linear the way a calibrated checkpoint would: e4m3 weights with a per-tensor weight_scale,
    input_scale = amax / 448 (a vector of equal values for the per-tensor class, dit_module.py:443), smooth_scale
    around 1. vector_jitter > 0 makes the per-channel divisors differ (they are divisors in div_clamp_to, :447)."""
    out_f, in_f = lin.weight.shape[1:]
    w = torch.randn(out_f, in_f, generator=g) / math.sqrt(in_f)
    ws = w.abs().max() / 448.0
    with torch.no_grad():
        lin.weight.copy_((w / ws).to(torch.float8_e4m3fn)[None])
        lin.weight_scale.fill_(float(ws))
        base = input_amax / 448.0
        if hasattr(lin, "smooth_scale"):
            lin.smooth_scale.copy_(1.0 + 0.25 * torch.rand(1, in_f, generator=g))
            lin.input_scale.fill_(base)
        else:
            lin.input_scale.copy_(base * (1.0 + vector_jitter * torch.rand(in_f, generator=g)))


def gen_fp8():
    dm = install_patches()
    dm.bmm_fp8 = bmm_fp8_cpu
    dm.range_mod_triton = range_mod_cpu
    init_dist(0, 1, 29591)
    from inference.common import InferenceParams
    from inference.common.config import EngineConfig
    g = torch.Generator().manual_seed(2024)
    fx = {}
    # (1) div_clamp_to on its own: per-channel and scalar divisors, values beyond the clamp, zeros, subnormal results
    x = (torch.randn(96, 256, generator=g) * 3.0)
    x[0, :8] = torch.tensor([0.0, -0.0, 1e-4, -1e-4, 500.0, -500.0, 3.9, 448.0])
    x = x.to(torch.bfloat16)
    sv = (4.0 / 448.0) * (1.0 + torch.rand(256, generator=g))
    ss = torch.tensor([0.011])
    fx.update({"dct.x": pk(x), "dct.scale_vec": sv.numpy(), "dct.scale_scalar": ss.numpy(),
               "dct.out_vec": dm.div_clamp_to(x, sv).view(torch.uint8).numpy(),
               "dct.out_scalar": dm.div_clamp_to(x, ss).view(torch.uint8).numpy()})
    # (2) the two linear classes on their own (per-tensor one with unequal per-channel divisors)
    pt = dm.PerTensorQuantizedFp8Linear(256, 384)
    pc = dm.PerChannelQuantizedFp8Linear(256, 128)
    init_fp8_linear(pt, g, vector_jitter=0.5)
    init_fp8_linear(pc, g)
    xin = torch.randn(40, 2, 256, generator=g).to(torch.bfloat16)
    with torch.no_grad():
        fx.update({"lin.x": pk(xin), "lin.pt.out": pk(pt(xin)), "lin.pc.out": pk(pc(xin))})
    for n_, m_ in (("pt", pt), ("pc", pc)):
        for k_, v_ in m_.state_dict().items():
            fx[f"lin.{n_}.{k_}"] = v_.view(torch.uint8).numpy() if v_.dtype == torch.float8_e4m3fn else pk(v_)
    # (3) a whole fp8 TransformerLayer (layer 1 of 4: every linear the mode covers is fp8, dit_module.py:407, 522, 535,
    # 875) through a two-step KV-cache sequence; `shared` = q / k / v / qx calibrated on their common input (equal
    # input_scale vectors), `split` = unequal ones
    mc, _ = make_configs(hidden=256, hq=4, hkv=2)
    ec = EngineConfig(cp_size=1, cp_strategy="none", ulysses_overlap_degree=1, fp8_quant=True)
    H = W = 4
    frames = 3
    clip = frames * H * W
    hidden = mc.hidden_size
    for variant in ("shared", "split"):
        torch.manual_seed(99)
        layer = dm.TransformerLayer(mc, ec, layer_number=1)
        with torch.no_grad():
            for name, p_ in layer.named_parameters():
                if p_.dtype == torch.float8_e4m3fn or name.endswith("_scale"):
                    continue
                if "layernorm" in name or "post_norm" in name:
                    p_.copy_((torch.randn(p_.shape, generator=g) * 0.2).to(p_.dtype))
                elif name.endswith("layer_norm.weight"):
                    p_.copy_((1.0 + torch.randn(p_.shape, generator=g) * 0.02).to(p_.dtype))
                elif name.endswith("bias"):
                    p_.copy_((torch.randn(p_.shape, generator=g) * 0.02).to(p_.dtype))
                else:
                    p_.copy_((torch.randn(p_.shape, generator=g) * (1.0 / math.sqrt(p_.shape[-1]))).to(p_.dtype))
        for name, sub in layer.named_modules():
            if isinstance(sub, (dm.PerTensorQuantizedFp8Linear, dm.PerChannelQuantizedFp8Linear)):
                amax = {"shared": 4.0, "split": 4.0 + 0.5 * len(name)}[variant] if "linear_qkv" in name else 6.0
                init_fp8_linear(sub, g, input_amax=amax)
        for name, sub in layer.named_modules():       # dit_model.py:621-638
            if "_xattn" in name:
                continue
            if any(k in name for k in ("q_layernorm", "k_layernorm", "self_attn_post_norm", "mlp_post_norm")):
                sub.float()
        layer.eval()
        pre = variant + "."
        for k_, v_ in layer.state_dict().items():
            fx[pre + "w." + k_] = v_.view(torch.uint8).numpy() if v_.dtype == torch.float8_e4m3fn else pk(v_)
        inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
        inf.key_value_memory_dict[1] = torch.zeros(4 * clip, 2, 256, dtype=torch.bfloat16)
        for tag, sp, dn, k_chunks, y_index, update in (("A", 0, 2, [(0, 1), (0, 2)], [50, 41], True),
                                                       ("B", 1, 2, [(0, 2), (1, 3)], [33, 90], True)):
            S = dn * clip
            hs = torch.randn(S, 1, hidden, generator=g).to(torch.bfloat16)
            kvs = torch.randn(int(sum(y_index)), hidden, generator=g).to(torch.bfloat16)
            cond = torch.randn(1, dn, int(hidden * mc.cond_hidden_ratio), generator=g).to(torch.bfloat16)
            cmap = torch.arange(dn).repeat_interleave(clip).reshape(1, -1).transpose(0, 1).contiguous()
            rope = rope_for(dm, (sp + dn) * frames, H, W, dn * frames)
            m = meta(dm, clip, sp, dn, 1, [[s_ * clip, e_ * clip] for s_, e_ in k_chunks], y_index)
            inf.update_kv_cache = update
            with torch.no_grad():
                core, xattn = layer.self_attention(hs, kvs, inf, rope, m)
                # the attention call above already appended this step's K/V: the layer call repeats the same write
                out = layer(hs, cond, cmap, kvs, rope, inf, m)
            t = pre + tag
            fx.update({f"{t}.hidden_states": pk(hs), f"{t}.key_value_states": pk(kvs), f"{t}.condition": pk(cond),
                       f"{t}.condition_map": cmap.numpy(), f"{t}.rope": pk(rope),
                       f"{t}.k_range": np.array([[s_ * clip, e_ * clip] for s_, e_ in k_chunks], dtype=np.int32),
                       f"{t}.y_index": np.array(y_index), f"{t}.meta": np.array([sp, dn, int(update)]),
                       f"{t}.core": pk(core), f"{t}.xattn": pk(xattn), f"{t}.out": pk(out)})
        fx[pre + "cache"] = pk(inf.key_value_memory_dict[1][:3 * clip])
    fx["clip"] = np.array(clip)
    fx["heads"] = np.array([4, 2])
    np.savez_compressed(os.path.join(OUT, "fp8_quant.npz"), **fx)
    print("fp8 fixture written")


# ------------------------------------------------------------------ scenario: the block stack (TransformerBlock)
def gen_block():
    """The caller of the layers: the reference's own TransformerBlock (dit_module.py:1373-1441) with three layers and
    fp8_quant on - layers 0 and 2 keep bf16 linears, layer 1 holds the fp8 ones (:407, 522, 535, 875) - and the final
    fp32 LayerNorm, through a two-forward KV-cache sequence."""
    dm = install_patches()
    dm.bmm_fp8 = bmm_fp8_cpu
    dm.range_mod_triton = range_mod_cpu
    init_dist(0, 1, 29593)
    from inference.common import InferenceParams
    from inference.common.config import EngineConfig, ModelConfig
    # hidden 256: the fused gating kernel works on 256-wide groups (every real model is a multiple of it)
    mc = ModelConfig(model_name="videodit_ardf", num_layers=3, hidden_size=256, ffn_hidden_size=512,
                     num_attention_heads=2, num_query_groups=2, kv_channels=128, layernorm_epsilon=1e-6,
                     apply_layernorm_1p=True, params_dtype=torch.bfloat16, xattn_cond_hidden_ratio=1.0)
    ec = EngineConfig(cp_size=1, cp_strategy="none", ulysses_overlap_degree=1, fp8_quant=True)
    torch.manual_seed(31)
    block = dm.TransformerBlock(mc, ec)
    g = torch.Generator().manual_seed(31)
    with torch.no_grad():
        for name, p_ in block.named_parameters():
            if p_.dtype == torch.float8_e4m3fn or name.endswith("_scale"):
                continue
            if "layernorm" in name or "post_norm" in name:
                p_.copy_((torch.randn(p_.shape, generator=g) * 0.2).to(p_.dtype))
            elif name.endswith("layer_norm.weight"):
                p_.copy_((1.0 + torch.randn(p_.shape, generator=g) * 0.02).to(p_.dtype))
            elif name.endswith("bias"):
                p_.copy_((torch.randn(p_.shape, generator=g) * 0.02).to(p_.dtype))
            else:
                p_.copy_((torch.randn(p_.shape, generator=g) * (1.0 / math.sqrt(p_.shape[-1]))).to(p_.dtype))
    for name, sub in block.named_modules():
        if isinstance(sub, (dm.PerTensorQuantizedFp8Linear, dm.PerChannelQuantizedFp8Linear)):
            init_fp8_linear(sub, g, input_amax=4.0 if "linear_qkv" in name else 6.0)
    for name, sub in block.named_modules():       # dit_model.py:621-638
        if "_xattn" in name:
            continue
        if any(k in name for k in ("q_layernorm", "k_layernorm", "self_attn_post_norm", "mlp_post_norm",
                                   "final_layernorm")):
            sub.float()
    block.eval()
    assert isinstance(block.layers[1].mlp.linear_fc1, dm.PerTensorQuantizedFp8Linear)
    assert isinstance(block.layers[0].mlp.linear_fc1, torch.nn.Linear)
    assert isinstance(block.layers[2].mlp.linear_fc1, torch.nn.Linear)
    H = W = 4
    frames = 3
    clip = frames * H * W
    hidden = mc.hidden_size
    fx = {"w." + k_: (v_.view(torch.uint8).numpy() if v_.dtype == torch.float8_e4m3fn else pk(v_))
          for k_, v_ in block.state_dict().items()}
    fx["clip"] = np.array(clip)
    fx["dims"] = np.array([3, hidden, 512, 2, 2])       # layers, hidden, ffn, Hq, Hkv
    inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
    for tag, sp, dn, k_chunks, y_index, update, extra in (
            ("A", 0, 2, [(0, 1), (0, 2)], [50, 41], True, True),
            ("B", 1, 2, [(0, 2), (1, 3)], [0, 64], True, False)):
        S = dn * clip
        hs = torch.randn(S, 1, hidden, generator=g).to(torch.bfloat16)
        kvs = torch.randn(int(sum(y_index)), hidden, generator=g).to(torch.bfloat16)
        cond = torch.randn(1, dn, int(hidden * mc.cond_hidden_ratio), generator=g).to(torch.bfloat16)
        cmap = torch.arange(dn).repeat_interleave(clip).reshape(1, -1).transpose(0, 1).contiguous()
        rope = rope_for(dm, (sp + dn) * frames, H, W, dn * frames)
        m = meta(dm, clip, sp, dn, 1, [[s_ * clip, e_ * clip] for s_, e_ in k_chunks], y_index, fwd_extra=extra)
        inf.update_kv_cache = update
        per_layer = []
        hooks = [l_.register_forward_hook(lambda mod, a, o: per_layer.append(o.clone())) for l_ in block.layers]
        with torch.no_grad():
            out = block(hs, cond, cmap, kvs, rope, inf, m)
        for h_ in hooks:
            h_.remove()
        assert out.dtype == torch.float32 and len(per_layer) == 3
        fx.update({f"{tag}.after_layer{i}": pk(t_) for i, t_ in enumerate(per_layer)})
        fx.update({f"{tag}.hidden_states": pk(hs), f"{tag}.key_value_states": pk(kvs), f"{tag}.condition": pk(cond),
                   f"{tag}.condition_map": cmap.numpy(), f"{tag}.rope": pk(rope),
                   f"{tag}.k_range": np.array([[s_ * clip, e_ * clip] for s_, e_ in k_chunks], dtype=np.int32),
                   f"{tag}.y_index": np.array(y_index), f"{tag}.meta": np.array([sp, dn, int(update), int(extra)]),
                   f"{tag}.out": out.numpy()})
    np.savez_compressed(os.path.join(OUT, "transformer_block.npz"), **fx)
    print("transformer block fixture written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "ranges"):
        gen_ranges()
    if which in ("all", "cp2"):
        gen_cp(2, 29551, 4, 2, "cp2")
    if which in ("all", "cp4"):
        gen_cp(4, 29561, 8, 2, "cp4")   # cp > Hkv: exercises KV-head replication (context_parallel.py:393-394)
    if which in ("all", "cso"):
        gen_cso(2, 29571)
    if which in ("all", "layer"):
        gen_layer()
    if which in ("two_blocks",):  # run on its own (a process group per scenario)
        gen_two_blocks()
    if which in ("fp8",):         # run on its own (a process group per scenario)
        gen_fp8()
    if which in ("block",):       # run on its own
        gen_block()
    if which in ("gating",):      # run on its own: it monkey-patches Tensor.is_cuda while it runs
        gen_gating()
    if which in ("all", "single"):
        gen_single_rank()
