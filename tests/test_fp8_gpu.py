"""GPU parity tests (B200) of the fp8_quant mode (engine_config.fp8_quant, dit_module.py:364-384, 405-410, 431-488, 520-540,
873-880): the activation quantiser, the fp8 instantiation of the fused projection GEMM and the whole fp8 DiT block,
called through the C-ABI, against the CPU oracle, the reference-generated fixture tests/golden/fp8_quant.npz and the
library GEMM the reference itself calls (cuBLASLt fp8 matmul).

Tolerances
  quantiser                         : byte exact (integer / byte work)
  fp8 GEMM + prologue vs oracle     : as the bf16 fused GEMM (max_abs 6e-2, rel-L2 6e-3): fp8 products are exact in fp32,
                                      the tensor core's summation order moves a few bf16 roundings of the GEMM output,
                                      which the LayerNorm's 1 / sigma amplifies on a handful of elements
  whole block vs reference fixture  : rel-L2 <= 2e-2, the per-step latent criterion of SURVEY.md section 8(c)
"""
import numpy as np
import pytest
import torch

import magi_oracle as O
from golden_io import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import magi_b200.functional as F_
    return F_


def bf(x, dev="cuda"):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev).to(torch.bfloat16)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).float().numpy()


def fp8_bytes(t):
    return t.view(torch.uint8).cpu().numpy()


def test_fp8_quantize_matches_reference_fixture(F):
    """magi_fp8_quantize on the inputs the reference's own div_clamp_to ran on (per-channel and scalar divisors, values
    beyond the clamp, signed zeros, subnormal results): byte exact."""
    g = Golden("fp8_quant.npz")
    x = g.torch("dct.x", "cuda")
    for kind in ("vec", "scalar"):
        out = F.fp8_quantize(x, torch.from_numpy(g.raw(f"dct.scale_{kind}")).cuda())
        assert out.dtype == torch.float8_e4m3fn and out.shape == x.shape
        assert np.array_equal(fp8_bytes(out), g.raw(f"dct.out_{kind}")), kind


@pytest.mark.parametrize("rows,cols,vector", [(1, 16, True), (777, 3072, True), (1000, 6144, False), (37, 48, True),
                                              (0, 256, False)])
def test_fp8_quantize_vs_oracle_and_eager(F, rows, cols, vector):
    """... and on seeded inputs against the oracle and against the reference's eager chain run on the same GPU (the
    literal lines of div_clamp_to): byte exact; a 3-D input keeps its shape; rows outside are untouched."""
    x = rnd((rows, cols), 11, scale=2.5)
    if rows:
        x[0, :4] = [0.0, 1e-5, 1000.0, -1000.0]
    sc = (np.random.default_rng(3).uniform(0.004, 0.02, size=cols if vector else 1)).astype(np.float32)
    xt, st = bf(x), torch.from_numpy(sc).cuda()
    out = F.fp8_quantize(xt, st)
    assert np.array_equal(fp8_bytes(out), O.e4m3_encode(O.div_clamp_to(x, sc)))
    eager = torch.clamp(xt.float() / st.float(), -448.0, 448.0).bfloat16().to(torch.float8_e4m3fn)
    assert torch.equal(out.view(torch.uint8), eager.view(torch.uint8))
    if rows % 3 == 0 and rows:
        out3 = F.fp8_quantize(xt.view(3, rows // 3, cols), st.view(1, -1) if vector else st)
        assert out3.shape == (3, rows // 3, cols) and torch.equal(out3.view(torch.uint8).view(rows, cols),
                                                                  out.view(torch.uint8))


def test_fp8_quantize_exhaustive_over_bf16_inputs(F):
    """Every bf16 bit pattern (NaNs aside) against 1024 divisors - random ones over the whole range the fast path accepts,
    'nice' ones whose quotients sit exactly on bf16 rounding ties (0.1, 3, 4 / 448, powers of two), negative ones and
    ones far outside the fast path's range (subnormal, 1e38) - byte exact against the reference's eager chain on the
    same GPU: 67 million divisions. This is the test of the kernel's arithmetic shortcut (reciprocal + two fma
    corrections instead of the IEEE division, csrc/fp8_quant.cu; tests/test_fp8_quant_algorithm.py restates it on the
    CPU)."""
    ncol = 1024
    bits = torch.arange(65536, dtype=torch.int32, device="cuda").to(torch.int16)
    x = bits.view(torch.bfloat16).view(-1, 1).expand(-1, ncol).contiguous()
    rng = np.random.default_rng(7)
    sc = np.concatenate([rng.uniform(0.004, 0.02, 500), rng.uniform(0.5, 2.0, 100), 2.0 ** rng.uniform(-60, 60, 400),
                         [1.0, 0.0078125, 3.0, 0.1, 7.0, 448.0, 0.3, 0.75, 1e-30, 1e30, 1e-42, 1e38, 3e38, 2.0 ** -126,
                          2.0 ** 126, 6.0, 4.0 / 448.0, -0.01, -3.0, 2.0 ** -60, 2.0 ** 60, 2.0 ** -61, 2.0 ** 61,
                          0.0123]]).astype(np.float32)
    assert sc.size == ncol
    st = torch.from_numpy(sc).cuda()
    out = F.fp8_quantize(x, st).view(torch.uint8)
    eager = torch.clamp(x.float() / st.float(), -448.0, 448.0).bfloat16().to(torch.float8_e4m3fn).view(torch.uint8)
    ok = ~torch.isnan(x.float())
    assert torch.equal(out[ok], eager[ok])
    sub = slice(0, None, 16)       # the oracle on every 16th column (numpy, CPU)
    assert np.array_equal(out[:, sub][ok[:, sub]].cpu().numpy(),
                          O.e4m3_encode(O.div_clamp_to(x[:, sub].float().cpu().numpy(), sc[sub]))[ok[:, sub].cpu().numpy()])
    # scalar divisor: the same inputs against each of a few divisors
    for s1 in (0.0123, 0.1, 1e-42):
        s1t = torch.tensor([s1], dtype=torch.float32, device="cuda")
        o1 = F.fp8_quantize(x[:, :16].contiguous(), s1t).view(torch.uint8)
        e1 = torch.clamp(x[:, :16].float() / s1t, -448.0, 448.0).bfloat16().to(torch.float8_e4m3fn).view(torch.uint8)
        assert torch.equal(o1[ok[:, :16]], e1[ok[:, :16]]), s1


def _fp8_weights(shape, seed):
    w = rnd(shape, seed, scale=shape[1] ** -0.5)
    ws = np.float32(np.abs(w).max() / 448.0)
    wq = O.e4m3_round(np.clip(w / ws, -448, 448))
    return wq, ws


@pytest.mark.parametrize("rows,hidden,Hq,Hkv,with_qx,batch", [
    (300, 256, 4, 2, True, 1), (1000, 3072, 6, 2, False, 1), (514, 512, 2, 2, True, 2), (4096, 3072, 24, 8, True, 1)])
def test_fp8_projection_gemm_vs_oracle(F, rows, hidden, Hq, Hkv, with_qx, batch):
    """The fp8 instantiation of magi_qkv_proj_norm_rope_fwd (tcgen05 kind::f8f6f4, accumulator scaled by
    input_scale * weight_scale per projection) against the oracle's composition of the reference steps -
    PerTensorQuantizedFp8Linear.forward (dit_module.py:446-457) for q / k / v / qx, then get_q / get_kv (:953-996) and
    the xattn q LayerNorm (:1008) - and against the library path on the same GPU (cuBLASLt fp8 matmul = what bmm_fp8
    calls, then this repository's prologue kernel)."""
    x = rnd((rows, hidden), 3, scale=1.0)
    heads = (Hq, Hkv, Hkv, Hq) if with_qx else (Hq, Hkv, Hkv)
    wq_ws = [_fp8_weights((h * 128, hidden), 4 + i) for i, h in enumerate(heads)]
    in_scale = np.full(hidden, 4.0 / 448.0, dtype=np.float32)
    lnp = [rnd((128,), 20 + i, scale=0.2) for i in range(6)]
    ang = np.random.default_rng(0).uniform(0, 6.28, size=(rows // batch, 48)).astype(np.float32)
    sin, cos = np.sin(ang), np.cos(ang)
    lin = [O.linear_fp8(x, O.e4m3_encode(wq), in_scale, in_scale, ws) for wq, ws in wq_ws]
    sq = rows // batch
    ref_q = O.get_q_or_k(lin[0].reshape(sq, batch, -1), Hq, lnp[0], lnp[1], sin, cos)
    ref_kv = O.get_kv(lin[1].reshape(sq, batch, -1), lin[2].reshape(sq, batch, -1), Hkv, lnp[2], lnp[3], sin, cos)
    t = lambda a: torch.from_numpy(a).cuda()
    w_fp8 = [torch.from_numpy(O.e4m3_encode(wq)).cuda().view(torch.float8_e4m3fn) for wq, _ in wq_ws]
    w_cat = F.stack_qkv_weights(*w_fp8)
    scales = [np.float32(in_scale[0]) * ws for _, ws in wq_ws] + ([] if with_qx else [np.float32(0)])
    out_scale = t(np.asarray(scales, dtype=np.float32))
    xq = F.fp8_quantize(bf(x), t(in_scale))
    cache = torch.zeros(rows + 50, Hkv, 256, dtype=torch.bfloat16, device="cuda")
    q, kv, qx = F.qkv_proj_norm_rope(xq, w_cat, Hq, Hkv, (t(lnp[0]), t(lnp[1])), (t(lnp[2]), t(lnp[3])), t(sin), t(cos),
                                     ln_qx=(bf(lnp[4]), bf(lnp[5])) if with_qx else None, batch=batch, cache=cache,
                                     cache_row_offset=50, cache_rows=rows - 7, out_scale=out_scale)

    def close(a, ref, what):
        a = a.float().cpu().numpy()
        err = np.abs(a - ref).max()
        rel = np.linalg.norm(a - ref) / np.linalg.norm(ref)
        assert a.shape == ref.shape and err <= 6e-2 and rel <= 6e-3, f"{what}: max_abs {err:.3e} rel_l2 {rel:.3e}"

    close(q, ref_q, "q")
    close(kv, ref_kv, "kv")
    if with_qx:
        close(qx, O.xattn_layer_norm(lin[3].reshape(rows, Hq, 128), O.bf16_round(lnp[4]), O.bf16_round(lnp[5])), "qx")
    assert torch.equal(cache[50:50 + rows - 7], kv[:rows - 7]) and float(cache[50 + rows - 7:].abs().sum()) == 0.0
    # the library path: cuBLASLt fp8 matmul per projection + the prologue kernel
    y = [F.fp8_scaled_linear(xq, w_, t(in_scale), t(np.asarray([ws], dtype=np.float32)))
         for w_, (_, ws) in zip(w_fp8, wq_ws)]
    q2, kv2 = F.qk_norm_rope(y[0], y[1], y[2], Hq, Hkv, (t(lnp[0]), t(lnp[1])), (t(lnp[2]), t(lnp[3])), t(sin), t(cos),
                             batch=batch)
    close(q, q2.float().cpu().numpy(), "q vs cuBLASLt path")
    close(kv, kv2.float().cpu().numpy(), "kv vs cuBLASLt path")
    close(y[0], lin[0], "cuBLASLt fp8 linear vs oracle")


@pytest.mark.parametrize("cls", ["pt", "pc"])
def test_fp8_linear_classes_match_reference_fixture(cls):
    """PerTensorQuantizedFp8Linear (unequal per-channel divisors) / PerChannelQuantizedFp8Linear with the reference's
    parameters and inputs: at most one bf16 ulp from the reference's output (fp32 summation order)."""
    from magi_b200.fp8_linear import PerChannelQuantizedFp8Linear, PerTensorQuantizedFp8Linear
    g = Golden("fp8_quant.npz")
    mod = (PerTensorQuantizedFp8Linear(256, 384) if cls == "pt" else PerChannelQuantizedFp8Linear(256, 128))
    sd = {}
    for k in g.files:
        if k.startswith(f"lin.{cls}.") and not k.endswith(".out"):
            a = g.raw(k)
            sd[k[len(f"lin.{cls}."):]] = (torch.from_numpy(a.copy()).view(torch.float8_e4m3fn) if a.dtype == np.uint8
                                          else g.torch(k))
    mod.load_state_dict(sd, strict=True)
    mod = mod.cuda()
    with torch.no_grad():
        out = mod(g.torch("lin.x", "cuda")).float().cpu().numpy()
    ref = g[f"lin.{cls}.out"]
    assert out.shape == ref.shape
    diff = np.abs(out - ref)
    assert (diff <= np.abs(ref) * 2.0 ** -7 + 1e-30).all() and (diff > 0).mean() < 2e-2


@pytest.mark.parametrize("variant,fused", [("shared", True), ("shared", False), ("split", True)])
def test_fp8_transformer_layer_matches_reference_golden(variant, fused):
    """The drop-in TransformerLayer with engine_config.fp8_quant (layer 1 of 4: fp8 q / qx / k / v, out-projection, fc1,
    fc2) against the reference's own fp8 TransformerLayer (tests/golden/fp8_quant.npz): the calibrated-checkpoint layout
    loads with strict=True; two forwards through the KV cache. `shared` runs the fused fp8 GEMM (or, with
    fused_qkv_proj off, the library GEMMs), `split` (unequal input scales) falls back by itself."""
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    from magi_b200.transformer_layer import TransformerLayer, high_precision_promote
    g = Golden("fp8_quant.npz")
    pre = variant + "."
    Hq, Hkv = [int(x) for x in g["heads"]]
    hidden = g.raw(pre + "w.self_attention.linear_qkv.layer_norm.weight").shape[0]
    mc = ModelConfig(num_layers=4, hidden_size=hidden, ffn_hidden_size=4 * hidden, num_attention_heads=Hq,
                     num_query_groups=Hkv)
    layer = high_precision_promote(TransformerLayer(mc, EngineConfig(fp8_quant=True, fused_qkv_proj=fused),
                                                    layer_number=1))
    sd = {}
    for name, key in g.weights(pre + "w.").items():
        a = g.raw(key)
        sd[name] = torch.from_numpy(a.copy()).view(torch.float8_e4m3fn) if a.dtype == np.uint8 else g.torch(key)
    layer.load_state_dict(sd, strict=True)
    layer = layer.cuda().eval()
    assert layer.self_attention.linear_qkv.fp8_fusable() == (variant == "shared")
    clip = int(g["clip"])
    inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
    inf.key_value_memory_dict[1] = torch.zeros(4 * clip, Hkv, 256, dtype=torch.bfloat16, device="cuda")
    i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32), device="cuda")
    for tag in ("A", "B"):
        t = pre + tag
        sp, dn, upd = [int(x) for x in g[f"{t}.meta"]]
        q_range, k_range = O.core_q_ranges(clip, dn, 1), g.raw(f"{t}.k_range")
        p = O.cross_attn_params(clip, dn, 1, g[f"{t}.y_index"])
        core = PackedCoreAttnParams(q_range=i32(q_range), k_range=i32(k_range), np_q_range=q_range, np_k_range=k_range,
                                    max_seqlen_q=clip, max_seqlen_k=O.max_seqlen_k(k_range))
        cross = PackedCrossAttnParams(q_ranges=i32(p["q_ranges"]), kv_ranges=i32(p["kv_ranges"]),
                                      cu_seqlens_q=i32(p["cu_seqlens_q"]), cu_seqlens_kv=i32(p["cu_seqlens_kv"]),
                                      max_seqlen_q=clip, max_seqlen_kv=800)
        meta = ModelMetaArgs(H=4, W=4, cp_pad_size=None, cp_split_sizes=None, slice_point=sp, denoising_range_num=dn,
                             range_num=sp + dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=False,
                             distill_nearly_clean_chunk=False, clip_token_nums=clip, enable_cuda_graph=False,
                             core_attn_params=core, cross_attn_params=cross)
        inf.update_kv_cache = bool(upd)
        hs, kvs = g.torch(f"{t}.hidden_states", "cuda"), g.torch(f"{t}.key_value_states", "cuda")
        rope = g.torch(f"{t}.rope", "cuda")
        core_out, x_out = layer.self_attention(hs, kvs, inf, rope, meta)
        out = layer(hs, g.torch(f"{t}.condition", "cuda"), g.torch(f"{t}.condition_map", "cuda"), kvs, rope, inf, meta)
        for name, got, ref in (("core", core_out, g[f"{t}.core"]), ("xattn", x_out, g[f"{t}.xattn"]),
                               ("out", out, g[f"{t}.out"])):
            got = got.float().cpu().numpy()
            rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            assert got.shape == ref.shape and rel < 2e-2, (variant, fused, tag, name, rel)
    cache = inf.key_value_memory_dict[1].float().cpu().numpy()[:3 * clip]
    ref_cache = g[pre + "cache"]
    err = np.abs(cache - ref_cache)
    assert (err <= np.maximum(np.abs(cache), np.abs(ref_cache)) * 2.0 ** -7 * 4 + 2e-2).all(), err.max()


def test_transformer_block_matches_reference_golden():
    """The drop-in TransformerBlock (the caller of the layers, dit_module.py:1373-1441) on the B200 against the
    reference's own block (tests/golden/transformer_block.npz: three layers, the middle one fp8 under fp8_quant, final
    fp32 LayerNorm): strict state-dict load under the reference's names, two forwards through the KV caches.
    Every layer on the reference's input of that layer: rel-L2 <= 2e-2 (the per-step criterion of SURVEY.md 8(c), as
    in the single-layer tests); final norm on the reference's last hidden state: 1e-5; end to end 5e-2 (the fp8 layer
    amplifies bf16-ulp input differences ~7x through e4m3 rounding flips and the final LayerNorm removes the large
    common component of the hidden state - the CPU oracle itself sits at 1.8e-2 from the reference's ATen run). With enable_cuda_graph the layer stack replays from one CUDA graph per forward
    signature and must agree with the eager block."""
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    from magi_b200.transformer_layer import TransformerBlock, high_precision_promote
    g = Golden("transformer_block.npz")
    L, hidden, ffn, Hq, Hkv = [int(x) for x in g["dims"]]
    mc = ModelConfig(num_layers=L, hidden_size=hidden, ffn_hidden_size=ffn, num_attention_heads=Hq,
                     num_query_groups=Hkv)
    sd = {}
    for name, key in g.weights().items():
        a = g.raw(key)
        sd[name] = torch.from_numpy(a.copy()).view(torch.float8_e4m3fn) if a.dtype == np.uint8 else g.torch(key)

    def build(graph):
        b = high_precision_promote(TransformerBlock(mc, EngineConfig(fp8_quant=True, enable_cuda_graph=graph)))
        b.load_state_dict(sd, strict=True)
        return b.cuda().eval()

    clip = int(g["clip"])
    i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32), device="cuda")

    def meta_for(tag):
        sp, dn, upd, extra = [int(x) for x in g[f"{tag}.meta"]]
        q_range, k_range = O.core_q_ranges(clip, dn, 1), g.raw(f"{tag}.k_range")
        p = O.cross_attn_params(clip, dn, 1, g[f"{tag}.y_index"])
        core = PackedCoreAttnParams(q_range=i32(q_range), k_range=i32(k_range), np_q_range=q_range, np_k_range=k_range,
                                    max_seqlen_q=clip, max_seqlen_k=O.max_seqlen_k(k_range))
        cross = PackedCrossAttnParams(q_ranges=i32(p["q_ranges"]), kv_ranges=i32(p["kv_ranges"]),
                                      cu_seqlens_q=i32(p["cu_seqlens_q"]), cu_seqlens_kv=i32(p["cu_seqlens_kv"]),
                                      max_seqlen_q=clip, max_seqlen_kv=800)
        return upd, ModelMetaArgs(H=4, W=4, cp_pad_size=None, cp_split_sizes=None, slice_point=sp,
                                  denoising_range_num=dn, range_num=sp + dn, extract_prefix_video_feature=False,
                                  fwd_extra_1st_chunk=bool(extra), distill_nearly_clean_chunk=False,
                                  clip_token_nums=clip, enable_cuda_graph=False, core_attn_params=core,
                                  cross_attn_params=cross)

    def rel(got, ref):
        got = got.float().cpu().numpy()
        ref = np.asarray(ref, np.float32).reshape(got.shape)
        return float(np.linalg.norm(got - ref) / np.linalg.norm(ref))

    def new_inf():
        inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
        for l in range(L):
            inf.key_value_memory_dict[l] = torch.zeros(4 * clip, Hkv, 256, dtype=torch.bfloat16, device="cuda")
        return inf

    block, inf = build(False), new_inf()
    for tag in ("A", "B"):          # (1) per layer, on the reference's inputs
        upd, meta = meta_for(tag)
        inf.update_kv_cache = bool(upd)
        args = (g.torch(f"{tag}.condition", "cuda"), g.torch(f"{tag}.condition_map", "cuda"),
                g.torch(f"{tag}.key_value_states", "cuda"), g.torch(f"{tag}.rope", "cuda"), inf, meta)
        for l, layer in enumerate(block.layers):
            x = g.torch(f"{tag}.hidden_states" if l == 0 else f"{tag}.after_layer{l - 1}", "cuda")
            r = rel(layer(x, *args), g[f"{tag}.after_layer{l}"])
            print(f"block {tag} layer {l} (reference input): rel-L2 {r:.3e}")
            assert r <= 2e-2, (tag, l, r)
        with torch.no_grad():
            normed = block.final_layernorm(g.torch(f"{tag}.after_layer{L - 1}", "cuda").float())
        assert rel(normed, g[f"{tag}.out"]) <= 1e-5
    eager, graphed = build(False), build(True)
    inf_e, inf_g = new_inf(), new_inf()
    for rep in range(2):            # (2) end to end, eager and graphed (second round: replays)
        for tag in ("A", "B"):
            upd, meta = meta_for(tag)
            inf_e.update_kv_cache = inf_g.update_kv_cache = bool(upd)
            tensors = [g.torch(f"{tag}.{k}", "cuda") for k in ("hidden_states", "condition", "condition_map",
                                                               "key_value_states", "rope")]
            out_e = eager(*tensors, inf_e, meta)
            out_g = graphed(*tensors, inf_g, meta)
            assert out_e.dtype == torch.float32 and out_g.dtype == torch.float32
            r, rg = rel(out_e, g[f"{tag}.out"]), rel(out_g, out_e.cpu().numpy())
            print(f"block {tag} end to end: rel-L2 {r:.3e}; graph vs eager {rg:.3e}")
            assert r <= 5e-2, (tag, r)          # measured on the B200: 1.8e-2 / 2.0e-2 (profiles/r02blk_pytest.log)
            # cuBLAS may pick other GEMM algorithms under stream capture, and the fp8 layer amplifies that: same bound
            # (measured: bit-identical)
            assert rg <= 5e-2, tag
    st = graphed._graphed
    assert st is not None and (st.disabled_reason is not None or (st.captures == 2 and st.replays == 4)), \
        (st.captures, st.replays, st.disabled_reason)
