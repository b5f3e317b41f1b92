"""GPU parity tests (B200): the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded inputs,
against the committed golden fixtures of the reference, and - at BASELINE.json's full sizes - through
size-independent properties and the flash-attn library the reference itself calls.

Tolerances (bf16 eps = 2^-8):
  attention output vs fp32 oracle : max_abs <= 2e-2 and rel-L2 <= 1e-2 for N(0,1) inputs
  prologue output vs fp32 oracle  : <= 1 bf16 ulp (|x| * 2^-7) element-wise
  data movement / cache append    : bit exact
"""
import numpy as np
import pytest
import torch

import magi_oracle as O
from golden_io import Golden

pytestmark = pytest.mark.gpu

ATTN_MAX_ABS = 2e-2
ATTN_REL_L2 = 1e-2


@pytest.fixture(scope="module")
def F():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import magi_b200.functional as F_
    return F_


def bf(x, dev="cuda"):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev).to(torch.bfloat16)


def rnd(shape, seed, scale=1.0):
    """bf16-representable N(0, scale) values as fp32 numpy (so oracle and kernel see identical inputs)."""
    g = torch.Generator().manual_seed(seed)
    t = (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)
    return t.float().numpy()


def i32(a):
    return torch.as_tensor(np.asarray(a, dtype=np.int32), device="cuda")


def check_attn(out, ref, what=""):
    out = out.float().cpu().numpy()
    err = np.abs(out - ref).max()
    rel = np.linalg.norm(out - ref) / max(np.linalg.norm(ref), 1e-9)
    assert err <= ATTN_MAX_ABS and rel <= ATTN_REL_L2, f"{what}: max_abs {err:.3e} rel_l2 {rel:.3e}"


# ------------------------------------------------------------------------------------------ attention vs oracle
RANGE_CASES = [
    # name, Tq, Tk, Hq, Hkv, q_ranges, k_ranges, cache_len
    ("one tile", 128, 128, 1, 1, [[0, 128]], [[0, 128]], 0),
    ("tiny", 5, 3, 2, 1, [[0, 5]], [[0, 3]], 0),
    ("cfg1 1536", 1536, 1536, 24, 8, [[0, 1536]], [[0, 1536]], 0),
    ("ragged 3 ranges", 1000, 1300, 3, 1, [[0, 300], [300, 650], [650, 1000]], [[0, 333], [100, 1300], [900, 1001]], 0),
    ("block causal window", 4 * 331, 4 * 331, 6, 2, [[i * 331, (i + 1) * 331] for i in range(4)],
     [[max(0, i - 1) * 331, (i + 1) * 331] for i in range(4)], 0),
    ("cache+new split mid-range", 700, 1500, 6, 2, [[0, 350], [350, 700]], [[0, 1150], [200, 1500]], 800),
    ("cache only", 300, 900, 4, 4, [[0, 300]], [[100, 640]], 900),
    ("empty k, uncovered rows", 600, 400, 2, 2, [[0, 200], [300, 600]], [[50, 50], [0, 400]], 0),
    ("unsorted ranges", 512, 700, 2, 1, [[256, 512], [0, 256]], [[0, 700], [300, 429]], 0),
    ("many short ranges", 64 * 9, 64 * 9, 2, 2, [[i * 64, (i + 1) * 64] for i in range(9)],
     [[0, (i + 1) * 64] for i in range(9)], 0),
]


@pytest.mark.parametrize("case", RANGE_CASES, ids=[c[0] for c in RANGE_CASES])
def test_attn_ranges_vs_oracle(F, case):
    name, Tq, Tk, Hq, Hkv, qr, kr, cache_len = case
    q, kv = rnd((Tq, Hq, 128), 1), rnd((Tk, Hkv, 256), 2)
    ref, ref_lse = O.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, round_output=False)
    tq, tkv = bf(q), bf(kv)
    if cache_len:
        kc, vc = tkv[:cache_len, :, :128], tkv[:cache_len, :, 128:]
        kn, vn = tkv[cache_len:, :, :128], tkv[cache_len:, :, 128:]
        if kn.shape[0] == 0:
            kn = vn = None
    else:
        kc = vc = None
        kn, vn = tkv[..., :128], tkv[..., 128:]
    if kn is None:   # everything lives in the cache segment
        out, lse = F.attn_ranges(tq, torch.empty(0, Hkv, 128, device="cuda", dtype=torch.bfloat16),
                                 torch.empty(0, Hkv, 128, device="cuda", dtype=torch.bfloat16), i32(qr), i32(kr),
                                 k_cache=kc, v_cache=vc, zero_fill=True)
    else:
        out, lse = F.attn_ranges(tq, kn, vn, i32(qr), i32(kr), k_cache=kc, v_cache=vc, zero_fill=True)
    check_attn(out, ref, name)
    lse = lse.cpu().numpy()
    fin = np.isfinite(ref_lse)
    assert np.array_equal(np.isfinite(lse), fin), "lse -inf pattern (uncovered rows / empty key ranges)"
    assert np.abs(lse[fin] - ref_lse[fin]).max() < 1e-3


@pytest.mark.parametrize("case", RANGE_CASES, ids=[c[0] for c in RANGE_CASES])
def test_pair_kernel_vs_oracle_and_single_kernel(F, case):
    """The CTA-pair kernel (attn_pair.cu: cta_group::2 MMAs, double-buffered S, quad-fragment softmax; `kernel=2`) on
    the same range layouts: against the fp32 oracle with the same tolerance, and against the default kernel (same
    math in a different association order: a few bf16 ulps)."""
    name, Tq, Tk, Hq, Hkv, qr, kr, cache_len = case
    q, kv = rnd((Tq, Hq, 128), 1), rnd((Tk, Hkv, 256), 2)
    ref, ref_lse = O.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, round_output=False)
    tq, tkv = bf(q), bf(kv)
    kc = vc = None
    if cache_len:
        kc, vc = tkv[:cache_len, :, :128], tkv[:cache_len, :, 128:]
    kn, vn = tkv[cache_len:, :, :128], tkv[cache_len:, :, 128:]
    outs = {}
    for kern in (1, 2):
        out, lse = F.attn_ranges(tq, kn, vn, i32(qr), i32(kr), k_cache=kc, v_cache=vc, zero_fill=True, kernel=kern)
        check_attn(out, ref, f"{name} kernel {kern}")
        fin = np.isfinite(ref_lse)
        assert np.abs(lse.cpu().numpy()[fin] - ref_lse[fin]).max() < 1e-3
        outs[kern] = out.float()
    assert (outs[1] - outs[2]).abs().max().item() <= 1.6e-2


def test_pair_kernel_rescale_path_and_many_items(F):
    """Pair kernel: rows whose running max jumps by far more than 2^8 (O rescaled in TMEM through the 16x256b fragment
    shape, behind the P V completion barrier) and a launch with several items per CTA pair and both K/V segments."""
    Tq, Tk, Hq, Hkv = 300, 1100, 4, 2
    q, k, v = rnd((Tq, Hq, 128), 1), rnd((Tk, Hkv, 128), 2), rnd((Tk, Hkv, 128), 3)
    ramp = np.ones(Tk, dtype=np.float32)
    ramp[256:384] = 6.0
    ramp[640:700] = 14.0
    ramp[1000:] = 0.25
    k = O.bf16_round(k * ramp[:, None, None])
    q[::7] = 0.0
    ref, ref_lse = O.attn_ranges(q, k, v, [[0, 300]], [[0, 1100]], round_output=False)
    out, lse = F.attn_ranges(bf(q), bf(k), bf(v), i32([[0, 300]]), i32([[0, 1100]]), kernel=2)
    check_attn(out, ref, "pair kernel, growing max")
    assert np.abs(lse.cpu().numpy() - ref_lse).max() < 2e-3
    qr = [[i * 520, (i + 1) * 520] for i in range(6)]
    kr = [[0, (i + 1) * 520] for i in range(6)]
    q, kv = rnd((3120, 24, 128), 5), rnd((3120, 8, 256), 6)
    ref, _ = O.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, round_output=False)
    tkv = bf(kv)
    out, _ = F.attn_ranges(bf(q), tkv[1000:, :, :128], tkv[1000:, :, 128:], i32(qr), i32(kr), k_cache=tkv[:1000, :, :128],
                           v_cache=tkv[:1000, :, 128:], kernel=2, num_sms=8)
    check_attn(out, ref, "pair kernel, 888 items on 4 pairs")


@pytest.mark.parametrize("kernel", [1, 2])
def test_padding_rows_inside_the_view_may_be_nan(F, kernel):
    """K / V rows that lie inside the tensor views but outside every key range (torch.empty padding: here NaN) must
    not reach the output: masked scores are forced to -inf whatever K holds, the 16-key groups of a partial last tile
    past the range are never multiplied, and the rows of the straddling group are zeroed in shared memory
    (flash-attn's predicated loads never read such rows). kernel 2 (experimental pair kernel) keeps
    the documented precondition instead: finite padding."""
    Tq, Tk, Hq, Hkv = 400, 1000, 4, 2
    qr, kr = [[0, 130], [130, 400]], [[7, 300], [450, 777]]
    q, kv = rnd((Tq, Hq, 128), 31), rnd((Tk, Hkv, 256), 32)
    ref, _ = O.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, round_output=False)
    bad = kv.copy()
    fill = np.nan if kernel == 1 else 1e4
    for a, b in ((0, 7), (300, 450), (777, 1000)):
        bad[a:b] = fill
    tkv = bf(bad)
    out, _ = F.attn_ranges(bf(q), tkv[..., :128], tkv[..., 128:], i32(qr), i32(kr), kernel=kernel)
    assert bool(torch.isfinite(out.float()).all())
    check_attn(out, ref, "NaN padding" if kernel == 1 else "finite padding")
    # the same through the cache / new-token seam with the seam inside the padding
    out2, _ = F.attn_ranges(bf(q), tkv[400:, :, :128], tkv[400:, :, 128:], i32(qr), i32(kr), k_cache=tkv[:400, :, :128],
                            v_cache=tkv[:400, :, 128:], kernel=kernel)
    check_attn(out2, ref, "padding across the seam")


@pytest.mark.parametrize("seed", range(6))
def test_attn_random_range_sets_vs_oracle(F, seed):
    """Randomised disjoint q ranges with arbitrary (overlapping, ragged) k ranges."""
    rng = np.random.default_rng(seed)
    Hkv = int(rng.choice([1, 2, 4]))
    Hq = Hkv * int(rng.choice([1, 3, 6]))
    n = int(rng.integers(1, 6))
    cuts = np.sort(rng.choice(np.arange(1, 900), size=n, replace=False))
    q_edges = np.concatenate([[0], cuts])
    Tq = int(q_edges[-1]) + int(rng.integers(0, 40))
    Tk = int(rng.integers(200, 1400))
    qr = [[int(q_edges[i]), int(q_edges[i + 1])] for i in range(n)]
    kr = []
    for _ in range(n):
        a = int(rng.integers(0, Tk - 1))
        b = int(rng.integers(a, Tk + 1))
        kr.append([a, b])
    cache_len = int(rng.integers(0, Tk)) if seed % 2 else 0
    q, kv = rnd((Tq, Hq, 128), 10 + seed), rnd((Tk, Hkv, 256), 20 + seed)
    ref, _ = O.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, round_output=False)
    tq, tkv = bf(q), bf(kv)
    kc = vc = None
    if cache_len:
        kc, vc = tkv[:cache_len, :, :128], tkv[:cache_len, :, 128:]
    out, _ = F.attn_ranges(tq, tkv[cache_len:, :, :128], tkv[cache_len:, :, 128:], i32(qr), i32(kr), k_cache=kc,
                           v_cache=vc, zero_fill=True)
    check_attn(out, ref, f"seed {seed}: n={n} Hq={Hq} Hkv={Hkv} Tq={Tq} Tk={Tk} cache={cache_len}")


def test_running_max_jumps_take_the_rescale_path(F):
    """Later key tiles carry much larger scores, so the running row max grows by far more than 2^8 between tiles:
    exercises the lazy-rescale branch (redo of the optimistic first half, O rescale in TMEM) that N(0,1) data never
    reaches; some rows jump, others (zero queries) do not."""
    Tq, Tk, Hq, Hkv = 300, 1100, 4, 2
    q, k, v = rnd((Tq, Hq, 128), 1), rnd((Tk, Hkv, 128), 2), rnd((Tk, Hkv, 128), 3)
    ramp = np.ones(Tk, dtype=np.float32)
    ramp[256:384] = 6.0
    ramp[640:700] = 14.0
    ramp[1000:] = 0.25
    k = O.bf16_round(k * ramp[:, None, None])
    q[::7] = 0.0            # rows whose scores stay flat
    qr, kr = [[0, 300]], [[0, 1100]]
    ref, ref_lse = O.attn_ranges(q, k, v, qr, kr, round_output=False)
    out, lse = F.attn_ranges(bf(q), bf(k), bf(v), i32(qr), i32(kr))
    check_attn(out, ref, "growing max")
    assert np.abs(lse.cpu().numpy() - ref_lse).max() < 2e-3
    # decreasing scores (max set by the first tile, later tiles underflow towards 0) as the mirror case
    k2 = O.bf16_round(k[::-1].copy())
    v2 = np.ascontiguousarray(v[::-1])
    ref2, _ = O.attn_ranges(q, k2, v2, qr, kr, round_output=False)
    out2, _ = F.attn_ranges(bf(q), bf(k2), bf(v2), i32(qr), i32(kr))
    check_attn(out2, ref2, "shrinking max")
    assert np.abs(ref - ref2).max() < 1e-4   # permutation invariance of the oracle itself


def test_drop_in_signatures(F):
    """flash_attn_func / flash_attn_varlen_func / flex_flash_attn_func call shapes of dit_module.py:1032-1133."""
    q, k, v = rnd((2, 200, 6, 128), 3), rnd((2, 333, 2, 128), 4), rnd((2, 333, 2, 128), 5)
    out = F.flash_attn_func(bf(q), bf(k), bf(v), deterministic=False)
    for b in range(2):
        ref, _ = O.attn_ranges(q[b], k[b], v[b], [[0, 200]], [[0, 333]], round_output=False)
        check_attn(out[b], ref, f"flash_attn_func batch {b}")
    cu_q, cu_k = [0, 120, 120, 400], [0, 50, 90, 131]     # middle q segment empty
    qv, kv_, vv = rnd((400, 4, 128), 6), rnd((131, 4, 128), 7), rnd((131, 4, 128), 8)
    out = F.flash_attn_varlen_func(bf(qv), bf(kv_), bf(vv), i32(cu_q), i32(cu_k), 280, 50, deterministic=False)
    check_attn(out, O.attn_varlen(qv, kv_, vv, cu_q, cu_k), "flash_attn_varlen_func")
    qr, kr = [[0, 150], [150, 400]], [[0, 131], [40, 90]]
    out, lse = F.flex_flash_attn_func(bf(qv), bf(kv_), bf(vv), i32(qr), i32(kr), 250, 131, softmax_scale=None,
                                      deterministic=False, disable_fwd_atomic_reduction=True)
    ref, ref_lse = O.attn_ranges(qv, kv_, vv, qr, kr, round_output=False)
    check_attn(out, ref, "flex_flash_attn_func")
    assert np.abs(lse.cpu().numpy() - ref_lse).max() < 1e-3
    with pytest.raises(ValueError):
        F.flex_flash_attn_func(bf(qv), bf(kv_), bf(vv), i32(qr), i32(kr), 250, 131, disable_fwd_atomic_reduction=False)
    # softmax_scale override
    out, _ = F.attn_ranges(bf(qv), bf(kv_), bf(vv), i32(qr), i32(kr), softmax_scale=0.05)
    check_attn(out, O.attn_ranges(qv, kv_, vv, qr, kr, softmax_scale=0.05, round_output=False)[0], "scale")


def test_attention_is_deterministic(F):
    q, kv = bf(rnd((777, 6, 128), 1)), bf(rnd((1111, 2, 256), 2))
    qr, kr = i32([[0, 400], [400, 777]]), i32([[0, 1111], [300, 900]])
    a, la = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)
    b, lb = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)
    assert torch.equal(a, b) and torch.equal(la, lb)


# ------------------------------------------------------------------------------------------ prologue vs oracle
def ulp_ok(out, ref, ulps=1.0):
    out = out.float().cpu().numpy()
    tol = np.maximum(np.abs(out), np.abs(ref)) * 2.0 ** -7 * ulps + 1e-6
    bad = np.abs(out - ref) > tol
    return not bad.any(), float(np.abs(out - ref).max()), int(bad.sum())


@pytest.mark.parametrize("batch", [1, 2])
def test_prologue_vs_oracle(F, batch):
    sq, Hq, Hkv = 333, 6, 2
    xq, xk, xv = rnd((sq, batch, Hq * 128), 1, 2.0), rnd((sq, batch, Hkv * 128), 2, 2.0), rnd((sq, batch, Hkv * 128), 3)
    wq, bq, wk, bk = rnd((128,), 4, 0.2), rnd((128,), 5, 0.2), rnd((128,), 6, 0.2), rnd((128,), 7, 0.2)
    sin, cos = O.rope_table(4, 9, 37)
    sin, cos = sin[-sq:], cos[-sq:]
    ref_q = O.get_q_or_k(xq, Hq, wq, bq, sin, cos)
    ref_kv = O.get_kv(xk, xv, Hkv, wk, bk, sin, cos)
    t32 = lambda a: torch.from_numpy(a).cuda()
    cache = torch.zeros(2000, Hkv, 256, dtype=torch.bfloat16, device="cuda")
    rows = sq * batch
    q, kv = F.qk_norm_rope(bf(xq).view(rows, -1), bf(xk).view(rows, -1), bf(xv).view(rows, -1), Hq, Hkv,
                           (t32(wq), t32(bq)), (t32(wk), t32(bk)), t32(sin), t32(cos), batch=batch, cache=cache,
                           cache_row_offset=100, cache_rows=rows - 50)
    ok, err, nbad = ulp_ok(q, ref_q)
    assert ok, f"q: max err {err}, {nbad} beyond 1 bf16 ulp"
    ok, err, nbad = ulp_ok(kv, ref_kv)
    assert ok, f"kv: max err {err}, {nbad} beyond 1 bf16 ulp"
    assert torch.equal(kv[..., 128:], bf(xv).view(rows, Hkv, 128)), "V must be a bit-exact copy"
    # fused KV-cache append == reference slice-assign (dit_module.py:929); untouched rows stay untouched
    assert torch.equal(cache[100:100 + rows - 50], kv[:rows - 50])
    assert cache[:100].abs().sum() == 0 and cache[100 + rows - 50:].abs().sum() == 0


def test_prologue_bf16_params_no_rope(F):
    """The cross-attention norms: bf16 parameters, `weight + 1` rounded to bf16, no RoPE (dit_module.py:1005-1021),
    K and V read out of the packed (tokens, Hkv, 256) projection."""
    T, Hq, Hkv = 211, 4, 2
    xq, kvx = rnd((T, Hq * 128), 1, 3.0), rnd((T, Hkv, 256), 2, 3.0)
    w, b = rnd((128,), 3, 0.3), rnd((128,), 4, 0.3)
    ref_q = O.xattn_layer_norm(xq.reshape(T, Hq, 128), w, b)
    ref_k = O.xattn_layer_norm(kvx[..., :128], w, b)
    tkv = bf(kvx)
    q, kv = F.qk_norm_rope(bf(xq), tkv[..., :128], tkv[..., 128:], Hq, Hkv, (bf(w), bf(b)), (bf(w), bf(b)), None, None)
    ok, err, nbad = ulp_ok(q, ref_q)
    assert ok, f"xq: {err} {nbad}"
    ok, err, nbad = ulp_ok(kv[..., :128], ref_k)
    assert ok, f"xk: {err} {nbad}"
    assert torch.equal(kv[..., 128:], tkv[..., 128:])


@pytest.mark.parametrize("mode", ["plain+cache", "send-layout", "gather", "cross-only", "ragged-heads"])
def test_one_prologue_launch_per_layer_matches_three(F, mode):
    """`cross=`: the cross-attention's qx / text k / text v in the SAME launch as q / k / v (fp32 LayerNorm + RoPE there,
    bf16 LayerNorm and no RoPE here, a different row count for the text): every output bit-identical to the three
    separate launches, whatever the routing of q / k / v (plain + KV-cache append, all-to-all send layout, fused
    all-gather), with the self-attention tensors absent, and with head counts that are not multiples of 8."""
    T, Tt = 333, 77
    Hq, Hkv = (6, 3) if mode == "ragged-heads" else (8, 2)
    xq, xk, xv = bf(rnd((T, Hq * 128), 1, 2.0)), bf(rnd((T, Hkv * 128), 2, 2.0)), bf(rnd((T, Hkv * 128), 3, 2.0))
    qx, tkv = bf(rnd((T, Hq * 128), 4, 2.0)), bf(rnd((Tt, Hkv, 256), 5, 2.0))
    ln = [torch.from_numpy(rnd((128,), 10 + i, 0.2)).cuda() for i in range(4)]
    lnx = [bf(rnd((128,), 20 + i, 0.2)) for i in range(4)]
    ang = torch.from_numpy(rnd((T, 48), 30, 3.0))
    sin, cos = torch.sin(ang).cuda(), torch.cos(ang).cuda()
    ref_qx, _ = F.qk_norm_rope(qx, None, None, Hq, Hkv, (lnx[0], lnx[1]), None, None, None)
    _, ref_tkv = F.qk_norm_rope(None, tkv[..., :128], tkv[..., 128:], Hq, Hkv, None, (lnx[2], lnx[3]), None, None)
    cross = dict(x_qx=qx, ln_qx=(lnx[0], lnx[1]), text_kv=tkv, ln_tk=(lnx[2], lnx[3]))
    main = (xq, xk, xv, Hq, Hkv, (ln[0], ln[1]), (ln[2], ln[3]), sin, cos)
    if mode == "cross-only":
        q, kv = F.qk_norm_rope(None, None, None, Hq, Hkv, None, None, None, None, cross=cross)
        assert q is None and kv is None
    elif mode in ("plain+cache", "ragged-heads"):
        c0 = torch.zeros(T + 50, Hkv, 256, dtype=torch.bfloat16, device="cuda")
        c1 = torch.zeros_like(c0)
        rq, rkv = F.qk_norm_rope(*main, cache=c0, cache_row_offset=50, cache_rows=T - 10)
        q, kv = F.qk_norm_rope(*main, cache=c1, cache_row_offset=50, cache_rows=T - 10, cross=cross)
        assert torch.equal(q, rq) and torch.equal(kv, rkv) and torch.equal(c0, c1)
    elif mode == "send-layout":
        rq, rkv = F.qk_norm_rope(*main, head_perm_cp=2)
        q, kv = F.qk_norm_rope(*main, head_perm_cp=2, cross=cross)
        assert torch.equal(q, rq) and torch.equal(kv, rkv)
    else:   # fused K/V all-gather into two "ranks'" buffers, this rank's stripe = rows [111, 222) of every chunk of 333
        g0 = [torch.zeros(3 * 333, Hkv, 256, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
        g1 = [torch.zeros_like(t) for t in g0]
        rq, _ = F.qk_norm_rope(*main, gather_kv_out=[t.data_ptr() for t in g0], gather_geometry=(111, 333, 111))
        q, _ = F.qk_norm_rope(*main, gather_kv_out=[t.data_ptr() for t in g1], gather_geometry=(111, 333, 111),
                              cross=cross)
        assert torch.equal(q, rq) and all(torch.equal(a, b) for a, b in zip(g0, g1))
    assert torch.equal(cross["qx"], ref_qx) and torch.equal(cross["tkv"], ref_tkv)


@pytest.mark.parametrize("cp,Hq,Hkv", [(2, 6, 2), (4, 8, 2), (8, 24, 8)])
def test_prologue_writes_ulysses_send_layout(F, cp, Hq, Hkv):
    """head_perm_cp: outputs equal `repeat_interleave` + "seq (cp hn) hd -> (cp seq) hn hd" of the plain outputs
    (context_parallel.py:393-396), bit exact."""
    sq = 97
    xq, xk, xv = rnd((sq, Hq * 128), 1), rnd((sq, Hkv * 128), 2), rnd((sq, Hkv * 128), 3)
    wq, bq = rnd((128,), 4, 0.2), rnd((128,), 5, 0.2)
    sin, cos = O.rope_table(1, 1, sq)
    t32 = lambda a: torch.from_numpy(a).cuda()
    args = (bf(xq), bf(xk), bf(xv), Hq, Hkv, (t32(wq), t32(bq)), (t32(wq), t32(bq)), t32(sin), t32(cos))
    q0, kv0 = F.qk_norm_rope(*args)
    q1, kv1 = F.qk_norm_rope(*args, head_perm_cp=cp)
    rep = cp // Hkv if cp > Hkv else 1
    exp_q = q0.view(sq, cp, Hq // cp, 128).permute(1, 0, 2, 3).reshape(cp * sq, Hq // cp, 128)
    kvr = kv0.repeat_interleave(rep, dim=1)
    exp_kv = kvr.view(sq, cp, kvr.shape[1] // cp, 256).permute(1, 0, 2, 3).reshape(cp * sq, -1, 256)
    assert torch.equal(q1, exp_q) and torch.equal(kv1, exp_kv)
    # the stand-alone re-layout kernel does the same, and mode 1 inverts it
    assert torch.equal(F.ulysses_relayout(q0.contiguous(), cp, 0), exp_q)
    assert torch.equal(F.ulysses_relayout(kv0.contiguous(), cp, 0, kv_repeat=rep), exp_kv)
    assert torch.equal(F.ulysses_relayout(exp_q.contiguous(), cp, 1), q0)
    wide = torch.zeros(sq, cp * (Hq // cp + 2), 128, dtype=torch.bfloat16, device="cuda")
    F.ulysses_relayout(exp_q.contiguous(), cp, 1, out=wide, dst_heads_total=Hq // cp + 2, dst_head_offset=1)
    w4 = wide.view(sq, cp, Hq // cp + 2, 128)
    assert torch.equal(w4[:, :, 1:-1].reshape(sq, Hq, 128), q0) and w4[:, :, 0].abs().sum() == 0


def test_kv_append_and_lse_merge(F):
    kv = bf(rnd((300, 2, 256), 1))
    cache = torch.zeros(1000, 2, 256, dtype=torch.bfloat16, device="cuda")
    F.kv_cache_append(cache, kv, 250, 123)
    assert torch.equal(cache[250:373], kv[:123]) and cache[:250].abs().sum() == 0 and cache[373:].abs().sum() == 0
    q, k, v = rnd((500, 4, 128), 2), rnd((700, 2, 128), 3), rnd((700, 2, 128), 4)
    tq, tk, tv = bf(q), bf(k), bf(v)
    qr = i32([[0, 500]])
    a, la = F.attn_ranges(tq, tk, tv, qr, i32([[0, 300]]))
    b, lb = F.attn_ranges(tq, tk, tv, qr, i32([[300, 700]]))
    e, le = F.attn_ranges(tq, tk, tv, qr, i32([[5, 5]]))          # empty partner
    a0 = a.clone()
    F.lse_merge(a, la, e, le)
    assert torch.equal(a, a0)
    F.lse_merge(a, la, b, lb)
    ref, ref_lse = O.attn_ranges(q, k, v, [[0, 500]], [[0, 700]], round_output=False)
    check_attn(a, ref, "merge of two key halves")
    assert np.abs(la.cpu().numpy() - ref_lse).max() < 1e-3


# ------------------------------------------------------------------------------------------ golden (reference) parity
def _build_module(g, cp_strategy="none"):
    from magi_b200.attention import FullyParallelAttention
    from magi_b200.common import EngineConfig, ModelConfig
    Hq = g["A.q"].shape[1] if "A.q" in g else int(g["heads"][0])
    Hkv = g["A.kv"].shape[1] if "A.kv" in g else int(g["heads"][1])
    hidden = g.raw("w.linear_qkv.layer_norm.weight").shape[0]
    mc = ModelConfig(num_layers=4, hidden_size=hidden, num_attention_heads=Hq, num_query_groups=Hkv)
    attn = FullyParallelAttention(mc, EngineConfig(cp_size=1, cp_strategy=cp_strategy), layer_number=1)
    sd = {}
    for name, key in g.weights().items():
        t = g.torch(key)
        sd[name] = t
    for n_ in ("q_layernorm", "k_layernorm"):   # the reference promotes these to fp32 (dit_model.py:621-633)
        m = getattr(attn, n_)
        m.weight.data = m.weight.data.float()
        m.bias.data = m.bias.data.float()
    missing, unexpected = attn.load_state_dict(sd, strict=True)
    return attn.cuda().eval()


@pytest.mark.parametrize("cp,Hq,Hkv", [(4, 8, 4), (4, 8, 2), (2, 6, 2), (8, 8, 1)])
def test_fused_all_to_all_routing_on_virtual_ranks(F, cp, Hq, Hkv):
    """The peer-memory routing of the fused Ulysses all-to-alls, exercised on ONE device: every "rank" is a set of
    local buffers whose addresses are handed to the kernels as peer pointers. The prologue must deliver exactly what
    all_to_all_input_split delivers (context_parallel.py:381-401, incl. KV-head replication), the attention
    epilogue exactly all_to_all_output_split + "(cp sq) hn hd -> sq (cp hn) hd" (:404-423, 596): bit-exact against the
    single-rank run of the same kernels."""
    clip, dn = 150, 3
    S = clip * dn
    split = [S // cp + (1 if i < S % cp else 0) for i in range(cp)]
    offs = np.concatenate([[0], np.cumsum(split)]).astype(np.int64)
    xq, xk, xv = bf(rnd((S, Hq * 128), 1)), bf(rnd((S, Hkv * 128), 2)), bf(rnd((S, Hkv * 128), 3))
    ln = [torch.from_numpy(rnd((128,), 10 + i, 0.2)).cuda() for i in range(4)]
    ang = torch.from_numpy(rnd((S, 48), 20, 3.0))
    sin, cos = torch.sin(ang).cuda(), torch.cos(ang).cuda()
    q_ranges = i32([[i * clip, (i + 1) * clip] for i in range(dn)])
    k_ranges = i32([[0, clip], [0, 2 * clip], [clip, 3 * clip]])
    q_full, kv_full = F.qk_norm_rope(xq, xk, xv, Hq, Hkv, (ln[0], ln[1]), (ln[2], ln[3]), sin, cos)
    out_full, _ = F.attn_ranges(q_full, kv_full[..., :128], kv_full[..., 128:], q_ranges, k_ranges)
    hn, hkv_loc = Hq // cp, max(1, Hkv // cp)
    recv_q = [torch.zeros(S, hn, 128, dtype=torch.bfloat16, device="cuda") for _ in range(cp)]
    recv_kv = [torch.zeros(S, hkv_loc, 256, dtype=torch.bfloat16, device="cuda") for _ in range(cp)]
    for me in range(cp):
        sl = slice(int(offs[me]), int(offs[me + 1]))
        F.qk_norm_rope(xq[sl], xk[sl], xv[sl], Hq, Hkv, (ln[0], ln[1]), (ln[2], ln[3]), sin[sl], cos[sl],
                       head_perm_cp=cp, peer_q_out=[t.data_ptr() for t in recv_q],
                       peer_kv_out=[t.data_ptr() for t in recv_kv], peer_row_offset=int(offs[me]),
                       peer_heads=(hn, hkv_loc))
    rep = cp // Hkv if cp > Hkv else 1
    for c in range(cp):
        assert torch.equal(recv_q[c], q_full[:, c * hn:(c + 1) * hn])
        k_lo = c // rep if cp > Hkv else c * hkv_loc
        assert torch.equal(recv_kv[c], kv_full[:, k_lo:k_lo + hkv_loc])
    out_recv = [torch.zeros(split[r], Hq, 128, dtype=torch.bfloat16, device="cuda") for r in range(cp)]
    for c in range(cp):
        o, l = F.attn_ranges(recv_q[c], recv_kv[c][..., :128], recv_kv[c][..., 128:], q_ranges, k_ranges,
                             return_lse=False, peer_out=([t.data_ptr() for t in out_recv], offs.tolist(), c * hn, Hq))
        assert o is None and l is None
    assert torch.equal(torch.cat(out_recv, 0), out_full)
    if Hq % 8 == 0:
        # fused_proj_input under the fused path (peer_proj): the same rows stored straight into slot n = 0 of every
        # owner's out-projection input "sq (hn n hd)", hn = 8 (attn_linear_proj, dit_module.py:1332-1338)
        proj = [torch.zeros(split[r], 2 * Hq * 128, dtype=torch.bfloat16, device="cuda") for r in range(cp)]
        for c in range(cp):
            F.attn_ranges(recv_q[c], recv_kv[c][..., :128], recv_kv[c][..., 128:], q_ranges, k_ranges,
                          return_lse=False, peer_out=([t.data_ptr() for t in proj], offs.tolist(), c * hn, Hq),
                          peer_proj=(0, 8))
        got = torch.cat(proj, 0).view(S, 8, 2, Hq // 8, 128)
        assert torch.equal(got[:, :, 0].reshape(S, Hq, 128), out_full)
        assert not got[:, :, 1].any()          # slot n = 1 (the cross-attention's) is left alone


@pytest.mark.parametrize("grouped", [False, True])
def test_kernel_writes_only_the_rows_its_ranges_cover(F, grouped):
    """Write canaries (the pool has no compute-sanitizer): the destination is pre-filled with a sentinel and sits
    between guard rows; ragged ranges that end inside a 128-row tile, start off a tile boundary and leave gaps must
    change exactly the rows they cover - with the 32-byte epilogue stores, in the plain layout and in the grouped
    out-projection-input layout (where the other half of every group belongs to the cross-attention)."""
    Tq, Hq, Hkv = 1000, 8, 2
    q, kv = bf(rnd((Tq, Hq, 128), 1)), bf(rnd((700, Hkv, 256), 2))
    qr_l, kr_l = [[3, 130], [131, 389], [600, 857]], [[0, 200], [100, 700], [5, 133]]
    qr, kr = i32(qr_l), i32(kr_l)
    ref, _ = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, return_lse=False, zero_fill=True)
    covered = torch.zeros(Tq, dtype=torch.bool, device="cuda")
    for a, b in qr_l:
        covered[a:b] = True
    sent = 7.0
    if not grouped:
        big = torch.full((Tq + 64, Hq, 128), sent, dtype=torch.bfloat16, device="cuda")
        out = big[32:32 + Tq]
        F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False)
        assert torch.equal(out[covered], ref[covered])
        assert bool((out[~covered] == sent).all()) and bool((big[:32] == sent).all()) and bool((big[32 + Tq:] == sent).all())
    else:
        big = torch.full((Tq + 64, 2 * Hq * 128), sent, dtype=torch.bfloat16, device="cuda")
        buf = big[32:32 + Tq]
        F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, return_lse=False, proj_input=(buf, 0, 8))
        v = buf.view(Tq, 8, 2, Hq // 8, 128)
        assert torch.equal(v[:, :, 0].reshape(Tq, Hq, 128)[covered], ref[covered])
        assert bool((v[:, :, 0][~covered] == sent).all()) and bool((v[:, :, 1] == sent).all())
        assert bool((big[:32] == sent).all()) and bool((big[32 + Tq:] == sent).all())


def test_dynamic_scheduler_is_grid_independent_and_rearms_its_counters(F):
    """The atomic work counter lives in the plan workspace and the last CTA of a launch re-arms it: the same plan used
    by many launches in a row, with grids of 1, 3 and all SMs (one CTA walking every item through the two-slot mailbox;
    more CTAs than items at the other end), and both mailbox flavours (index / decoded item, chosen by the key count),
    must give the same bits every time."""
    Hq, Hkv = 6, 2
    rng = np.random.default_rng(5)
    for total_k in (900, 9000):                  # <= 8192 keys: decoded mailbox; above: index mailbox
        Tq = 2100
        q, kv = bf(rnd((Tq, Hq, 128), 1)), bf(rnd((total_k, Hkv, 256), 2))
        cuts = np.sort(rng.choice(np.arange(1, Tq), size=11, replace=False))
        edges = np.concatenate([[0], cuts, [Tq]])
        qr_l = [[int(a), int(b)] for a, b in zip(edges[:-1], edges[1:])]
        kr_l = []
        for _ in qr_l:
            a = int(rng.integers(0, total_k - 1))
            kr_l.append([a, int(rng.integers(a + 1, total_k + 1))])
        qr, kr = i32(qr_l), i32(kr_l)
        plan = F.attn_plan(qr, kr, num_heads_q=Hq, total_q=Tq, total_k=total_k)
        ref, ref_lse = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)       # per-call work list, full grid
        for sms in (1, 3, 0, 3, 1, 0, 0):
            out, lse = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, plan=plan, num_sms=sms)
            assert torch.equal(out, ref) and torch.equal(lse, ref_lse), (total_k, sms)


def test_output_rows_that_are_only_16_byte_aligned(F):
    """The epilogue writes 32-byte vectors when every output row is 32-byte aligned (the normal case) and falls back to
    16-byte stores otherwise: a destination whose token stride is a multiple of 8 but not of 16 elements, and one whose
    base is offset by 16 bytes, must give the same bits as the aligned destination."""
    Tq, Hq, Hkv = 300, 4, 2
    q, kv = bf(rnd((Tq, Hq, 128), 1)), bf(rnd((500, Hkv, 256), 2))
    qr, kr = i32([[0, 130], [130, 300]]), i32([[0, 200], [100, 500]])
    ref, _ = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, return_lse=False)
    wide = torch.zeros(Tq, Hq * 128 + 8, dtype=torch.bfloat16, device="cuda")           # token stride 520: 16-byte rows
    out = wide[:, :Hq * 128].unflatten(1, (Hq, 128))
    F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False)
    assert torch.equal(out, ref) and not wide[:, Hq * 128:].any()
    flat = torch.zeros(Tq * Hq * 128 + 8, dtype=torch.bfloat16, device="cuda")           # base offset by 16 bytes
    out2 = flat[8:].view(Tq, Hq, 128)
    F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out2, return_lse=False)
    assert torch.equal(out2, ref) and not flat[:8].any()


def test_overlapped_launch_gives_the_same_bits(F):
    """magi_attn_args.launch_dependency = 1: the cross-attention launch that may start under the tail of the
    self-attention launch before it (programmatic dependent launch, wait moved to the end of the kernel) - same bits
    as the serialised launch, for both launches, repeatedly, and the launch after the pair sees both finished."""
    clip, Hq, Hkv = 1170, 3, 1                      # per-rank shapes of cp_ulysses with 8 ranks, scaled down
    S = 4 * clip
    q, kv = bf(rnd((S, Hq, 128), 1)), bf(rnd((S, Hkv, 256), 2))
    qr = i32([[i * clip, (i + 1) * clip] for i in range(4)])
    kr = i32([[0, (i + 1) * clip] for i in range(4)])
    xq, xkv = bf(rnd((S // 8, 24, 128), 3)), bf(rnd((300, 8, 256), 4))
    xqr, xkr = i32([[0, 200], [200, S // 8]]), i32([[0, 120], [120, 300]])
    plan = F.attn_plan(qr, kr, num_heads_q=Hq, total_q=S, total_k=S)
    xplan = F.attn_plan(xqr, xkr, num_heads_q=24, total_q=S // 8, total_k=300)
    ref, _ = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, return_lse=False)
    xref, _ = F.attn_ranges(xq, xkv[..., :128], xkv[..., 128:], xqr, xkr, return_lse=False)
    torch.cuda.synchronize()
    with pytest.raises(ValueError):                  # needs a plan and a caller-provided destination
        F.attn_ranges(xq, xkv[..., :128], xkv[..., 128:], xqr, xkr, return_lse=False, independent_of_prev=True)
    for _ in range(5):
        out, xout = torch.full_like(ref, float("nan")), torch.full_like(xref, float("nan"))
        F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False, plan=plan)
        F.attn_ranges(xq, xkv[..., :128], xkv[..., 128:], xqr, xkr, out=xout, return_lse=False, plan=xplan,
                      independent_of_prev=True)
        both = out.float().sum() + xout.float().sum()        # a plain launch after the pair: must see both complete
        assert torch.equal(out, ref) and torch.equal(xout, xref)
        assert torch.isfinite(both)


@pytest.mark.parametrize("cp,clip", [(2, 150), (3, 100), (8, 333)])
def test_fused_kv_gather_on_virtual_ranks(F, cp, clip):
    """cp_ring_kv with the K/V all-gather fused into the prologue's stores, exercised on ONE device: every "rank" owns
    a stripe of every chunk and a gathered K/V buffer whose address the kernels get as a peer pointer. After all
    ranks' prologues every buffer must hold exactly the single-rank packed K/V (chunk-major), rows outside the stripes
    untouched; the attention of every rank's query stripe against it, with the unmodified k ranges, must be
    bit-identical to the single-rank rows (same kernel, same per-row key order)."""
    from magi_b200.context_parallel import ring_split_sizes
    dn, Hq, Hkv, sp = 3, 6, 2, 1
    S = clip * dn
    sizes = ring_split_sizes(clip, cp)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    xq, xk, xv = bf(rnd((S, Hq * 128), 1)), bf(rnd((S, Hkv * 128), 2)), bf(rnd((S, Hkv * 128), 3))
    cache = bf(rnd((sp * clip, Hkv, 256), 4))
    ln = [torch.from_numpy(rnd((128,), 10 + i, 0.2)).cuda() for i in range(4)]
    ang = torch.from_numpy(rnd((S, 48), 20, 3.0))
    sin, cos = torch.sin(ang).cuda(), torch.cos(ang).cuda()
    q_ranges = i32([[i * clip, (i + 1) * clip] for i in range(dn)])
    k_ranges = i32([[0, 2 * clip], [clip, 3 * clip], [0, 4 * clip]])       # absolute rows: cache prefix ++ in-flight
    q_full, kv_full = F.qk_norm_rope(xq, xk, xv, Hq, Hkv, (ln[0], ln[1]), (ln[2], ln[3]), sin, cos)
    out_full, _ = F.attn_ranges(q_full, kv_full[..., :128], kv_full[..., 128:], q_ranges, k_ranges,
                                k_cache=cache[..., :128], v_cache=cache[..., 128:])
    pad = 5                                                                   # guard rows around every buffer
    bufs = [torch.full((S + 2 * pad, Hkv, 256), 7.0, dtype=torch.bfloat16, device="cuda") for _ in range(cp)]
    stripe = lambda t, r: t.unflatten(0, (dn, clip))[:, int(offs[r]):int(offs[r + 1])].flatten(0, 1).contiguous()
    qs = []
    for me in range(cp):
        q_me, kv_me = F.qk_norm_rope(stripe(xq, me), stripe(xk, me), stripe(xv, me), Hq, Hkv, (ln[0], ln[1]),
                                     (ln[2], ln[3]), stripe(sin, me), stripe(cos, me),
                                     gather_kv_out=[b[pad:].data_ptr() for b in bufs],
                                     gather_geometry=(sizes[me], clip, int(offs[me])))
        assert kv_me is None
        qs.append(q_me)
    for b in bufs:
        assert torch.equal(b[pad:pad + S], kv_full)
        assert bool((b[:pad] == 7.0).all()) and bool((b[pad + S:] == 7.0).all())
    for me in range(cp):
        n = sizes[me]
        qr = i32([[i * n, (i + 1) * n] for i in range(dn)])
        kv = bufs[me][pad:pad + S]
        o, _ = F.attn_ranges(qs[me], kv[..., :128], kv[..., 128:], qr, k_ranges, k_cache=cache[..., :128],
                             v_cache=cache[..., 128:])
        assert torch.equal(o, stripe(out_full, me))
    with pytest.raises(ValueError):
        F.qk_norm_rope(stripe(xq, 0), stripe(xk, 0), stripe(xv, 0), Hq, Hkv, (ln[0], ln[1]), (ln[2], ln[3]),
                       stripe(sin, 0), stripe(cos, 0), gather_kv_out=[bufs[0].data_ptr()],
                       gather_geometry=(sizes[0], sizes[0] - 1, 0))          # a stripe larger than a chunk


def test_cso_unshuffle_and_full_attention(F):
    """CSO pieces with the real kernels: the receive-side gather "(cp dn spb) -> dn (cp spb)"[:, :clip]
    (dit_module.py:1215-1224) is bit exact and can land in the KV-cache slot; full_attention on the padded query
    chunk (dit_module.py:1068-1096) matches the oracle's cso_core_attention."""
    cp, dn, clip, Hq, Hkv = 4, 3, 203, 4, 2
    spb = -(-clip // cp)
    recv = bf(rnd((cp * dn * spb, Hkv, 256), 41))
    exp = recv.view(cp, dn, spb, Hkv, 256).permute(1, 0, 2, 3, 4).reshape(dn, cp * spb, Hkv, 256)[:, :clip].flatten(0, 1)
    got = F.cso_unshuffle(recv, cp, dn, clip)
    assert torch.equal(got, exp)
    cache = torch.zeros(5 * clip, Hkv, 256, dtype=torch.bfloat16, device="cuda")
    F.cso_unshuffle(recv, cp, dn, clip, out=cache[clip:clip + dn * clip])
    assert torch.equal(cache[clip:clip + dn * clip], exp) and not cache[:clip].any() and not cache[(dn + 1) * clip:].any()
    # per-rank slices of every padded chunk -> oracle (all ranks simulated) vs kernel (one launch per chunk)
    qs = [rnd((dn * spb, Hq, 128), 50 + r) for r in range(cp)]
    kvs = [rnd((dn * spb, Hkv, 256), 60 + r) for r in range(cp)]
    k_range = np.array([[0, clip], [0, 2 * clip], [clip, 3 * clip]], dtype=np.int32)
    ref = O.cso_core_attention(qs, kvs, dn, clip, k_range)
    kv_full = np.concatenate([np.concatenate([kvs[r][d * spb:(d + 1) * spb] for r in range(cp)], 0)[:clip]
                              for d in range(dn)], 0)
    kv_dev = bf(kv_full)
    for d in range(dn):
        q = bf(np.concatenate([qs[r][d * spb:(d + 1) * spb] for r in range(cp)], 0))
        out, _ = F.attn_ranges(q, kv_dev[..., :128], kv_dev[..., 128:], i32([[0, cp * spb]]), i32(k_range[d:d + 1]))
        exp_o = np.concatenate([ref[r][d * spb:(d + 1) * spb] for r in range(cp)], 0)
        check_attn(out, exp_o, f"cso chunk {d}")


@pytest.mark.parametrize("tag", ["attn_b1", "mlp_b1", "attn_b2"])
def test_bias_modulate_add_vs_reference_golden(F, tag):
    """Fused adaLN gating chain kernel against the reference's bias_modulate_add output (tests/golden/gating.npz)
    and the oracle: <= 1 bf16 ulp (fp32 LayerNorm summation order differs)."""
    g = Golden("gating.npz")
    x = g.torch(f"{tag}.x", "cuda")
    res, gate = g.torch(f"{tag}.residual", "cuda"), g.torch(f"{tag}.gate", "cuda")
    cmap = g.torch(f"{tag}.cmap", "cuda")
    w, b = g.torch(f"{tag}.w", "cuda"), g.torch(f"{tag}.b", "cuda")
    out = F.bias_modulate_add(x, res, cmap, gate, (w, b, 1e-6, True))
    assert out.dtype == torch.bfloat16 and out.shape == x.shape
    ref = g[f"{tag}.out"]
    got = out.float().cpu().numpy()
    tol = np.maximum(np.abs(ref), np.abs(got)) * 2.0 ** -7 + 1e-6
    assert (np.abs(got - ref) <= tol).all(), float(np.abs(got - ref).max())
    # int32 map, fp32 gate: same result
    out2 = F.bias_modulate_add(x, res, cmap.to(torch.int32), gate.float(), (w, b, 1e-6, True))
    assert torch.equal(out, out2)


@pytest.mark.parametrize("hidden,rows,xdt", [(3072, 1000, torch.float32), (6144, 517, torch.float32),
                                             (3072, 333, torch.bfloat16), (256, 77, torch.float32),
                                             (12288, 64, torch.bfloat16)])
def test_bias_modulate_add_sizes_vs_oracle(F, hidden, rows, xdt):
    """Model widths (4.5B: 3072 = one warp per row, 24B: 6144 = two warps per row) and the extremes of the supported
    range against the oracle."""
    gen = torch.Generator().manual_seed(hidden + rows)
    x = (torch.randn(rows, 1, hidden, generator=gen) * 2).to(xdt)
    res = torch.randn(rows, 1, hidden, generator=gen).to(torch.bfloat16)
    gate = torch.tanh(torch.randn(1, 5, hidden, generator=gen)).to(torch.bfloat16)
    cmap = torch.randint(0, 5, (rows, 1), generator=gen)
    w, b = torch.randn(hidden, generator=gen) * 0.1, torch.randn(hidden, generator=gen) * 0.1
    ref = O.bias_modulate_add(x.float().numpy(), res.float().numpy(), cmap.numpy(), gate.float().numpy(), w.numpy(),
                              b.numpy())
    out = F.bias_modulate_add(x.cuda(), res.cuda(), cmap.cuda(), gate.cuda(), (w.cuda(), b.cuda(), 1e-6, True))
    got = out.float().cpu().numpy()
    tol = np.maximum(np.abs(ref), np.abs(got)) * 2.0 ** -7 + 1e-6
    assert (np.abs(got - ref) <= tol).all(), float(np.abs(got - ref).max())


@pytest.mark.parametrize("gated,rows,f", [(True, 1000, 16384), (False, 777, 12288), (True, 3, 8), (False, 1, 4096)])
def test_mlp_activation_vs_oracle_and_libraries(F, gated, rows, f):
    """magi_mlp_activation (fused SwiGLU / GELU between fc1 and fc2, dit_module.py:545-548) vs the oracle (<= 1 bf16
    ulp: fast-exp vs exp), and against what the reference itself calls: flashinfer.activation.silu_and_mul (when the
    wheel imports) and torch.nn.functional.gelu."""
    x = rnd((rows, (2 if gated else 1) * f), 77, scale=2.5)
    ref = O.mlp_activation(x, gated)
    tx = bf(x)
    out = F.mlp_activation(tx, gated)
    assert out.shape == (rows, f) and out.dtype == torch.bfloat16
    got = out.float().cpu().numpy()
    # <= 1 bf16 ulp; 1 + erf(x / sqrt 2) cancels for x << 0, where implementations differ in the last bit more often
    assert np.all(np.abs(got - ref) <= np.abs(ref) * 2.0 ** -6 + 1e-5)
    assert (got != ref).mean() < (0.01 if gated else 0.05)
    if gated:
        try:
            import flashinfer
            lib = flashinfer.activation.silu_and_mul(tx)
        except Exception:
            lib = None
    else:
        lib = torch.nn.functional.gelu(tx)
    if lib is not None:
        d = (out.float() - lib.float()).abs()
        assert bool((d <= lib.float().abs() * 2.0 ** -6 + 1e-5).all())
        assert float((d > 0).float().mean()) < (0.01 if gated else 0.05)
    # strided input rows (a view into a wider buffer) and a caller-provided output
    wide = torch.zeros(rows, tx.shape[1] + 64, dtype=torch.bfloat16, device="cuda")
    wide[:, :tx.shape[1]] = tx
    out2 = torch.empty(rows, f, dtype=torch.bfloat16, device="cuda")
    F.mlp_activation(wide[:, :tx.shape[1]], gated, out=out2)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("cp,Hq,Hkv,fp8", [(2, 4, 2, False), (4, 8, 2, False), (8, 8, 1, False), (2, 4, 2, True),
                                            (8, 8, 1, True)])
def test_fused_projection_gemm_peer_routing_on_virtual_ranks(F, cp, Hq, Hkv, fp8):
    """The Ulysses input all-to-all fused into the projection GEMM's epilogue (magi_qkv_proj_args.peer_*), on ONE device:
    every "rank" projects its own rows and stores each head slice into the owning rank's receive buffer. What arrives
    must be, bit for bit, the head slices of the single-rank fused GEMM over all rows (incl. KV-head replication when
    cp > Hkv); the cross-attention query stays local. fp8: the same through the fp8_quant instantiation (e4m3 operands,
    per-projection output scale)."""
    S, hidden = 700, 256
    split = [S // cp + (1 if i < S % cp else 0) for i in range(cp)]
    offs = np.concatenate([[0], np.cumsum(split)]).astype(np.int64)
    x = bf(rnd((S, hidden), 1))
    ws = [bf(rnd((h * 128, hidden), 2 + i, hidden ** -0.5)) for i, h in enumerate((Hq, Hkv, Hkv, Hq))]
    kw = {}
    if fp8:
        x = F.fp8_quantize(x, torch.full((hidden,), 4.0 / 448.0, device="cuda"))
        ws = [(w_.float() * 4000.0).clamp(-448.0, 448.0).to(torch.float8_e4m3fn) for w_ in ws]
        kw = dict(out_scale=torch.tensor([2.1e-6, 2.3e-6, 2.2e-6, 2.4e-6], device="cuda"))
    w = F.stack_qkv_weights(*ws)
    ln = [torch.from_numpy(rnd((128,), 10 + i, 0.2)).cuda() for i in range(4)]
    lnx = (bf(rnd((128,), 20, 0.2)), bf(rnd((128,), 21, 0.2)))
    ang = torch.from_numpy(rnd((S, 48), 30, 3.0))
    sin, cos = torch.sin(ang).cuda(), torch.cos(ang).cuda()
    q_full, kv_full, qx_full = F.qkv_proj_norm_rope(x, w, Hq, Hkv, (ln[0], ln[1]), (ln[2], ln[3]), sin, cos, ln_qx=lnx,
                                                    **kw)
    assert float(q_full.float().abs().mean()) > 0.1 and float(kv_full[..., 128:].float().abs().mean()) > 0.01
    hn, hkv_loc = Hq // cp, max(1, Hkv // cp)
    recv_q = [torch.zeros(S, hn, 128, dtype=torch.bfloat16, device="cuda") for _ in range(cp)]
    recv_kv = [torch.zeros(S, hkv_loc, 256, dtype=torch.bfloat16, device="cuda") for _ in range(cp)]
    for me in range(cp):
        sl = slice(int(offs[me]), int(offs[me + 1]))
        q, kv, qx = F.qkv_proj_norm_rope(x[sl], w, Hq, Hkv, (ln[0], ln[1]), (ln[2], ln[3]), sin[sl], cos[sl], ln_qx=lnx,
                                         peer_q_out=[t.data_ptr() for t in recv_q],
                                         peer_kv_out=[t.data_ptr() for t in recv_kv], peer_row_offset=int(offs[me]),
                                         **kw)
        assert q is None and kv is None and torch.equal(qx, qx_full[sl])
    rep = cp // Hkv if cp > Hkv else 1
    for c in range(cp):
        assert torch.equal(recv_q[c], q_full[:, c * hn:(c + 1) * hn])
        k_lo = c // rep if cp > Hkv else c * hkv_loc
        assert torch.equal(recv_kv[c], kv_full[:, k_lo:k_lo + hkv_loc])


@pytest.mark.parametrize("rows,hidden,Hq,Hkv,with_qx,batch,fp32_ln", [
    (300, 256, 4, 2, True, 1, True), (1000, 3072, 6, 2, False, 1, True), (514, 512, 2, 1, True, 2, False)])
def test_fused_projection_gemm_vs_oracle(F, rows, hidden, Hq, Hkv, with_qx, batch, fp32_ln):
    """magi_qkv_proj_norm_rope_fwd (the q / k / v / qx projections with the QK-LayerNorm + RoPE prologue as the GEMM
    epilogue) against the oracle's composition of the reference steps: nn.Linear in bf16
    (dit_module.py:405-425) -> get_q / get_kv (:953-996) / xattn q LayerNorm (:1008), and against this repository's
    unfused path (cuBLAS GEMMs + magi_qk_norm_rope_fwd) on the same inputs. Tolerance: the prologue's 1 bf16 ulp plus
    the ulp flips of the bf16-rounded GEMM output that different fp32 accumulation orders cause (a few 1e-2 after the
    LayerNorm's 1 / sigma amplification on a handful of elements): max_abs 6e-2, rel-L2 6e-3."""
    x = rnd((rows, hidden), 3, scale=1.0)
    ws = [rnd((h * 128, hidden), 4 + i, scale=hidden ** -0.5) for i, h in enumerate((Hq, Hkv, Hkv, Hq))]
    if not with_qx:
        ws = ws[:3]
    lnp = [rnd((128,), 20 + i, scale=0.2) for i in range(6)]
    ang = np.random.default_rng(0).uniform(0, 6.28, size=(rows // batch, 48)).astype(np.float32)
    sin, cos = np.sin(ang), np.cos(ang)
    lin = [O.linear_bf16(x, w_) for w_ in ws]
    sq = rows // batch
    ref_q = O.get_q_or_k(lin[0].reshape(sq, batch, -1), Hq, lnp[0], lnp[1], sin, cos)
    ref_kv = O.get_kv(lin[1].reshape(sq, batch, -1), lin[2].reshape(sq, batch, -1), Hkv, lnp[2], lnp[3], sin, cos)
    dt = torch.float32 if fp32_ln else torch.bfloat16
    t = lambda a, d=dt: torch.from_numpy(a).cuda().to(d)
    w_cat = F.stack_qkv_weights(*[bf(w_) for w_ in ws])
    cache = torch.zeros(rows + 50, Hkv, 256, dtype=torch.bfloat16, device="cuda")
    q, kv, qx = F.qkv_proj_norm_rope(bf(x), w_cat, Hq, Hkv, (t(lnp[0]), t(lnp[1])), (t(lnp[2]), t(lnp[3])),
                                     torch.from_numpy(sin).cuda(), torch.from_numpy(cos).cuda(),
                                     ln_qx=(bf(lnp[4]), bf(lnp[5])) if with_qx else None, batch=batch,
                                     cache=cache, cache_row_offset=50, cache_rows=rows - 7)

    def close(a, ref, what):
        a = a.float().cpu().numpy()
        err = np.abs(a - ref).max()
        rel = np.linalg.norm(a - ref) / np.linalg.norm(ref)
        assert a.shape == ref.shape and err <= 6e-2 and rel <= 6e-3, f"{what}: max_abs {err:.3e} rel_l2 {rel:.3e}"

    if fp32_ln:   # the oracle's get_q_or_k models fp32 LayerNorm parameters
        close(q, ref_q, "q")
        close(kv, ref_kv, "kv")
    else:
        close(kv[..., 128:], ref_kv[..., 128:], "v")
    if with_qx:
        ref_qx = O.xattn_layer_norm(lin[3].reshape(rows, Hq, 128), O.bf16_round(lnp[4]), O.bf16_round(lnp[5]))
        close(qx, ref_qx, "qx")
    assert torch.equal(cache[50:50 + rows - 7], kv[:rows - 7]) and float(cache[50 + rows - 7:].abs().sum()) == 0.0
    # the unfused path of this repository on the same inputs
    xt = bf(x)
    y = [torch.nn.functional.linear(xt, bf(w_)) for w_ in ws]
    q2, kv2 = F.qk_norm_rope(y[0], y[1], y[2], Hq, Hkv, (t(lnp[0]), t(lnp[1])), (t(lnp[2]), t(lnp[3])),
                             torch.from_numpy(sin).cuda(), torch.from_numpy(cos).cuda(), batch=batch)
    close(q, q2.float().cpu().numpy(), "q vs unfused")
    close(kv, kv2.float().cpu().numpy(), "kv vs unfused")


def test_full_attention_block_matches_reference_golden():
    """The drop-in FullyParallelAttention replays the reference's own forward sequence (steps A-E of
    oracle/gen_golden.py: KV-cache fill, fwd_extra_1st_chunk, update_kv_cache False/True, distill extra range,
    batched uncond pass) and must reproduce core_attn_out, xattn_out and the cache contents."""
    from magi_b200.common import InferenceParams, ModelMetaArgs, PackedCoreAttnParams, PackedCrossAttnParams
    g = Golden("attention_single_rank.npz")
    attn = _build_module(g)
    clip = int(g["clip"])
    inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
    inf.key_value_memory_dict[1] = torch.zeros(4 * clip, attn.model_config.num_query_groups, 256, dtype=torch.bfloat16,
                                               device="cuda")
    for s in ["A", "B", "C", "D", "E"]:
        sp, dn, bs, upd, extra, ncc, has_inf = [int(x) for x in g[f"{s}.meta"]]
        p = O.cross_attn_params(clip, dn, bs, g[f"{s}.y_index"])
        q_range = g.raw(f"{s}.q_range")
        k_range = g.raw(f"{s}.k_range")
        core = PackedCoreAttnParams(q_range=i32(q_range), k_range=i32(k_range), np_q_range=q_range, np_k_range=k_range,
                                    max_seqlen_q=clip, max_seqlen_k=O.max_seqlen_k(k_range))
        cross = PackedCrossAttnParams(q_ranges=i32(p["q_ranges"]), kv_ranges=i32(p["kv_ranges"]),
                                      cu_seqlens_q=i32(p["cu_seqlens_q"]), cu_seqlens_kv=i32(p["cu_seqlens_kv"]),
                                      max_seqlen_q=clip, max_seqlen_kv=800)
        meta = ModelMetaArgs(H=4, W=4, cp_pad_size=None, cp_split_sizes=None, slice_point=sp, denoising_range_num=dn,
                             range_num=sp + dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=bool(extra),
                             distill_nearly_clean_chunk=bool(ncc), clip_token_nums=clip, enable_cuda_graph=False,
                             core_attn_params=core, cross_attn_params=cross)
        params = inf if has_inf else None
        if params is not None:
            params.update_kv_cache = bool(upd)
        core_out, x_out = attn(g.torch(f"{s}.hidden_states", "cuda"), g.torch(f"{s}.key_value_states", "cuda"), params,
                               g.torch(f"{s}.rope", "cuda"), meta)
        ref_core, ref_x = g[f"{s}.core_attn_out"], g[f"{s}.xattn_out"]
        assert tuple(core_out.shape) == ref_core.shape and tuple(x_out.shape) == ref_x.shape
        # cuBLAS vs ATen-CPU GEMMs feed both paths: tolerance is the attention tolerance of the module docstring
        check_attn(core_out, ref_core, f"step {s} core_attn_out")
        check_attn(x_out, ref_x, f"step {s} xattn_out")
        if f"{s}.cache_after" in g:
            cache = inf.key_value_memory_dict[1].float().cpu().numpy()
            ref_cache = g[f"{s}.cache_after"]
            written = np.abs(ref_cache).sum(axis=(1, 2)) > 0
            assert np.array_equal(np.abs(cache).sum(axis=(1, 2)) > 0, written), f"step {s}: cache rows written"
            # K goes through a cuBLAS GEMM + LayerNorm + RoPE: compare within a few bf16 ulps; V likewise via GEMM
            err = np.abs(cache - ref_cache)
            tol = np.maximum(np.abs(cache), np.abs(ref_cache)) * 2.0 ** -7 * 4 + 2e-2
            assert (err <= tol).all(), f"step {s}: cache contents, max err {err.max()}"
        if s == "B":
            assert x_out[:clip].abs().sum() == 0   # zero caption tokens -> zeros (dit_module.py:1124 semantics)


def test_transformer_layer_matches_reference_golden():
    """The drop-in TransformerLayer (attention hot path + fused gating chains + torch GEMMs) against the reference's own
    TransformerLayer.forward (tests/golden/transformer_layer.npz): reference state_dict loads by name, two forwards
    (no cache, then one cached chunk). Tolerance: SURVEY.md section 8(c) per-step latent criterion, rel-L2 <= 2e-2
    (cuBLAS vs ATen-CPU bf16 GEMMs)."""
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    from magi_b200.transformer_layer import TransformerLayer, high_precision_promote
    g = Golden("transformer_layer.npz")
    hidden = g.raw("w.self_attention.linear_qkv.layer_norm.weight").shape[0]
    Hq, Hkv = [int(x) for x in g["heads"]]
    mc = ModelConfig(num_layers=4, hidden_size=hidden, ffn_hidden_size=4 * hidden, num_attention_heads=Hq,
                     num_query_groups=Hkv)
    layer = high_precision_promote(TransformerLayer(mc, EngineConfig(), layer_number=0))
    sd = {k[2:]: g.torch(k) for k in g.files if k.startswith("w.")}
    layer.load_state_dict(sd, strict=True)
    layer = layer.cuda().eval()
    assert layer.self_attn_post_norm.weight.dtype == torch.float32
    clip = int(g["clip"])
    inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
    # as in the fixture: step A (slice_point 0, no extra chunk) does not touch the cache, step B reads chunk 0 of it
    inf.key_value_memory_dict[0] = torch.zeros(4 * clip, Hkv, 256, dtype=torch.bfloat16, device="cuda")
    for tag in ("A", "B"):
        sp, dn, upd = [int(x) for x in g[f"{tag}.meta"]]
        q_range, k_range = O.core_q_ranges(clip, dn, 1), g.raw(f"{tag}.k_range")
        p = O.cross_attn_params(clip, dn, 1, g[f"{tag}.y_index"])
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
        out = layer(g.torch(f"{tag}.hidden_states", "cuda"), g.torch(f"{tag}.condition", "cuda"),
                    g.torch(f"{tag}.condition_map", "cuda"), g.torch(f"{tag}.key_value_states", "cuda"),
                    g.torch(f"{tag}.rope", "cuda"), inf, meta)
        ref = g[f"{tag}.out"]
        got = out.float().cpu().numpy()
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert got.shape == ref.shape and rel < 2e-2, (tag, rel)


def test_kv_offload_matches_resident_cache():
    """kv_offload (dit_module.py:803-811, 914-929): caches in pinned host memory, prefix staged to the device (the next
    layer's prefix prefetched on a copy stream), new rows written back asynchronously. Two layers replay the golden
    step sequence with and without offload: outputs and cache contents must be bit-identical."""
    from magi_b200.attention import KVOffload
    from magi_b200.common import InferenceParams, ModelMetaArgs, PackedCoreAttnParams, PackedCrossAttnParams
    g = Golden("attention_single_rank.npz")
    clip = int(g["clip"])

    def run(offload):
        layers = []
        for ln_ in (1, 2):
            m = _build_module(g)
            m.layer_number = ln_
            m.engine_config.kv_offload = offload
            layers.append(m)
        inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
        outs = []
        for rep in range(2):                      # second pass: every layer's cache exists -> prefetch path
            for s in ["A", "B", "C", "D"]:
                sp, dn, bs, upd, extra, ncc, has_inf = [int(x) for x in g[f"{s}.meta"]]
                if not has_inf or bs != 1:
                    continue
                p = O.cross_attn_params(clip, dn, bs, g[f"{s}.y_index"])
                q_range, k_range = g.raw(f"{s}.q_range"), g.raw(f"{s}.k_range")
                core = PackedCoreAttnParams(q_range=i32(q_range), k_range=i32(k_range), np_q_range=q_range,
                                            np_k_range=k_range, max_seqlen_q=clip, max_seqlen_k=O.max_seqlen_k(k_range))
                cross = PackedCrossAttnParams(q_ranges=i32(p["q_ranges"]), kv_ranges=i32(p["kv_ranges"]),
                                              cu_seqlens_q=i32(p["cu_seqlens_q"]), cu_seqlens_kv=i32(p["cu_seqlens_kv"]),
                                              max_seqlen_q=clip, max_seqlen_kv=800)
                meta = ModelMetaArgs(H=4, W=4, cp_pad_size=None, cp_split_sizes=None, slice_point=sp,
                                     denoising_range_num=dn, range_num=sp + dn, extract_prefix_video_feature=False,
                                     fwd_extra_1st_chunk=bool(extra), distill_nearly_clean_chunk=bool(ncc),
                                     clip_token_nums=clip, enable_cuda_graph=False, core_attn_params=core,
                                     cross_attn_params=cross)
                inf.update_kv_cache = bool(upd)
                hs = g.torch(f"{s}.hidden_states", "cuda")
                for m in layers:
                    c, x = m(hs, g.torch(f"{s}.key_value_states", "cuda"), inf, g.torch(f"{s}.rope", "cuda"), meta)
                    outs.append((c.clone(), x.clone()))
                    hs = (hs.float() + 0.1 * c.float().mean(-1, keepdim=True)).to(torch.bfloat16)   # layer 2 sees other data
        torch.cuda.synchronize()
        if offload:
            KVOffload.of(inf, torch.device("cuda", 0)).synchronize()
        return outs, inf

    ref_outs, ref_inf = run(False)
    off_outs, off_inf = run(True)
    for (c0, x0), (c1, x1) in zip(ref_outs, off_outs):
        assert torch.equal(c0, c1) and torch.equal(x0, x1)
    st = off_inf.kv_offload_state
    assert st.prefetch_hits > 0 and st.d2h_bytes > 0 and st.h2d_bytes > 0
    for ln_ in (1, 2):
        host = off_inf.key_value_memory_dict[ln_]
        assert not host.is_cuda and host.is_pinned()
        dev_cache = ref_inf.key_value_memory_dict[ln_]
        rows = 3 * clip          # rows written by the sequence (the rest of torch.empty storage is undefined)
        assert torch.equal(host[:rows].cuda(), dev_cache[:rows])


def test_reference_named_pieces_match_golden():
    """get_q / get_kv / get_qkv / get_xqkv / adjust_key_and_value_for_inference / full_attention keep the reference's
    names and argument meaning (dit_module.py:933-1096); their outputs are compared with what the reference's own
    methods produced on the same weights and inputs (golden step C: slice_point 1, cache from steps A-B)."""
    from magi_b200.common import InferenceParams, ModelMetaArgs, PackedCoreAttnParams
    g = Golden("attention_single_rank.npz")
    attn = _build_module(g)
    clip = int(g["clip"])
    s = "C"
    hs = g.torch(f"{s}.hidden_states", "cuda")
    rope = g.torch(f"{s}.rope", "cuda")
    sin_emb, cos_emb = rope.tensor_split(2, -1)
    mixed = attn.linear_qkv.forward_ln(hs)

    def close(a, ref, ulps=6.0, atol=3e-2):   # cuBLAS vs ATen-CPU GEMM feeding a LayerNorm: a few bf16 ulps
        a = a.float().cpu().numpy()
        tol = np.maximum(np.abs(a), np.abs(ref)) * 2.0 ** -7 * ulps + atol
        assert a.shape == ref.shape and (np.abs(a - ref) <= tol).all(), float(np.abs(a - ref).max())

    q = attn.get_q(mixed, cos_emb, sin_emb)
    kv = attn.get_kv(mixed, cos_emb, sin_emb)
    close(q, g[f"{s}.q"])
    close(kv, g[f"{s}.kv"])
    q2, k2, v2 = attn.get_qkv(mixed, cos_emb, sin_emb)
    assert torch.equal(q2, q) and torch.equal(k2, kv[..., :128]) and torch.equal(v2, kv[..., 128:])
    assert torch.equal(attn.get_k(mixed, cos_emb, sin_emb), kv[..., :128])
    assert torch.equal(attn.get_v(mixed), kv[..., 128:])
    xq, xk, xv = attn.get_xqkv(mixed, g.torch(f"{s}.key_value_states", "cuda"))
    close(xq, g[f"{s}.xq"])
    close(xk, g[f"{s}.xk"])
    close(xv, g[f"{s}.xv"], ulps=2.0, atol=2e-2)
    # KV cache: replay the reference's cache contents of step B, then the reference's adjust_key_and_value call of C
    sp, dn = 1, 2
    inf = InferenceParams(1, 4 * clip)
    inf.key_value_memory_dict[1] = g.torch("B.cache_after", "cuda").clone()
    inf.update_kv_cache = False
    k_range, q_range = g.raw(f"{s}.k_range"), g.raw(f"{s}.q_range")
    meta = ModelMetaArgs(H=4, W=4, cp_pad_size=None, cp_split_sizes=None, slice_point=sp, denoising_range_num=dn,
                         range_num=sp + dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=False,
                         distill_nearly_clean_chunk=False, clip_token_nums=clip, enable_cuda_graph=False,
                         core_attn_params=PackedCoreAttnParams(i32(q_range), i32(k_range), q_range, k_range, clip,
                                                               O.max_seqlen_k(k_range)), cross_attn_params=None)
    kv_ref = g.torch(f"{s}.kv", "cuda")                       # the reference's own new-token K|V
    key, value = attn.adjust_key_and_value_for_inference(kv_ref, inf, meta)
    assert key.shape[0] == sp * clip + dn * clip
    assert torch.equal(key[:sp * clip], inf.key_value_memory_dict[1][:sp * clip, :, :128])
    assert torch.equal(value[sp * clip:], kv_ref[..., 128:])
    assert torch.equal(inf.key_value_memory_dict[1], g.torch("C.cache_after", "cuda"))   # update_kv_cache False
    # full_attention(i) on the reference's q / materialised k, v == rows of the reference's core_attn_out
    q_ref = g.torch(f"{s}.q", "cuda")
    ref_core = g[f"{s}.core_attn_out"].reshape(dn * clip, -1, 128)
    for i in range(dn):
        o = attn.full_attention(1, meta, q_ref[i * clip:(i + 1) * clip], key, value, i)
        check_attn(o, ref_core[i * clip:(i + 1) * clip], f"full_attention range {i}")


def test_fused_out_projection_input_layout():
    """both attention kernels write straight into the reference's out-projection input
    "sq b (n hn hd) -> sq b (hn n hd)", n = 2, hn = 8 (attn_linear_proj, dit_module.py:1332-1338). Must equal
    cat + rearrange of the two separate outputs bit for bit (batch 1 and the batched uncond pass)."""
    from magi_b200.attention import FullyParallelAttention
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    torch.manual_seed(0)
    Hq, Hkv, hidden, clip = 16, 8, 256, 150
    mc = ModelConfig(num_layers=2, hidden_size=hidden, num_attention_heads=Hq, num_query_groups=Hkv)
    attn = FullyParallelAttention(mc, EngineConfig(), layer_number=0)
    with torch.no_grad():
        for n_, p_ in attn.named_parameters():
            p_.copy_((torch.randn(p_.shape) * (0.2 if "layernorm" in n_ else p_.shape[-1] ** -0.5)).to(p_.dtype))
    for n_ in ("q_layernorm", "k_layernorm"):
        getattr(attn, n_).float()
    attn = attn.cuda().eval()
    for bs, dn in ((1, 3), (2, 1)):
        S = dn * clip
        y_index = [40, 77, 13][:dn * bs] if bs == 1 else [40, 77]
        p = O.cross_attn_params(clip, dn, bs, y_index)
        q_range = O.core_q_ranges(clip, dn, bs)
        k_range = np.array([[0, (i + 1) * clip] for i in range(dn)], dtype=np.int32) if bs == 1 else q_range[:1].copy()
        core = PackedCoreAttnParams(i32(q_range), i32(k_range), q_range, k_range, clip, O.max_seqlen_k(k_range))
        cross = PackedCrossAttnParams(i32(p["q_ranges"]), i32(p["kv_ranges"]), i32(p["cu_seqlens_q"]),
                                      i32(p["cu_seqlens_kv"]), clip, 800)
        meta = ModelMetaArgs(H=1, W=1, cp_pad_size=None, cp_split_sizes=None, slice_point=0, denoising_range_num=dn,
                             range_num=dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=False,
                             distill_nearly_clean_chunk=False, clip_token_nums=clip, enable_cuda_graph=False,
                             core_attn_params=core, cross_attn_params=cross)
        hs = torch.randn(S, bs, hidden, device="cuda").to(torch.bfloat16)
        kvs = torch.randn(sum(y_index), hidden, device="cuda").to(torch.bfloat16)
        ang = torch.rand(S, 48, device="cuda") * 6.28
        rope = torch.cat([torch.sin(ang), torch.cos(ang)], -1)
        inf = InferenceParams(1, 4 * clip) if bs == 1 else None
        core_out, x_out = attn(hs, kvs, inf, rope, meta)
        fused = attn(hs, kvs, inf, rope, meta, fused_proj_input=True)
        ref = FullyParallelAttention.attn_linear_proj_input(core_out, x_out)
        # the reference's own expression of the same re-layout
        from einops import rearrange
        ref2 = rearrange(torch.cat([core_out, x_out], dim=2), "sq b (n hn hd) -> sq b (hn n hd)", n=2, hn=8)
        assert torch.equal(ref, ref2)
        assert fused.shape == ref.shape and torch.equal(fused, ref), f"bs={bs}"


def test_hot_path_is_cuda_graph_capturable(F):
    """include/magi_attn.h promises no host reads of the ranges, no device sync and no hidden state: the prologue +
    range attention sequence is captured into a CUDA graph once and replayed on new inputs (same buffers)."""
    S, Hq, Hkv = 700, 6, 2
    xq = torch.zeros(S, Hq * 128, device="cuda", dtype=torch.bfloat16)
    xk = torch.zeros(S, Hkv * 128, device="cuda", dtype=torch.bfloat16)
    xv = torch.zeros(S, Hkv * 128, device="cuda", dtype=torch.bfloat16)
    ln = [torch.randn(128, device="cuda") * 0.2 for _ in range(4)]
    ang = torch.rand(S, 48, device="cuda") * 6.28
    sin, cos = torch.sin(ang), torch.cos(ang)
    qr, kr = i32([[0, 350], [350, 700]]), i32([[0, 350], [100, 700]])
    cache = torch.zeros(S, Hkv, 256, device="cuda", dtype=torch.bfloat16)

    def run():
        q, kv = F.qk_norm_rope(xq, xk, xv, Hq, Hkv, (ln[0], ln[1]), (ln[2], ln[3]), sin, cos, cache=cache,
                               cache_row_offset=0, cache_rows=S)
        return F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)

    def fill(seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        for t in (xq, xk, xv):
            t.copy_(torch.randn(t.shape, device="cuda", generator=g).to(torch.bfloat16))

    fill(1)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        run()                                  # warm-up outside capture
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_g, lse_g = run()
    for seed in (2, 3):
        fill(seed)
        graph.replay()
        torch.cuda.synchronize()
        out_e, lse_e = run()                   # eager on the same inputs
        assert torch.equal(out_g, out_e) and torch.equal(lse_g, lse_e), f"replay {seed}"
        assert torch.equal(cache, torch.cat([cache[..., :128], xv.view(S, Hkv, 128)], -1))


# ------------------------------------------------------------------------------------------ full-size properties
CLIP = 9360   # 480x832, BASELINE.json configs[1]


def _cfg2_inputs(Hq=24, Hkv=8, chunks=4, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(chunks * CLIP, Hq, 128, device="cuda", dtype=torch.bfloat16, generator=g)
    kv = torch.randn(chunks * CLIP, Hkv, 256, device="cuda", dtype=torch.bfloat16, generator=g)
    qr = i32([[i * CLIP, (i + 1) * CLIP] for i in range(chunks)])
    kr = i32([[0, (i + 1) * CLIP] for i in range(chunks)])
    return q, kv, qr, kr


def test_fullsize_constant_v_gives_constant_rows(F):
    """softmax rows sum to one: with V[:, h, d] = c[h, d] every output row must equal c (to bf16 rounding of P)."""
    q, kv, qr, kr = _cfg2_inputs()
    c = torch.randn(8, 128, device="cuda").to(torch.bfloat16)
    kv[..., 128:] = c[None]
    out, _ = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)
    exp = c.repeat_interleave(3, dim=0)[None].float()
    err = (out.float() - exp).abs().max().item()
    assert err < 2e-2, err


def test_fullsize_cache_split_and_key_split_consistency(F):
    """(a) reading the first two chunks from the 'cache' segment == reading everything from one buffer;
    (b) attention over [0, 4 clips) == LSE-merge of attention over two halves of the key range."""
    q, kv, qr, kr = _cfg2_inputs(seed=1)
    k, v = kv[..., :128], kv[..., 128:]
    full, lse = F.attn_ranges(q, k, v, qr, kr)
    cut = 2 * CLIP - 77    # not tile aligned
    split, lse2 = F.attn_ranges(q, k[cut:], v[cut:], qr, kr, k_cache=k[:cut], v_cache=v[:cut])
    assert (full.float() - split.float()).abs().max().item() < 1e-2
    assert (lse - lse2).abs().max().item() < 1e-3
    q3 = q[3 * CLIP:]
    whole, lw = F.attn_ranges(q3, k, v, i32([[0, CLIP]]), i32([[0, 4 * CLIP]]))
    a, la = F.attn_ranges(q3, k, v, i32([[0, CLIP]]), i32([[0, 15000]]))
    b, lb = F.attn_ranges(q3, k, v, i32([[0, CLIP]]), i32([[15000, 4 * CLIP]]))
    F.lse_merge(a, la, b, lb)
    assert (whole.float() - a.float()).abs().max().item() < 1.5e-2
    assert (lw - la).abs().max().item() < 1e-3
    assert torch.equal(full[3 * CLIP:], whole)      # same rectangle -> bit identical regardless of the other ranges


def test_fullsize_matches_flash_attn_library(F):
    """B-REF-GPU: the reference's own fallback (dit_module.py:1046-1065) = a Python loop of flash_attn_func over the
    ranges, run with the flash-attn wheel of this image on identical inputs (both bf16 in/out, fp32 accumulate).
    Bound: max_abs <= 2e-2 (two bf16 roundings of O(1) values) and rel-L2 <= 1e-2."""
    fa = pytest.importorskip("flash_attn")
    q, kv, qr, kr = _cfg2_inputs(seed=2)
    k, v = kv[..., :128].contiguous(), kv[..., 128:].contiguous()
    out, _ = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)
    try:
        refs = [fa.flash_attn_func(q[None, i * CLIP:(i + 1) * CLIP], k[None, :(i + 1) * CLIP], v[None, :(i + 1) * CLIP])
                for i in range(4)]
    except RuntimeError as e:   # pragma: no cover - wheel without sm_100 kernels
        pytest.skip(f"flash-attn cannot run here: {e}")
    ref = torch.cat([r[0] for r in refs], dim=0)
    err = (out.float() - ref.float()).abs().max().item()
    rel = ((out.float() - ref.float()).norm() / ref.float().norm()).item()
    assert err <= ATTN_MAX_ABS and rel <= ATTN_REL_L2, (err, rel)


def test_two_block_dit_matches_reference_golden():
    """BASELINE.json configs[0] in miniature (the DiT cut to 2 blocks): two drop-in TransformerLayers, loaded from the
    reference's state_dicts, chained through the three-forward KV-cache sequence of tests/golden/dit_two_blocks.npz
    (extra clean-chunk forward filling both caches; window advanced, reading them; update_kv_cache False with an empty
    caption segment) against the reference's own two TransformerLayers run on CPU. Tolerance: the per-step latent
    criterion of SURVEY.md section 8(c), rel-L2 <= 2e-2, on the hidden state after both blocks and on the contents of
    both KV caches at the end (the CPU oracle's composition sits at 4.3e-3 on the same fixture)."""
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    from magi_b200.transformer_layer import TransformerLayer, high_precision_promote
    g = Golden("dit_two_blocks.npz")
    hidden = g.raw("w0.self_attention.linear_qkv.layer_norm.weight").shape[0]
    Hq, Hkv = [int(x) for x in g["heads"]]
    mc = ModelConfig(num_layers=2, hidden_size=hidden, ffn_hidden_size=4 * hidden, num_attention_heads=Hq,
                     num_query_groups=Hkv)
    layers = []
    for l in (0, 1):
        layer = high_precision_promote(TransformerLayer(mc, EngineConfig(), layer_number=l))
        layer.load_state_dict({k[3:]: g.torch(k) for k in g.files if k.startswith(f"w{l}.")}, strict=True)
        layers.append(layer.cuda().eval())
    clip = int(g["clip"])
    inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)   # both layer caches are allocated lazily
    for tag in ("A", "B", "C"):
        sp, dn, upd, extra = [int(x) for x in g[f"{tag}.meta"]]
        q_range, k_range = O.core_q_ranges(clip, dn, 1), g.raw(f"{tag}.k_range")
        p = O.cross_attn_params(clip, dn, 1, g[f"{tag}.y_index"])
        core = PackedCoreAttnParams(q_range=i32(q_range), k_range=i32(k_range), np_q_range=q_range, np_k_range=k_range,
                                    max_seqlen_q=clip, max_seqlen_k=O.max_seqlen_k(k_range))
        cross = PackedCrossAttnParams(q_ranges=i32(p["q_ranges"]), kv_ranges=i32(p["kv_ranges"]),
                                      cu_seqlens_q=i32(p["cu_seqlens_q"]), cu_seqlens_kv=i32(p["cu_seqlens_kv"]),
                                      max_seqlen_q=clip, max_seqlen_kv=800)
        meta = ModelMetaArgs(H=4, W=4, cp_pad_size=None, cp_split_sizes=None, slice_point=sp, denoising_range_num=dn,
                             range_num=sp + dn, extract_prefix_video_feature=False, fwd_extra_1st_chunk=bool(extra),
                             distill_nearly_clean_chunk=False, clip_token_nums=clip, enable_cuda_graph=False,
                             core_attn_params=core, cross_attn_params=cross)
        inf.update_kv_cache = bool(upd)
        x = g.torch(f"{tag}.hidden_states", "cuda")
        for layer in layers:
            x = layer(x, g.torch(f"{tag}.condition", "cuda"), g.torch(f"{tag}.condition_map", "cuda"),
                      g.torch(f"{tag}.key_value_states", "cuda"), g.torch(f"{tag}.rope", "cuda"), inf, meta)
        ref = g[f"{tag}.out"]
        got = x.float().cpu().numpy()
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert got.shape == ref.shape and rel < 2e-2, (tag, rel)
    for l in (0, 1):
        ref = g[f"cache{l}"]
        got = inf.key_value_memory_dict[l][:3 * clip].float().cpu().numpy()
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert rel < 2e-2, (l, rel)


def test_graphed_layer_stack_replays_bit_identical_to_eager():
    """EngineConfig.enable_cuda_graph (config.py:107) honoured by magi_b200.cuda_graph.GraphedLayerStack: the two-block
    stack of the fixture above run through its three-forward KV-cache sequence TWICE - eagerly, and as one CUDA graph
    per forward signature (the second pass replays every graph on other hidden states and changed range / caption
    tensors) - must give bit-identical hidden states and KV-cache contents."""
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    from magi_b200.cuda_graph import GraphedLayerStack
    from magi_b200.transformer_layer import TransformerLayer, high_precision_promote
    g = Golden("dit_two_blocks.npz")
    hidden = g.raw("w0.self_attention.linear_qkv.layer_norm.weight").shape[0]
    Hq, Hkv = [int(x) for x in g["heads"]]
    mc = ModelConfig(num_layers=2, hidden_size=hidden, ffn_hidden_size=4 * hidden, num_attention_heads=Hq,
                     num_query_groups=Hkv)
    clip = int(g["clip"])

    def run(graphed):
        ec = EngineConfig(enable_cuda_graph=graphed)
        layers = []
        for l in (0, 1):
            layer = high_precision_promote(TransformerLayer(mc, ec, layer_number=l))
            layer.load_state_dict({k[3:]: g.torch(k) for k in g.files if k.startswith(f"w{l}.")}, strict=True)
            layers.append(layer.cuda().eval())
        stack = GraphedLayerStack(layers, ec)
        inf = InferenceParams(max_batch_size=1, max_sequence_length=4 * clip)
        outs = []
        for rep in range(2):
            for tag in ("A", "B", "C"):
                sp, dn, upd, extra = [int(x) for x in g[f"{tag}.meta"]]
                q_range, k_range = O.core_q_ranges(clip, dn, 1), g.raw(f"{tag}.k_range")
                p = O.cross_attn_params(clip, dn, 1, g[f"{tag}.y_index"])
                core = PackedCoreAttnParams(q_range=i32(q_range), k_range=i32(k_range), np_q_range=q_range,
                                            np_k_range=k_range, max_seqlen_q=clip, max_seqlen_k=O.max_seqlen_k(k_range))
                cross = PackedCrossAttnParams(q_ranges=i32(p["q_ranges"]), kv_ranges=i32(p["kv_ranges"]),
                                              cu_seqlens_q=i32(p["cu_seqlens_q"]), cu_seqlens_kv=i32(p["cu_seqlens_kv"]),
                                              max_seqlen_q=clip, max_seqlen_kv=800)
                meta = ModelMetaArgs(H=4, W=4, cp_pad_size=None, cp_split_sizes=None, slice_point=sp,
                                     denoising_range_num=dn, range_num=sp + dn, extract_prefix_video_feature=False,
                                     fwd_extra_1st_chunk=bool(extra), distill_nearly_clean_chunk=False,
                                     clip_token_nums=clip, enable_cuda_graph=graphed, core_attn_params=core,
                                     cross_attn_params=cross)
                inf.update_kv_cache = bool(upd)
                x = g.torch(f"{tag}.hidden_states", "cuda")
                y = g.torch(f"{tag}.key_value_states", "cuda")
                if rep == 1:                      # same signature, other data
                    x = (x.float() * 0.5 + 0.25).to(torch.bfloat16)
                    y = (y.float() * -1.0).to(torch.bfloat16)
                out = stack(x, g.torch(f"{tag}.condition", "cuda"), g.torch(f"{tag}.condition_map", "cuda"), y,
                            g.torch(f"{tag}.rope", "cuda"), inf, meta)
                outs.append(out.clone())
        torch.cuda.synchronize()
        return outs, [inf.key_value_memory_dict[l][:3 * clip].clone() for l in (0, 1)], stack

    eager_outs, eager_caches, _ = run(False)
    graph_outs, graph_caches, stack = run(True)
    assert stack.disabled_reason is None, stack.disabled_reason
    assert stack.captures == 3 and stack.replays == 6        # one graph per signature, replayed in both passes
    for a, b in zip(eager_outs, graph_outs):
        assert torch.equal(a, b)
    for a, b in zip(eager_caches, graph_caches):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_apply_rotary_emb_dropin(F, dtype):
    """`flash_apply_rotary_emb(q, cos_emb, sin_emb)` (dit_module.py:961, 978) as a stand-alone call, on the layout the
    reference passes: the (b, sq, hn, hd) TRANSPOSED view of the (sq, b, hn, hd) LayerNorm output, fp32; also bf16,
    in place, and a rotary_dim == head_dim table. Oracle: apply_rotary_neox (flash-attn's apply_rotary_emb_torch).
    fp32: the two sides may contract the multiply-adds differently -> 4 fp32 ulp at |x| <= 8; bf16: one rounding."""
    sq, b, hn, hd = 37, 2, 3, 128
    x_sb = torch.from_numpy(rnd((sq, b, hn, hd), 5)).to("cuda").to(dtype)
    sin, cos = O.rope_table(3, 4, 4)                                     # (48 rows, 48) each
    sin, cos = sin[:sq], cos[:sq]
    x = x_sb.transpose(0, 1)                                             # (b, sq, hn, hd), non-contiguous
    out = F.apply_rotary_emb(x, torch.from_numpy(cos).cuda(), torch.from_numpy(sin).cuda())
    ref = O.apply_rotary_neox(x.float().cpu().numpy(), cos, sin)
    assert out.shape == x.shape and out.dtype == dtype
    if dtype == torch.float32:
        np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=4e-6)
    else:
        assert torch.equal(out.cpu(), torch.from_numpy(ref).to(torch.bfloat16)) or \
            np.abs(out.float().cpu().numpy() - ref).max() <= 2.0 ** -7 * 8
    assert torch.equal(out[..., 96:], x[..., 96:])                       # dims past rotary_dim pass through
    # in place on a contiguous tensor == out of place
    y = x.contiguous()
    r = F.apply_rotary_emb(y, torch.from_numpy(cos).cuda(), torch.from_numpy(sin).cuda(), inplace=True)
    assert r.data_ptr() == y.data_ptr() and torch.equal(y, out)
    # full-width rotation (rotary_dim == head_dim)
    g = np.random.default_rng(3)
    ang = g.uniform(0, 6.28, (sq, 64)).astype(np.float32)
    out2 = F.apply_rotary_emb(x, torch.from_numpy(np.cos(ang)).cuda(), torch.from_numpy(np.sin(ang)).cuda())
    ref2 = O.apply_rotary_neox(x.float().cpu().numpy(), np.cos(ang), np.sin(ang))
    tol = 4e-6 if dtype == torch.float32 else 2.0 ** -7 * 8
    assert np.abs(out2.float().cpu().numpy() - ref2).max() <= tol


def test_maxsize_24b_720p_sampled_rows_vs_oracle(F):
    """The largest single-GPU state of BASELINE.json (configs[3]'s per-GPU problem before sharding: 24B heads 48 / 8,
    720x1280 -> 21 600 tokens per chunk, 5 ranges in flight = 108 000 query rows, keys up to 3 chunks, noise2clean
    windows that start mid-sequence) in ONE launch; the fp32 oracle is evaluated on sampled (row, head) pairs - first,
    last and a random row of every range, first / last / random head - where it costs a 64 800 x 128 product each.
    Same tolerances as everywhere: out max_abs <= 2e-2, lse <= 1e-3. Also: the launch is deterministic."""
    clip, Hq, Hkv = 21600, 48, 8
    k_chunks = [(0, 1), (0, 2), (0, 3), (1, 4), (3, 5)]
    g = torch.Generator(device="cuda").manual_seed(11)
    q = torch.randn(5 * clip, Hq, 128, device="cuda", dtype=torch.bfloat16, generator=g)
    kv = torch.randn(5 * clip, Hkv, 256, device="cuda", dtype=torch.bfloat16, generator=g)
    qr = i32([[i * clip, (i + 1) * clip] for i in range(5)])
    kr = i32([[s * clip, e * clip] for s, e in k_chunks])
    out, lse = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)
    out2, lse2 = F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)
    assert torch.equal(out, out2) and torch.equal(lse, lse2)
    rng = np.random.default_rng(5)
    worst_o = worst_l = 0.0
    for i, (s, e) in enumerate(k_chunks):
        for row in (i * clip, i * clip + int(rng.integers(1, clip - 1)), (i + 1) * clip - 1):
            for h in (0, int(rng.integers(1, Hq - 1)), Hq - 1):
                hk = h // (Hq // Hkv)
                keys = kv[s * clip:e * clip, hk].float().cpu().numpy()
                ref_o, ref_l = O.attn_ranges(q[row:row + 1, h:h + 1].float().cpu().numpy(), keys[:, None, :128],
                                             keys[:, None, 128:], [[0, 1]], [[0, keys.shape[0]]], round_output=False)
                worst_o = max(worst_o, float(np.abs(out[row, h].float().cpu().numpy() - ref_o[0, 0]).max()))
                worst_l = max(worst_l, abs(float(lse[row, h]) - float(ref_l[0, 0])))
    assert worst_o <= ATTN_MAX_ABS and worst_l <= 1e-3, (worst_o, worst_l)
