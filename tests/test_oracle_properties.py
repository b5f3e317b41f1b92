"""Property tests (hypothesis) of the CPU oracle on random ragged range sets: the invariants the GPU parity tests rely
on at full size (tests/test_gpu_parity.py::test_fullsize_*) hold for the checker itself, and its two forms - the numpy
oracle and the torch fp32 CPU path that bench.py times - agree. CPU only."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import magi_oracle as O
import torch_cpu_path as T


@st.composite
def range_sets(draw):
    """Disjoint ascending q ranges (some empty, some rows uncovered) and arbitrary k ranges (some empty)."""
    n = draw(st.integers(1, 4))
    hkv = draw(st.sampled_from([1, 2]))
    hq = hkv * draw(st.sampled_from([1, 3]))
    tk = draw(st.integers(1, 90))
    q_edges = sorted(draw(st.lists(st.integers(0, 70), min_size=2 * n, max_size=2 * n)))
    q_ranges = [[q_edges[2 * i], q_edges[2 * i + 1]] for i in range(n)]
    k_ranges = []
    for _ in range(n):
        a = draw(st.integers(0, tk))
        b = draw(st.integers(a, tk))
        k_ranges.append([a, b])
    seed = draw(st.integers(0, 2 ** 16))
    return hq, hkv, max(q_edges[-1], 1), tk, np.array(q_ranges), np.array(k_ranges), seed


def _inputs(hq, hkv, tq, tk, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((tq, hq, 128)).astype(np.float32), rng.standard_normal((tk, hkv, 128)).astype(np.float32),
            rng.standard_normal((tk, hkv, 128)).astype(np.float32))


@settings(max_examples=40, deadline=None)
@given(range_sets())
def test_numpy_oracle_and_torch_cpu_path_agree(case):
    hq, hkv, tq, tk, qr, kr, seed = case
    q, k, v = _inputs(hq, hkv, tq, tk, seed)
    out, lse = O.attn_ranges(q, k, v, qr, kr, round_output=False)
    got = T.core_attention(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(v), qr, kr).numpy()
    rows = np.concatenate([np.arange(a, b) for a, b in qr]) if len(qr) else np.zeros(0, int)
    assert got.shape[0] == len(rows)
    np.testing.assert_allclose(got, out[rows], rtol=0, atol=3e-5)
    # rows of an empty key interval: zeros and lse = -inf; rows covered by no range: untouched zeros
    for (a, b), (c, d) in zip(qr, kr):
        if d <= c and b > a:
            assert not out[a:b].any() and np.isneginf(lse[a:b]).all()
    covered = np.zeros(tq, bool)
    covered[rows] = True
    assert not out[~covered].any() and np.isneginf(lse[~covered]).all()


@settings(max_examples=25, deadline=None)
@given(range_sets(), st.integers(0, 90))
def test_key_split_plus_lse_merge_is_the_whole_range(case, cut):
    """The identity behind the ring-KV exchange and the two-segment (cache ++ new) kernel: attending [c, m) and [m, d)
    separately and merging by log-sum-exp equals attending [c, d)."""
    hq, hkv, tq, tk, qr, kr, seed = case
    q, k, v = _inputs(hq, hkv, tq, tk, seed)
    whole, lse = O.attn_ranges(q, k, v, qr, kr, round_output=False)
    mid = np.clip(cut, kr[:, 0], kr[:, 1])
    a, la = O.attn_ranges(q, k, v, qr, np.stack([kr[:, 0], mid], 1), round_output=False)
    b, lb = O.attn_ranges(q, k, v, qr, np.stack([mid, kr[:, 1]], 1), round_output=False)
    m, lm = O.lse_merge(a, la, b, lb)
    ok = np.isfinite(lse)
    np.testing.assert_allclose(lm[ok], lse[ok], rtol=0, atol=2e-5)
    assert np.array_equal(np.isfinite(lm), ok)
    np.testing.assert_allclose(m, O.bf16_round(whole), rtol=0, atol=2.0 ** -7 * max(1.0, float(np.abs(whole).max())))


@settings(max_examples=25, deadline=None)
@given(range_sets())
def test_constant_v_gives_constant_rows_and_gqa_maps_by_group(case):
    hq, hkv, tq, tk, qr, kr, seed = case
    q, k, v = _inputs(hq, hkv, tq, tk, seed)
    const = np.broadcast_to(np.arange(hkv, dtype=np.float32)[None, :, None] + 1.0, v.shape)
    out, lse = O.attn_ranges(q, k, const, qr, kr, round_output=False)
    g = hq // hkv
    for (a, b), (c, d) in zip(qr, kr):
        if b > a and d > c:
            for h in range(hq):          # softmax weights sum to 1: every covered row equals its kv head's constant
                np.testing.assert_allclose(out[a:b, h], h // g + 1.0, rtol=0, atol=2e-5)


def test_mlp_activation_follows_the_published_definitions():
    """silu_and_mul (flashinfer, dit_module.py:546): act(float(a)) * float(b), one rounding; gelu: torch's erf form on
    a bf16 tensor. Checked against torch on the CPU (the same formulas through another implementation)."""
    import torch
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(37, 2 * 96, generator=g) * 3).to(torch.bfloat16)
    a, b = x[:, :96].float(), x[:, 96:].float()
    want = (torch.nn.functional.silu(a) * b).to(torch.bfloat16).float().numpy()
    got = O.mlp_activation(x.float().numpy(), True)
    assert np.abs(got - want).max() <= np.abs(want).max() * 2.0 ** -7   # <= 1 bf16 ulp (exp implementations differ)
    assert (got != want).mean() < 0.01
    want = torch.nn.functional.gelu(x).float().numpy()
    got = O.mlp_activation(x.float().numpy(), False)
    # 1 + erf(x / sqrt 2) cancels for x << 0: implementations differ in the last bf16 bit on a few per cent of elements
    assert (got != want).mean() < 0.05 and np.all(np.abs(got - want) <= np.abs(want) * 2.0 ** -6 + 1e-5)
