"""Host-side range / metadata construction (magi_b200.meta) against the reference's golden tables and the oracle.
CPU only: the builders are integer arithmetic plus one non-blocking copy per table."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import magi_b200  # noqa: E402,F401
import magi_oracle as O  # noqa: E402
from golden_io import Golden  # noqa: E402
from magi_b200 import meta  # noqa: E402


def test_kv_ranges_match_the_reference_tables():
    """tests/golden/ranges.npz holds the outputs of the reference's own generate_default_kvrange /
    generate_noise2clean_kvrange (video_generate.py:454-508) for states of the 64-step, window-4 schedule."""
    r = Golden("ranges.npz")
    i = 0
    while f"kv{i}.args" in r:
        a = [int(x) for x in r[f"kv{i}.args"]]
        sp, dn, bs, steps = a[0], a[1], a[2], a[3:]
        got = meta.noise2clean_kv_range(sp, dn, [5, 4, 3, 2], 1, steps, 64, bs, 9360)
        assert got.dtype == np.int32 and np.array_equal(got, r[f"kv{i}.noise2clean"]), a
        assert np.array_equal(meta.default_kv_range(sp, dn, bs, 9360), r[f"kv{i}.default"]), a
        assert np.array_equal(meta.kv_range_for_denoising_video(sp, dn, steps, [], 1, 64, bs, 9360),
                              r[f"kv{i}.default"])
        assert np.array_equal(meta.kv_range_for_denoising_video(sp, dn, steps, [5, 4, 3, 2], 1, 64, bs, 9360),
                              r[f"kv{i}.noise2clean"])
        i += 1
    assert i >= 4


@pytest.mark.parametrize("seed", range(8))
def test_kv_ranges_match_the_oracle_on_random_states(seed):
    rng = np.random.default_rng(seed)
    sp, dn, bs = int(rng.integers(0, 6)), int(rng.integers(1, 6)), int(rng.integers(1, 4))
    stages = [int(x) for x in sorted(rng.integers(1, 7, size=int(rng.choice([1, 2, 4]))), reverse=True)]
    num_steps = len(stages) * int(rng.integers(1, 9))
    steps = [int(x) for x in rng.integers(0, num_steps + 1, size=dn)]
    clean = int(rng.choice([-1, 1, 2]))
    chunk = int(rng.integers(1, 3000))
    assert np.array_equal(meta.noise2clean_kv_range(sp, dn, stages, clean, steps, num_steps, bs, chunk),
                          O.generate_noise2clean_kvrange(sp, dn, stages, clean, steps, num_steps, bs, chunk))
    assert np.array_equal(meta.default_kv_range(sp, dn, bs, chunk), O.generate_default_kvrange(sp, dn, bs, chunk))


def test_uncondition_and_nearly_clean_ranges():
    """dit_model.py:92-101 (linspace of chunk edges) and :544-552 (self-only range after the last key)."""
    assert meta.kv_range_for_uncondition(3, 1536).tolist() == [[0, 1536], [1536, 3072], [3072, 4608]]
    kv = meta.default_kv_range(1, 3, 1, 100)                   # [0,200) [0,300) [0,400)
    cat = meta.append_nearly_clean_range(kv, 100)
    assert cat.tolist() == [[0, 200], [0, 300], [0, 400], [400, 500]] and cat.dtype == np.int32


def test_errors():
    with pytest.raises(ValueError):
        meta.noise2clean_kv_range(0, 2, [5, 4, 3, 2], 1, [0], 64, 1, 10)        # one step per chunk
    with pytest.raises(ValueError):
        meta.noise2clean_kv_range(0, 1, [5, 4, 3], 1, [0], 64, 1, 10)           # 64 % 3 != 0
    with pytest.raises(ValueError):
        meta.core_attn_params(np.zeros((3, 2), np.int32), 10, 2, 1, "cpu")


@pytest.mark.parametrize("bs,dn", [(1, 4), (2, 1), (3, 1)])
def test_packed_params_match_the_oracle(bs, dn):
    clip, L = 96, 40
    rng = np.random.default_rng(bs * 10 + dn)
    y_index = rng.integers(0, L + 1, size=bs * dn)
    y_index[0] = 0                                             # a slot without caption tokens
    mask = torch.from_numpy((np.arange(L)[None, :] < y_index[:, None]).astype(np.int64)).reshape(bs, dn, 1, L)
    kv = meta.default_kv_range(2, dn, bs, clip)
    core = meta.core_attn_params(kv, clip, dn, bs, "cpu")
    assert np.array_equal(core.np_q_range, O.core_q_ranges(clip, dn, bs))
    assert torch.equal(core.q_range, torch.from_numpy(O.core_q_ranges(clip, dn, bs)))
    assert torch.equal(core.k_range, torch.from_numpy(kv)) and core.k_range.dtype == torch.int32
    assert core.max_seqlen_k == O.max_seqlen_k(kv) and core.max_seqlen_q == clip
    cross = meta.cross_attn_params(mask, clip, dn, bs, 800)
    ref = O.cross_attn_params(clip, dn, bs, y_index)
    for name, got in (("q_ranges", cross.q_ranges), ("kv_ranges", cross.kv_ranges),
                      ("cu_seqlens_q", cross.cu_seqlens_q), ("cu_seqlens_kv", cross.cu_seqlens_kv)):
        assert got.dtype == torch.int32 and np.array_equal(got.numpy(), ref[name]), name
    assert cross.max_seqlen_q == clip and cross.max_seqlen_kv == 800
    # condition_map, dit_model.py:203-205 restated literally
    lit = torch.arange(bs * dn).repeat_interleave(clip).reshape(bs, -1).transpose(0, 1).contiguous()
    assert torch.equal(meta.condition_map(bs, dn, clip, "cpu"), lit)
    # caption packing, dit_model.py:208
    y = torch.randn(bs * dn, 1, L, 8)
    lit_y = torch.masked_select(y.squeeze(1), mask.reshape(bs * dn, L).unsqueeze(-1).bool()).reshape(-1, 8)
    assert torch.equal(meta.pack_caption_tokens(y, mask), lit_y)


def test_build_meta_args_single_rank():
    clip, dn, sp, L, C = 64, 3, 1, 16, 8
    x = torch.randn(dn * clip, 1, C)
    rope = torch.randn(dn * clip, 96)
    mask = torch.ones(dn, L, dtype=torch.int64)
    kv = meta.append_nearly_clean_range(meta.default_kv_range(sp, dn - 1, 1, clip), clip)
    x2, cmap, rope2, m = meta.build_meta_args(kv, mask, x, rope, batch_size=1, denoising_range_num=dn, slice_point=sp,
                                              range_num=sp + dn, H=4, W=4, caption_max_length=800,
                                              fwd_extra_1st_chunk=True, distill_nearly_clean_chunk=True)
    assert x2 is x and rope2 is rope and cmap.shape == (dn * clip, 1)
    assert m.clip_token_nums == clip and m.slice_point == sp and m.denoising_range_num == dn and m.range_num == sp + dn
    assert m.cp_split_sizes is None and m.distill_nearly_clean_chunk and m.fwd_extra_1st_chunk
    assert m.core_attn_params.np_k_range.tolist() == [[0, 2 * clip], [0, 3 * clip], [3 * clip, 4 * clip]]
    assert m.core_attn_params.max_seqlen_k == 4 * clip
    assert m.cross_attn_params.cu_seqlens_kv.tolist() == [0, L, 2 * L, 3 * L]
