"""Test infrastructure: runs the reference's OWN FullyParallelAttention (imported from /root/reference, CUDA-only
third-party kernels replaced by their published math exactly as in gen_golden.py) on RANDOM configurations and state
sequences and compares the numpy oracle's composition with it. Executed in a subprocess by
tests/test_live_reference.py (the stand-ins patch torch.cuda globally), only where /root/reference exists.

    python oracle/live_attention_check.py [n_cases]   ->  one JSON line {"cases": n, "worst_rel": r, ...}
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))

import gen_golden as G  # noqa: E402
import magi_oracle as O  # noqa: E402
from golden_io import f32  # noqa: E402


def val(t):
    return f32(G.pk(t))


def main(n_cases):
    dm = G.install_patches()
    import socket
    with socket.socket() as s_:           # a free rendezvous port for the single-rank gloo group the reference needs
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    G.init_dist(0, 1, port)
    from inference.common import InferenceParams
    rng = np.random.default_rng(2024)
    worst = dict(core_rel=0.0, xattn_rel=0.0, cache_rel=0.0)
    for case in range(n_cases):
        hq, hkv = [(4, 2), (6, 2), (4, 4), (8, 2), (6, 3)][int(rng.integers(0, 5))]
        mc, ec = G.make_configs(hidden=128, hq=hq, hkv=hkv)
        attn = G.build_attention(dm, mc, ec, seed=1000 + case)
        w = {k[2:]: f32(v) for k, v in G.state_to_np(attn).items()}
        H = W = 4
        frames = int(rng.integers(1, 4))
        clip = frames * H * W
        g = torch.Generator().manual_seed(5000 + case)
        inf = InferenceParams(max_batch_size=1, max_sequence_length=8 * clip)
        inf.key_value_memory_dict[1] = torch.zeros(8 * clip, hkv, 256, dtype=torch.bfloat16)
        cache = np.zeros((8 * clip, hkv, 256), dtype=np.float32)
        # a random walk through the schedule: the window only advances after an extra clean-chunk forward cached it
        sp = 0
        for step in range(4):
            dn = int(rng.integers(1, 4))
            extra = bool(rng.integers(0, 2)) or sp > 0
            update = bool(rng.integers(0, 4) > 0)
            distill = bool(rng.integers(0, 3) == 0) and dn >= 2
            width = rng.integers(1, 4, size=dn)
            k_chunks = [(max(0, sp + j + 1 - int(width[j])), sp + j + 1) for j in range(dn)]
            if distill:      # dit_model.py:544-552: the last range is the self-only duplicate behind the last key
                top = max(e for _, e in k_chunks[:-1])
                k_chunks[-1] = (top, top + 1)
            y_index = [int(v) for v in rng.integers(0, 40, size=dn)]
            if sum(y_index) == 0:
                y_index[-1] = 7
            S = dn * clip
            hs = torch.randn(S, 1, mc.hidden_size, generator=g).to(torch.bfloat16)
            kvs = torch.randn(sum(y_index), mc.hidden_size, generator=g).to(torch.bfloat16)
            rope = G.rope_for(dm, (sp + dn) * frames, H, W, dn * frames)
            k_range = np.array([[s * clip, e * clip] for s, e in k_chunks], dtype=np.int32)
            m = G.meta(dm, clip, sp, dn, 1, k_range.tolist(), y_index, extra, distill)
            inf.update_kv_cache = update
            with torch.no_grad():
                core, xattn = attn(hs, kvs, inf, rope, m)
            o_core, o_x = O.fully_parallel_attention(w, val(hs), val(kvs), val(rope), cache, True, sp, dn, clip,
                                                     O.core_q_ranges(clip, dn, 1), k_range, y_index, update,
                                                     fwd_extra_1st_chunk=extra, distill_nearly_clean_chunk=distill)
            for name, a, b in (("core_rel", o_core, val(core)), ("xattn_rel", o_x, val(xattn))):
                rel = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-9))
                worst[name] = max(worst[name], rel)
            used = (sp + dn - (1 if distill else 0)) * clip if (extra or sp > 0) and update else sp * clip
            if used > 0:
                ref_cache = val(inf.key_value_memory_dict[1])[:used]
                rel = float(np.linalg.norm(cache[:used] - ref_cache) / max(np.linalg.norm(ref_cache), 1e-9))
                worst["cache_rel"] = max(worst["cache_rel"], rel)
            if extra and update and rng.integers(0, 2) and sp + dn < 5:
                sp += 1     # the first in-flight chunk is finished and cached: the window moves on
    print(json.dumps(dict(cases=n_cases, **worst)))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
