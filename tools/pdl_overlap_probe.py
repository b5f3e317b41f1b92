"""Does the cross-attention launch really overlap the tail of the self-attention launch before it
(magi_attn_args.launch_dependency = 1)? One GPU, the per-rank shapes of cfg2 under cp_ulysses with 8 ranks: the
self-attention has 4 ranges x 37 query tiles x 3 heads = 444 items for 148 CTAs (half of the CTAs run out of work one
chunk-unit early), the cross-attention 4680 local rows x 24 heads. Times the pair back to back, A/B/A/B. GPU box only.
usage: python tools/pdl_overlap_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import magi_b200.functional as F  # noqa: E402

dev = "cuda"
clip, cp, Hq, Hkv = 9360, 8, 24, 8
hq, hkv = Hq // cp, 1
S = 4 * clip
q = torch.randn(S, hq, 128, device=dev, dtype=torch.bfloat16)
kv = torch.randn(S, hkv, 256, device=dev, dtype=torch.bfloat16)
qr = torch.tensor([[i * clip, (i + 1) * clip] for i in range(4)], dtype=torch.int32, device=dev)
kr = torch.tensor([[0, (i + 1) * clip] for i in range(4)], dtype=torch.int32, device=dev)
out = torch.empty_like(q)
rows = S // cp
xq = torch.randn(rows, Hq, 128, device=dev, dtype=torch.bfloat16)
xkv = torch.randn(512, Hkv, 256, device=dev, dtype=torch.bfloat16)
xqr = torch.tensor([[0, rows]], dtype=torch.int32, device=dev)
xkr = torch.tensor([[0, 512]], dtype=torch.int32, device=dev)
xout = torch.empty_like(xq)
plan = F.attn_plan(qr, kr, num_heads_q=hq, total_q=S, total_k=S)
xplan = F.attn_plan(xqr, xkr, num_heads_q=Hq, total_q=rows, total_k=512)


def pair(independent):
    F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False, plan=plan)
    F.attn_ranges(xq, xkv[..., :128], xkv[..., 128:], xqr, xkr, out=xout, return_lse=False, plan=xplan,
                  independent_of_prev=independent)


def timed(fn, n=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


only_self = lambda: F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False, plan=plan)
only_cross = lambda: F.attn_ranges(xq, xkv[..., :128], xkv[..., 128:], xqr, xkr, out=xout, return_lse=False, plan=xplan)
ref = xout.clone()
for rnd in range(3):
    a = timed(only_self)
    b = timed(only_cross)
    c = timed(lambda: pair(False))
    d = timed(lambda: pair(True))
    print(f"round {rnd}: self {a:7.1f} us | cross {b:6.1f} us | pair serial {c:7.1f} us | pair overlapped {d:7.1f} us "
          f"(hidden {c - d:5.1f} of {b:5.1f} us)", flush=True)
only_cross()
torch.cuda.synchronize()
ref = xout.clone()
pair(True)
torch.cuda.synchronize()
print("cross output identical with and without the overlap:", bool(torch.equal(ref, xout)))
