"""Reads the per-ITEM clock64 stamps of a -DMAGI_TRACE build (tools/build_variants.sh trace="-DMAGI_TRACE") for the
short-item regime - the text cross-attention of cfg2 (4 K/V tiles per item) or cfg1's self-attention - and prints
where the time of an item goes in each role of CTA 0 (softmax warp 0, MMA issuer, TMA warp). GPU box only.
usage: MAGI_ATTN_LIB=magi-1_b200/variants/lib_trace.so python tools/trace_items.py [cross|cfg1]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import magi_b200.functional as F  # noqa: E402
from magi_b200 import _lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "cross"
dev = "cuda"
if what == "cross":
    clip, Hq, Hkv, text = 9360, 24, 8, [512, 512, 512, 512]
    q = torch.randn(4 * clip, Hq, 128, device=dev, dtype=torch.bfloat16)
    kv = torch.randn(sum(text), Hkv, 256, device=dev, dtype=torch.bfloat16)
    qr = torch.tensor([[i * clip, (i + 1) * clip] for i in range(4)], dtype=torch.int32, device=dev)
    cu = np.concatenate([[0], np.cumsum(text)])
    kr = torch.tensor([[cu[i], cu[i + 1]] for i in range(4)], dtype=torch.int32, device=dev)
else:
    clip, Hq, Hkv = 1536, 24, 8
    q = torch.randn(clip, Hq, 128, device=dev, dtype=torch.bfloat16)
    kv = torch.randn(clip, Hkv, 256, device=dev, dtype=torch.bfloat16)
    qr = torch.tensor([[0, clip]], dtype=torch.int32, device=dev)
    kr = torch.tensor([[0, clip]], dtype=torch.int32, device=dev)
out = torch.empty_like(q)
plan = F.attn_plan(qr, kr, num_heads_q=Hq, total_q=q.shape[0], total_k=kv.shape[0])
for _ in range(3):
    F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False, plan=plan)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False, plan=plan)
e1.record()
torch.cuda.synchronize()
buf = np.zeros((3, 128, 8), dtype=np.int64)
rc = _lib.lib().magi_debug_read_itrace(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.nbytes))
assert rc == 0, rc
s, m, t = (buf[i].astype(np.float64) for i in range(3))
n = int((s[:, 0] > 0).sum())
lo, hi = 2, max(3, min(n - 2, 20))
print(f"[{what}] launch {e0.elapsed_time(e1) * 1e3:.1f} us, CTA 0 ran {n} items; averages over items {lo}..{hi - 1} (cycles)")
per = np.diff(s[lo:hi + 1, 0]).mean()
print(f"softmax warp 0: item period {per:.0f} | item known -> first S ready {np.mean(s[lo:hi, 1] - s[lo:hi, 0]):.0f} | "
      f"tiles {np.mean(s[lo:hi, 2] - s[lo:hi, 1]):.0f} | wait O complete {np.mean(s[lo:hi, 3] - s[lo:hi, 2]):.0f} | "
      f"epilogue {np.mean(s[lo:hi, 4] - s[lo:hi, 3]):.0f} | -> next item known {np.mean(s[lo + 1:hi + 1, 0] - s[lo:hi, 4]):.0f}")
print(f"MMA issuer    : item known -> Q landed {np.mean(m[lo:hi, 1] - m[lo:hi, 0]):.0f} | -> K0 landed "
      f"{np.mean(m[lo:hi, 2] - m[lo:hi, 1]):.0f} | issue S0 S1 {np.mean(m[lo:hi, 3] - m[lo:hi, 2]):.0f} | rest of the item "
      f"{np.mean(m[lo:hi, 4] - m[lo:hi, 3]):.0f} | -> next item known {np.mean(m[lo + 1:hi + 1, 0] - m[lo:hi, 4]):.0f}")
print(f"TMA warp      : drawn+decoded -> Q buffer free {np.mean(t[lo:hi, 1] - t[lo:hi, 0]):.0f} | issue Q "
      f"{np.mean(t[lo:hi, 2] - t[lo:hi, 1]):.0f} | K/V loads of the item {np.mean(t[lo:hi, 3] - t[lo:hi, 2]):.0f} | "
      f"draw + decode of the next {np.mean(t[lo + 1:hi + 1, 0] - t[lo:hi, 3]):.0f}")
print(f"cross-role    : Q issued -> MMA sees Q {np.mean(m[lo:hi, 1] - t[lo:hi, 2]):.0f} | MMA first S issued -> softmax sees S "
      f"{np.mean(s[lo:hi, 1] - m[lo:hi, 3]):.0f} | softmax epilogue done -> MMA item end {np.mean(m[lo:hi, 4] - s[lo:hi, 4]):.0f} | "
      f"TMA draws next item {np.mean(t[lo + 1:hi + 1, 0] - s[lo:hi, 0]):.0f} after softmax started this one")
