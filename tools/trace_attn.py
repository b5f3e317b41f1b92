"""Reads the clock64 trace of a -DMAGI_TRACE build (tools/build_variants.sh trace=...) for one long item and prints
the average per-tile phase durations of the softmax warpgroups and the MMA issuer of CTA 0. GPU box only.
usage: MAGI_ATTN_LIB=magi-1_b200/variants/lib_trace.so python tools/trace_attn.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import magi_b200.functional as F  # noqa: E402
from magi_b200 import _lib  # noqa: E402

clip, Hq, Hkv = 9360, 24, 8
q = torch.randn(clip, Hq, 128, device="cuda", dtype=torch.bfloat16)
kv = torch.randn(4 * clip, Hkv, 256, device="cuda", dtype=torch.bfloat16)
qr = torch.tensor([[0, clip]], dtype=torch.int32, device="cuda")
kr = torch.tensor([[0, 4 * clip]], dtype=torch.int32, device="cuda")
for _ in range(2):
    F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr)
torch.cuda.synchronize()
buf = np.zeros((3, 512, 8), dtype=np.int64)
L = _lib.lib()
rc = L.magi_debug_read_trace(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.nbytes))
assert rc == 0, rc
lo, hi = 20, 280   # steady state tiles of item 0 (293 tiles)
for t in (0, 1):
    b = buf[t, lo:hi].astype(np.float64)
    names = ["wait S", "tcgen05.ld S", "(max, j=0 only)", "half0 exp+st+arrive", "half1 exp+st+wait", "arrive"]
    d = [b[:, i + 1] - b[:, i] for i in range(6)]
    period = np.diff(b[:, 0]).mean()
    print(f"softmax WG{t}: period {period:.0f} cyc | " + " | ".join(f"{n} {x.mean():.0f}" for n, x in zip(names, d)))
m = buf[2, lo:hi].astype(np.float64)
names = ["issue S0+commit", "wait o_free", "PV1 (waits+issue), S1, commits", "wait V", "PV0 (waits+issue)", "-"]
d = [m[:, i + 1] - m[:, i] for i in range(6)]
period = np.diff(m[:, 0]).mean()
loop = m[1:, 0] - m[:-1, 6]
print(f"MMA issuer : period {period:.0f} cyc | " + " | ".join(f"{n} {x.mean():.0f}" for n, x in zip(names, d))
      + f" | wait K (loop top) {loop.mean():.0f}")
# latency from S0 issue (MMA stamp 1) to softmax WG0 seeing it (softmax stamp 1) and from arrive to MMA wake
s0 = buf[0, lo:hi].astype(np.float64)
print(f"S0 commit issued -> softmax0 woke: {(s0[:, 1] - m[:, 1]).mean():.0f} cyc ; softmax0 arrive -> MMA past wait P0: "
      f"{(m[:, 5] - s0[:, 6]).mean():.0f} cyc")
