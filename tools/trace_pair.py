"""Reads the clock64 trace of a -DMAGI_TRACE build of the CTA-pair attention kernel (one long item) and prints the
average per-tile phase durations of softmax warps 0 and 4 (same SM sub-partition) and of the MMA issuer, leader CTA of
pair 0. GPU box only.  usage: MAGI_ATTN_LIB=magi-1_b200/variants/lib_ptrace.so python tools/trace_pair.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MAGI_ATTN_KERNEL"] = "pair"
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
buf = np.zeros((4, 512, 8), dtype=np.int64)
L = _lib.lib()
rc = L.magi_debug_read_trace_pair(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.nbytes))
assert rc == 0, rc
lo, hi = 20, 280   # steady-state tiles of the first item (293 tiles)
for t in (0, 1):
    b = buf[t, lo:hi].astype(np.float64)
    names = ["wait S", "ld S + arrive s_free", "exps", "max + vote", "st P", "fence + arrive p_full"]
    d = [b[:, i + 1] - b[:, i] for i in range(6)]
    print(f"softmax warp {4 * t}: period {np.diff(b[:, 0]).mean():.0f} cyc | "
          + " | ".join(f"{n} {x.mean():.0f}" for n, x in zip(names, d))
          + f" | loop tail {(b[1:, 0] - b[:-1, 6]).mean():.0f}")
ms, mp = buf[2, lo:hi].astype(np.float64), buf[3, lo:hi].astype(np.float64)
print(f"issuer S : period {np.diff(ms[:, 0]).mean():.0f} | wait K {(ms[:, 1] - ms[:, 0]).mean():.0f} | wait s_free "
      f"{(ms[:, 2] - ms[:, 1]).mean():.0f} | issue+commit {(ms[:, 3] - ms[:, 2]).mean():.0f}")
print(f"issuer PV: period {np.diff(mp[:, 0]).mean():.0f} | wait V {(mp[:, 1] - mp[:, 0]).mean():.0f} | wait p_full "
      f"{(mp[:, 2] - mp[:, 1]).mean():.0f} | issue+commit {(mp[:, 3] - mp[:, 2]).mean():.0f}")
s0 = buf[0, lo:hi].astype(np.float64)
s1 = buf[1, lo:hi].astype(np.float64)
print(f"S(g) issued -> warp 0 woke: {(s0[:, 1] - ms[:, 3]).mean():.0f} cyc (includes the 512-cycle MMA) ; "
      f"warp 0 arrived p_full -> issuer past wait P(g): {(mp[:, 2] - s0[:, 6]).mean():.0f} ; "
      f"warp 4 arrived -> {(mp[:, 2] - s1[:, 6]).mean():.0f}")
print(f"PV(g) issued -> S(g+2) issued: {(ms[2:, 3] - mp[:-2, 3]).mean():.0f} ; S(g+1) issued -> PV(g) issued: "
      f"{(mp[:-1, 3] - ms[1:, 3]).mean():.0f}")
rel = lambda a: (a - ms[:, 3:4])
print("timeline of one tile relative to its S issue (mean cycles):")
print("  warp0 stamps", np.round(rel(s0[:, :7]).mean(0)).astype(int).tolist())
print("  warp4 stamps", np.round(rel(s1[:, :7]).mean(0)).astype(int).tolist())
print("  PV stamps   ", np.round(rel(mp[:, :4]).mean(0)).astype(int).tolist())
