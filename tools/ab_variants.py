"""A/B/A/B timing of library variants (magi-1_b200/variants/lib_*.so) inside ONE process, interleaved, in the
power-capped regime: the cfg2 self-attention launch back to back, `--rounds` passes over all variants in turn, so
that box-to-box and warm-up / thermal drift (which is as large as the differences being measured when every variant
runs in its own process one after the other, tools/bench_variants.py) cancels. GPU box only.
usage: python tools/ab_variants.py [--rounds 4] [--launches 40] [name ...]"""
import argparse
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import magi_b200.functional as F  # noqa: E402
from magi_b200 import _lib  # noqa: E402


def use(path):
    # lib_force0* / lib_force1*: copies of a build that are loaded with MAGI_SCHED_DECODED forced (read once per library)
    name = os.path.basename(path)
    os.environ.pop("MAGI_SCHED_DECODED", None)
    if name.startswith("lib_force"):
        os.environ["MAGI_SCHED_DECODED"] = name[9]
    _lib.LIB_PATH = path
    _lib._lib = None
    _lib.lib()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--launches", type=int, default=40)
    ap.add_argument("names", nargs="*")
    args = ap.parse_args()
    libs = {os.path.basename(p)[4:-3]: p for p in sorted(glob.glob(os.path.join(ROOT, "magi-1_b200", "variants", "lib_*.so")))}
    if args.names:
        libs = {k: v for k, v in libs.items() if k in args.names}
    dev = "cuda"
    clip, Hq, Hkv = 9360, 24, 8
    q = torch.randn(4 * clip, Hq, 128, device=dev, dtype=torch.bfloat16)
    kv = torch.randn(4 * clip, Hkv, 256, device=dev, dtype=torch.bfloat16)
    qr = torch.tensor([[i * clip, (i + 1) * clip] for i in range(4)], dtype=torch.int32, device=dev)
    kr = torch.tensor([[0, (i + 1) * clip] for i in range(4)], dtype=torch.int32, device=dev)
    out = torch.empty_like(q)
    flops = 4 * 128 * Hq * clip * clip * 10
    # cross-attention of the same state (text lengths as bench.py's cfg2)
    text = [50, 512, 512, 512]   # bench.py WORKLOADS["cfg2"]["text"]
    xkv = torch.randn(sum(text), Hkv, 256, device=dev, dtype=torch.bfloat16)
    cu = [0]
    for t in text:
        cu.append(cu[-1] + t)
    xkr = torch.tensor([[cu[i], cu[i + 1]] for i in range(4)], dtype=torch.int32, device=dev)
    xflops = 4 * 128 * Hq * clip * sum(text)
    res = {k: [] for k in libs}
    xres = {k: [] for k in libs}
    for rnd in range(args.rounds + 1):   # round 0 = warm-up of every variant (and of the power state)
        for name, path in libs.items():
            use(path)
            plan = F.attn_plan(qr, kr, num_heads_q=Hq, total_q=q.shape[0], total_k=kv.shape[0])
            fn = lambda: F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False, plan=plan)
            for _ in range(8):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.launches):
                fn()
            e1.record()
            xplan = F.attn_plan(qr, xkr, num_heads_q=Hq, total_q=q.shape[0], total_k=xkv.shape[0])
            xfn = lambda: F.attn_ranges(q, xkv[..., :128], xkv[..., 128:], qr, xkr, out=out, return_lse=False, plan=xplan)
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            xfn()
            x0.record()
            for _ in range(20):
                xfn()
            x1.record()
            torch.cuda.synchronize()
            if rnd > 0:
                res[name].append(e0.elapsed_time(e1) / args.launches)
                xres[name].append(x0.elapsed_time(x1) / 20 * 1e3)
    best = min(sum(v) / len(v) for v in res.values())
    for name in libs:
        v, x = res[name], xres[name]
        m = sum(v) / len(v)
        print(f"{name:14s} self {m:7.3f} ms = {flops / m / 1e9:7.1f} TFLOP/s ({100 * (m / best - 1):+5.1f} %)  rounds "
              + " ".join(f"{t:.3f}" for t in v) + f" | cross {sum(x) / len(x):6.1f} us = {xflops / (sum(x) / len(x)) / 1e6:6.1f} TFLOP/s",
              flush=True)


if __name__ == "__main__":
    main()
