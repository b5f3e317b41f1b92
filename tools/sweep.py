"""BASELINE.json configs[4]: attention microbench sweep on one B200 (run on the GPU box).

    python tools/sweep.py > gpurun_out/sweep.log

Chunk-causal self-attention (window-5 block-causal ranges: chunk i attends chunks [max(0, i-4), i]) for total token
counts 1.5k .. 256k split into chunks of 1536 / 9360 / 21600 tokens, head_dim 128, Hq in {24, 48}, Hkv = 8, packed
[K|V] input, and the T5-length cross-attention (every chunk of queries against its own caption of 50 / 256 / 800
tokens). FLOPs = 4 * 128 * Hq * sum_i Lq_i * Lk_i. CUDA events, 2 warm-up + best of 5; the
inputs exceed L2 from 24k tokens on; smaller cases are launch/latency bound and reported as latency."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import magi_b200.functional as F  # noqa: E402

CASES = [(1536, 1536), (6144, 1536), (24576, 1536), (65536, 9360), (131072, 21600), (262144, 21600)]


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = "cuda"
    i32 = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
    print("| tokens | chunk | chunks | Hq | self-attn TFLOP | ms | TFLOP/s | cross Lk=50 us | Lk=256 us | Lk=800 us (TFLOP/s) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for total, clip in CASES:
        n = max(1, total // clip)
        T = n * clip
        for Hq in (24, 48):
            q = torch.randn(T, Hq, 128, device=dev, dtype=torch.bfloat16)
            kv = torch.randn(T, 8, 256, device=dev, dtype=torch.bfloat16)
            out = torch.empty_like(q)
            qr = i32([[i * clip, (i + 1) * clip] for i in range(n)])
            kr = i32([[max(0, i - 4) * clip, (i + 1) * clip] for i in range(n)])
            flops = 4 * 128 * Hq * sum(clip * (i + 1 - max(0, i - 4)) * clip for i in range(n))
            ms = timed(lambda: F.attn_ranges(q, kv[..., :128], kv[..., 128:], qr, kr, out=out, return_lse=False))
            cross = []
            for lk in (50, 256, 800):
                tkv = torch.randn(n * lk, 8, 256, device=dev, dtype=torch.bfloat16)
                xkr = i32([[i * lk, (i + 1) * lk] for i in range(n)])
                xms = timed(lambda: F.attn_ranges(q, tkv[..., :128], tkv[..., 128:], qr, xkr, out=out, return_lse=False))
                cross.append((xms, 4 * 128 * Hq * T * lk / (xms * 1e-3) / 1e12))
            print(f"| {T} | {clip} | {n} | {Hq} | {flops / 1e12:.3f} | {ms:.3f} | {flops / (ms * 1e-3) / 1e12:.0f} | "
                  f"{cross[0][0] * 1e3:.0f} | {cross[1][0] * 1e3:.0f} | {cross[2][0] * 1e3:.0f} ({cross[2][1]:.0f}) |", flush=True)
            del q, kv, out
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
