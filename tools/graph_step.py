"""GPU box: the attention hot path of one denoise forward (bench.py's HotPath step) launched eagerly vs replayed from
ONE CUDA graph, for the launch-bound small configurations (BASELINE configs[0] = cfg1:
1536 tokens, 2 blocks, 7 launches per layer through ctypes). Every entry point of libmagi_attn.so takes the stream and
never synchronises or allocates, so the whole step is capturable (tests/test_gpu_parity.py::
test_hot_path_is_cuda_graph_capturable); this tool measures what the capture buys.

    python tools/graph_step.py [--workload cfg1] [--iters 50]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg1", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    w = bench.WORKLOADS[a.workload]
    torch.cuda.set_device(0)
    hp = bench.HotPath(w, torch.device("cuda", 0))
    eager_ms = timed(hp.step, a.iters)
    ref_core, ref_x = hp.core_out.clone(), hp.x_out.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):          # warm-up on the capture stream (workspace allocations happen here)
        hp.step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        hp.step()
    hp.core_out.zero_()
    hp.x_out.zero_()
    graph_ms = timed(graph.replay, a.iters)
    same = bool(torch.equal(hp.core_out, ref_core) and torch.equal(hp.x_out, ref_x))
    self_f, cross_f = bench.attn_flops(w)
    flops = (self_f + cross_f) * w["layers"]
    print(json.dumps({"workload": a.workload, "layers": w["layers"], "eager_ms_per_step": round(eager_ms, 4),
                      "graph_ms_per_step": round(graph_ms, 4), "speedup": round(eager_ms / graph_ms, 2),
                      "eager_tflops": round(flops / eager_ms / 1e9, 1), "graph_tflops": round(flops / graph_ms / 1e9, 1),
                      "replay_bit_identical_to_eager": same}))
    assert same, "graph replay differs from the eager step"
