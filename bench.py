#!/usr/bin/env python
"""bench.py - chunk-causal DiT attention hot path on B200 (BASELINE.json metric).

One "step" = the attention hot path of ONE denoise forward of the named workload over all transformer layers:
per layer the fused QK-LayerNorm+RoPE prologue (with KV-cache append), the range-masked chunk-causal self-attention
and the text cross-attention (its two prologues + varlen attention). The projection GEMMs around it are cuBLAS glue
outside the path and are not executed or counted.

Default workload (N = 1): BASELINE.json configs[1] - MAGI-1 4.5B-distill (34 layers, 24 q / 8 kv heads x 128),
480x832 -> 9360 tokens per chunk, 4 chunks with KV cache, the maximum-work state of the schedule: the first step of
a stage, which re-forwards the just-finished clean chunk (fwd_extra_1st_chunk) next to 3 denoising chunks and writes
all 4 to the KV cache: q ranges 4 x 9360, k ranges [(0,1),(0,2),(0,3),(0,4)] chunks, 10.77 TFLOP of self-attention per
layer. N > 1 runs the SAME workload context-parallel (Ulysses head all-to-all over NCCL): strong scaling.

    python bench.py                                  # N=1, prints one JSON line
    torchrun --nproc-per-node N bench.py --gpus N    # context parallel
    python bench.py --impl reference                 # the reference's CPU torch path (oracle port) on the host cores
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Hq, Hkv, layers, clip tokens, k ranges in chunk units (one q chunk each), caption tokens per chunk)
    "cfg2": dict(desc="MAGI-1 4.5B-distill 34L 24q/8kv x128, 480x832 (9360 tok/chunk), 4 chunks + KV cache, "
                      "max-work state k=[(0,1),(0,2),(0,3),(0,4)]", Hq=24, Hkv=8, layers=34, clip=9360,
                 k_chunks=[(0, 1), (0, 2), (0, 3), (0, 4)], text=[50, 512, 512, 512], slice_point=0),
    "cfg2-cache": dict(desc="cfg2, window advanced: 3 chunks in flight reading cached chunk 0, k=[(0,2),(0,3),(0,4)]",
                       Hq=24, Hkv=8, layers=34, clip=9360, k_chunks=[(0, 2), (0, 3), (0, 4)], text=[512, 512, 512],
                       slice_point=1),
    "cfg3": dict(desc="MAGI-1 4.5B-distill 34L, 480x832, pipelined denoising: 3 chunks in flight behind one cached chunk "
                      "with the noise2clean windows + the distill nearly-clean extra self-only range, "
                      "k=[(0,2),(0,3),(1,4),(4,5)] (mixed varlen ranges)", Hq=24, Hkv=8, layers=34, clip=9360,
                 k_chunks=[(0, 2), (0, 3), (1, 4), (4, 5)], text=[512, 512, 512, 512], slice_point=1),
    "cfg1": dict(desc="MAGI-1 4.5B cut to 2 blocks, 256x256 (1536 tok), 1 chunk", Hq=24, Hkv=8, layers=2, clip=1536,
                 k_chunks=[(0, 1)], text=[512], slice_point=0),
    "cfg4": dict(desc="MAGI-1 24B 48L 48q/8kv x128, 720x1280 (21600 tok/chunk), peak state 5 ranges "
                      "k=[(0,1),(0,2),(0,3),(0,4),(0,5)]", Hq=48, Hkv=8, layers=48, clip=21600,
                 k_chunks=[(0, 1), (0, 2), (0, 3), (0, 4), (0, 5)], text=[50, 512, 512, 512, 512], slice_point=0),
}


def attn_flops(w):
    """4 * hd * Hq * sum_i Lq_i * Lk_i over the unmasked rectangles; self + cross."""
    clip = w["clip"]
    self_f = 4 * 128 * w["Hq"] * sum(clip * (e - s) * clip for s, e in w["k_chunks"])
    cross_f = 4 * 128 * w["Hq"] * sum(clip * t for t in w["text"])
    return self_f, cross_f


# dram__bytes_read.sum + dram__bytes_write.sum of ONE self-attention launch of the named workload: NOT measured by the
# run that prints it (a number taken under a profiler is never a bench value) but read off the committed `ncu --set
# full` capture named next to it. The algorithmic minimum (q + out + packed kv once) is 230 + 230 + 153 = 613 MB: K/V
# re-reads by the 3 query heads of a GQA group and by the 37 query tiles of a chunk are served by L2.
NCU_TRAFFIC = {"cfg2": (824810240, "ncu capture profiles/r02zz_ncu_raw.csv (attn_fwd_kernel: 622.31 MB read + "
                                   "202.50 MB written per launch), not re-measured by this run")}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def prologue_bytes(w):
    """Algorithmic bytes of the single-GPU layer's ONE prologue launch (SURVEY.md section 8(d) extended by what rides in
    the launch): projection outputs in (q, k, v, qx, text [K|V]) + the sin / cos rows, and out: q, packed [K|V], its
    KV-cache copy, the cross-attention query and the normalised text [K|V]."""
    S, T = w["clip"] * len(w["k_chunks"]), sum(w["text"])
    hq, hkv = w["Hq"] * 128 * 2, w["Hkv"] * 128 * 2          # bytes of one row of q / of k (or v)
    read = S * (2 * hq + 2 * hkv + 96 * 4) + T * 2 * hkv
    write = S * (2 * hq + 2 * hkv + 2 * hkv) + T * 2 * hkv
    return read + write


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clock and throttle reasons of this rank's GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------- the hot path
class HotPath:
    """Synthetic inputs of one denoise forward's attention hot path, resident in HBM, and the per-layer launch
    sequence through the public API (magi_b200.functional -> C-ABI)."""

    def __init__(self, w, device, cp=1, rank=0, seed=1234, fused=True, ring=False, ring_fused=True, layers=None):
        import torch
        import magi_b200.functional as F
        if layers is not None:
            w = dict(w, layers=layers)
        self.torch, self.F, self.w, self.cp, self.rank, self.dev = torch, F, w, cp, rank, device
        self.ring = ring and cp > 1
        self.ring_fused = self.ring and ring_fused
        self.ring_multicast = False
        self.fused = fused and cp > 1 and not self.ring
        Hq, Hkv, clip = w["Hq"], w["Hkv"], w["clip"]
        self.S = clip * len(w["k_chunks"])
        from magi_b200 import context_parallel as cpm
        n_chunks = len(w["k_chunks"])
        if self.ring:
            # cp_ring_kv: every chunk striped across the ranks (context_parallel.ring_split_sizes)
            self.ring_sizes = cpm.ring_split_sizes(clip, cp)
            self.split = [n_chunks * s_ for s_ in self.ring_sizes]
        elif cp > 1:
            self.split = cpm.ulysses_split_sizes(self.S, cp)
        else:
            self.split = [self.S]
        self.S_loc = self.split[rank]
        off = sum(self.split[:rank])
        g = torch.Generator(device=device).manual_seed(seed + rank)     # this rank's token shard
        gw = torch.Generator(device=device).manual_seed(seed + 7919)    # replicated tensors: same on every rank

        def rnd(*shape, scale=1.0, dtype=torch.bfloat16, gen=None):
            return (torch.randn(*shape, device=device, generator=gen or g) * scale).to(dtype)

        # projection outputs of the LOCAL token shard (all heads) - the inputs of the hot path
        self.inputs = dict(xq=rnd(self.S_loc, Hq * 128), qx=rnd(self.S_loc, Hq * 128), xk=rnd(self.S_loc, Hkv * 128),
                           xv=rnd(self.S_loc, Hkv * 128), text_kv=rnd(sum(w["text"]), Hkv, 256, gen=gw))
        ang = torch.rand(self.S_loc, 48, device=device, generator=g) * 6.2831853
        self.inputs["sin"], self.inputs["cos"] = torch.sin(ang), torch.cos(ang)
        self.ln = {k: (rnd(128, scale=0.02, dtype=torch.float32, gen=gw), rnd(128, scale=0.02, dtype=torch.float32, gen=gw))
                   for k in ("q", "k")}
        self.lnx = {k: (rnd(128, scale=0.02, gen=gw), rnd(128, scale=0.02, gen=gw)) for k in ("q", "k")}
        # global (single-GPU order) row index of every local row: contiguous shard, or (chunk, stripe row) for ring-KV
        if self.ring:
            nr, so = self.ring_sizes[rank], sum(self.ring_sizes[:rank])
            self.global_rows = torch.cat([torch.arange(i * clip + so, i * clip + so + nr, device=device)
                                          for i in range(n_chunks)])
        else:
            self.global_rows = torch.arange(off, off + self.S_loc, device=device)
        i32 = lambda a: torch.tensor(a, dtype=torch.int32, device=device)
        sp = w["slice_point"]
        n = len(w["k_chunks"])
        self.q_ranges = i32([[i * clip, (i + 1) * clip] for i in range(n)])
        self.k_ranges = i32([[s * clip, e * clip] for s, e in w["k_chunks"]])
        # cross-attention segments restricted to the local shard (context_parallel.py:134-224)
        cu_k = [0]
        for t in w["text"]:
            cu_k.append(cu_k[-1] + t)
        xq_r, xk_r = [], []
        for i in range(n):
            a, b = max(i * clip, off), min((i + 1) * clip, off + self.S_loc)
            if self.ring:                         # local rows are (chunk, stripe row)
                nr = self.ring_sizes[rank]
                a, b = off + i * nr, off + (i + 1) * nr
            if a < b:
                xq_r.append([a - off, b - off])
                xk_r.append([cu_k[i], cu_k[i + 1]])
        self.x_q_ranges, self.x_k_ranges = i32(xq_r), i32(xk_r)
        # per-layer KV cache in post-all-to-all layout (dit_module.py:794-819)
        self.hkv_loc = Hkv if self.ring else max(1, Hkv // cp)   # ring-KV keeps all heads (and a replicated cache)
        self.hq_loc = Hq // cp
        self.cache_rows = (sp + n) * clip
        self.start = sp * clip
        if self.fused or self.ring_fused:
            # fused all-to-alls / all-gather: caches and receive buffers live in symmetric (NVLink peer-mapped) memory
            import torch.distributed._symmetric_memory as symm
            from magi_b200 import parallel_state as mpu
            self.caches, self.cache_ptrs, self.cache_mc = [], [], []
            for _ in range(w["layers"]):
                c = symm.empty((self.cache_rows, self.hkv_loc, 256), dtype=torch.bfloat16, device=device)
                h_ = symm.rendezvous(c, mpu.get_cp_group().group_name)
                self.cache_ptrs.append([int(p_) for p_ in h_.buffer_ptrs])
                self.cache_mc.append(int(getattr(h_, "multicast_ptr", 0) or 0))
                self.caches.append(c)
        if self.ring_fused:
            self.ring_ws = cpm.RingGatherWorkspace.get(mpu.get_cp_group(), self.S, Hkv, device)
            self.ring_multicast = cpm.ring_multicast_enabled() and all(self.cache_mc)
            e = torch.arange(n + 1, dtype=torch.int32) * self.ring_sizes[rank]
            self.stripe_q_ranges = torch.stack([e[:-1], e[1:]], 1).contiguous().to(device)
        elif self.fused:
            self.ws = cpm.UlyssesPeerWorkspace.get(mpu.get_cp_group(), self.S, max(self.split), self.hq_loc,
                                                   self.hkv_loc, Hq, device)
            self.row_off, self.bounds, self.head_off = cpm.fused_ulysses_plan(self.split, rank, Hq, cp)
        else:
            self.caches = [torch.empty(self.cache_rows, self.hkv_loc, 256, dtype=torch.bfloat16, device=device)
                           for _ in range(w["layers"])]
        if sp > 0:
            for c in self.caches:
                c[:self.start].normal_()
        self.core_out = torch.empty(self.S_loc, Hq, 128, dtype=torch.bfloat16, device=device)
        self.x_out = torch.empty(self.S_loc, Hq, 128, dtype=torch.bfloat16, device=device)
        self.attn_events = []
        self._plans = {}
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.inputs.values())
        self.d2h_bytes = 2 * self.core_out.numel() * 2

    # phase marks (diagnostic steps only, phase_breakdown_leg): an event after every phase of a layer
    marks = None

    def mark(self, name):
        if self.marks is not None:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    # -- one layer, single GPU
    def layer(self, l, timed_events=False):
        F, w, inp, torch = self.F, self.w, self.inputs, self.torch
        cache = self.caches[l]
        # ONE prologue launch per layer: q / k / v (+ KV-cache append) and the cross-attention's qx / text k / text v.
        # Everything the cross-attention needs (and both plans) thereby exists BEFORE the self-attention launch, so the
        # cross-attention launch that follows it depends on nothing the self-attention kernel touches
        cross = self.cross_args(inp)
        q, kv = F.qk_norm_rope(inp["xq"], inp["xk"], inp["xv"], w["Hq"], w["Hkv"], self.ln["q"], self.ln["k"],
                               inp["sin"], inp["cos"], cache=cache, cache_row_offset=self.start, cache_rows=self.S,
                               cross=cross)
        self.mark("prologue")
        kc = vc = None
        if self.start > 0:
            kc, vc = cache[:self.start, :, :128], cache[:self.start, :, 128:]
        x_prepared = self.cross_prepared(cross)
        plan = self.plan("self", q, self.start + kv.shape[0], self.q_ranges, self.k_ranges)
        if timed_events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        F.attn_ranges(q, kv[..., :128], kv[..., 128:], self.q_ranges, self.k_ranges, k_cache=kc, v_cache=vc,
                      out=self.core_out, return_lse=False, plan=plan)
        if timed_events:
            e1.record()
            self.attn_events.append((e0, e1))
        self.mark("self_attention")
        self.cross_attend(x_prepared, independent=not timed_events and self.marks is None)
        self.mark("cross_attention")

    def cross_args(self, inp):
        """`cross=` of F.qk_norm_rope: the cross-attention's tensors ride in the layer's one prologue launch."""
        return dict(x_qx=inp["qx"], ln_qx=self.lnx["q"], text_kv=inp["text_kv"], ln_tk=self.lnx["k"])

    def cross_prepared(self, cross):
        xq, xkv = cross["qx"], cross["tkv"]
        return xq, xkv, self.plan("cross", xq, xkv.shape[0], self.x_q_ranges, self.x_k_ranges)

    def cross_prepare(self, inp):
        """Two separate prologue launches (the NCCL paths, whose q and kv prologues are separate launches already)."""
        F, w = self.F, self.w
        xq, _ = F.qk_norm_rope(inp["qx"], None, None, w["Hq"], w["Hkv"], self.lnx["q"], None, None, None)
        tkv = inp["text_kv"]
        _, xkv = F.qk_norm_rope(None, tkv[..., :128], tkv[..., 128:], w["Hq"], w["Hkv"], None, self.lnx["k"], None,
                                None)
        return xq, xkv, self.plan("cross", xq, xkv.shape[0], self.x_q_ranges, self.x_k_ranges)

    def cross_attend(self, prepared, independent=False):
        """independent: this launch directly follows the self-attention of the same layer and shares no tensor with
        it (inputs, plan and self.x_out all exist since before that launch): it may overlap the self-attention's tail
        (magi_attn_args.launch_dependency). Off while per-launch events bracket the self-attention kernel."""
        xq, xkv, plan = prepared
        self.F.attn_ranges(xq, xkv[..., :128], xkv[..., 128:], self.x_q_ranges, self.x_k_ranges, out=self.x_out,
                           return_lse=False, plan=plan, independent_of_prev=independent and self.pdl_overlap)

    pdl_overlap = os.environ.get("MAGI_BENCH_PDL_OVERLAP", "1") != "0"

    def cross(self, inp):
        self.cross_attend(self.cross_prepare(inp))

    # -- one layer, Ulysses context parallel (dit_module.py:1175-1205 / context_parallel.py:503-597)
    def layer_cp(self, l, timed_events=False):
        from magi_b200 import context_parallel as cpm
        F, w, inp, torch, cp = self.F, self.w, self.inputs, self.torch, self.cp
        cache = self.caches[l]
        _, kv_send = F.qk_norm_rope(None, inp["xk"], inp["xv"], w["Hq"], w["Hkv"], None, self.ln["k"], inp["sin"],
                                    inp["cos"], head_perm_cp=cp)
        # the KV all-to-all lands directly in the cache slot (every row of this state is cached)
        kv, h_kv = cpm.all_to_all_input_split(kv_send, self.split, send_layout=True,
                                              out=cache[self.start:self.start + self.S])
        q_send, _ = F.qk_norm_rope(inp["xq"], None, None, w["Hq"], w["Hkv"], self.ln["q"], None, inp["sin"],
                                   inp["cos"], head_perm_cp=cp)
        q, h_q = cpm.all_to_all_input_split(q_send, self.split, send_layout=True)
        x_prepared = self.cross_prepare(inp)   # overlaps the two input all-to-alls
        h_kv.wait()
        h_q.wait()
        kc = vc = None
        if self.start > 0:
            kc, vc = cache[:self.start, :, :128], cache[:self.start, :, 128:]

        def core(qg, kg, vg, out, kv_heads):
            a, b = kv_heads
            if timed_events:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            F.attn_ranges(qg, kg, vg, self.q_ranges, self.k_ranges, k_cache=None if kc is None else kc[:, a:b],
                          v_cache=None if vc is None else vc[:, a:b], out=out, return_lse=False,
                          plan=self.plan("self", qg, self.start + kg.shape[0], self.q_ranges, self.k_ranges))
            if timed_events:
                e1.record()
                self.attn_events.append((e0, e1))
            return out

        def relayout_out(recv, cp_, result, q_head, lo):
            F.ulysses_relayout(recv, cp_, 1, out=result, dst_heads_total=q_head, dst_head_offset=lo)

        core_out, _ = cpm.UlyssesScheduler.get_attn_and_xattn_base(
            q, kv[..., :128], kv[..., 128:], core, lambda: self.cross_attend(x_prepared), self.ulysses_overlap_degree, 1, cp,
            self.split, relayout_out=relayout_out)
        self.core_out = core_out

    # -- one layer, Ulysses with both all-to-alls fused into the kernels' stores (FullyParallelAttention.
    #    _forward_ulysses_fused): one prologue launch writes q and [K|V] into the owners' buffers (K/V straight into
    #    the cache slot), the attention epilogue writes every output row into its owner's buffer; two barriers
    def layer_cp_fused(self, l, timed_events=False):
        F, w, inp, torch, cp = self.F, self.w, self.inputs, self.torch, self.cp
        cache, ws = self.caches[l], self.ws
        row_bytes = self.hkv_loc * 256 * 2
        cross = self.cross_args(inp)
        F.qk_norm_rope(inp["xq"], inp["xk"], inp["xv"], w["Hq"], w["Hkv"], self.ln["q"], self.ln["k"], inp["sin"],
                       inp["cos"], head_perm_cp=cp, peer_q_out=ws.q_ptrs,
                       peer_kv_out=[p_ + self.start * row_bytes for p_ in self.cache_ptrs[l]],
                       peer_row_offset=self.row_off, peer_heads=(self.hq_loc, self.hkv_loc), cross=cross)
        self.mark("prologue")
        x_prepared = self.cross_prepared(cross)
        ws.barrier()
        self.mark("barrier_a")
        kv = cache[self.start:self.start + self.S]
        kc = vc = None
        if self.start > 0:
            kc, vc = cache[:self.start, :, :128], cache[:self.start, :, 128:]
        if timed_events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        F.attn_ranges(ws.q, kv[..., :128], kv[..., 128:], self.q_ranges, self.k_ranges, k_cache=kc, v_cache=vc,
                      return_lse=False, peer_out=(ws.out_ptrs, self.bounds, self.head_off, w["Hq"]),
                      plan=self.plan("self", ws.q, self.start + kv.shape[0], self.q_ranges, self.k_ranges))
        if timed_events:
            e1.record()
            self.attn_events.append((e0, e1))
        self.mark("self_attention")
        self.cross_attend(x_prepared, independent=not timed_events and self.marks is None)
        self.mark("cross_attention")
        ws.barrier()
        self.mark("barrier_b")
        self.core_out = ws.out[:self.S_loc]

    # -- one layer, ring-KV (FullyParallelAttention._forward_ring): every rank keeps its query shard with all heads, the
    #    packed K/V blocks of the new tokens travel round the ring (NCCL send/recv over NVLink, next block in flight
    #    while the current one is attended), partial results merged by log-sum-exp; the cache is replicated
    def layer_cp_ring(self, l, timed_events=False):
        from magi_b200 import context_parallel as cpm
        F, w, inp, torch = self.F, self.w, self.inputs, self.torch
        cache = self.caches[l]
        q, kv_local = F.qk_norm_rope(inp["xq"], inp["xk"], inp["xv"], w["Hq"], w["Hkv"], self.ln["q"], self.ln["k"],
                                     inp["sin"], inp["cos"])
        dev = self.dev

        def attn(q_, blk, cache_prefix, qr, kr):
            t = lambda a: torch.as_tensor(a, dtype=torch.int32).to(dev, non_blocking=True)
            kc = vc = None
            if cache_prefix is not None:
                kc, vc = cache_prefix[..., :128], cache_prefix[..., 128:]
            if timed_events:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            r = F.attn_ranges(q_, blk[..., :128], blk[..., 128:], t(qr), t(kr), k_cache=kc, v_cache=vc, zero_fill=True)
            if timed_events:
                e1.record()
                self.attn_events.append((e0, e1))
            return r

        n = len(w["k_chunks"])
        out, _ = cpm.RingKVScheduler.run(q, kv_local, cache, w["k_chunks"], w["slice_point"], w["clip"], n,
                                         self.ring_sizes, n, attn, F.lse_merge)
        self.core_out = out
        self.cross(inp)

    # -- one layer, ring-KV with the exchange fused into the prologue (FullyParallelAttention._forward_ring_fused): the
    #    K/V stores of the prologue ARE the all-gather (NVLink peer stores or one NVSwitch multicast store per row),
    #    straight into every rank's replicated cache slot; one device barrier; one attention launch per rank
    def layer_cp_ring_fused(self, l, timed_events=False):
        F, w, inp, torch = self.F, self.w, self.inputs, self.torch
        cache = self.caches[l]
        slot = self.start * w["Hkv"] * 256 * 2
        targets = [self.cache_mc[l] + slot] if self.ring_multicast else [p_ + slot for p_ in self.cache_ptrs[l]]
        cross = self.cross_args(inp)
        q, _ = F.qk_norm_rope(inp["xq"], inp["xk"], inp["xv"], w["Hq"], w["Hkv"], self.ln["q"], self.ln["k"],
                              inp["sin"], inp["cos"], gather_kv_out=targets,
                              gather_geometry=(self.ring_sizes[self.rank], w["clip"], sum(self.ring_sizes[:self.rank])),
                              gather_multicast=self.ring_multicast, cross=cross)
        x_prepared = self.cross_prepared(cross)
        self.ring_ws.barrier()
        kv = cache[self.start:self.start + self.S]
        kc = vc = None
        if self.start > 0:
            kc, vc = cache[:self.start, :, :128], cache[:self.start, :, 128:]
        if timed_events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        F.attn_ranges(q, kv[..., :128], kv[..., 128:], self.stripe_q_ranges, self.k_ranges, k_cache=kc, v_cache=vc,
                      out=self.core_out, return_lse=False,
                      plan=self.plan("self", q, self.start + kv.shape[0], self.stripe_q_ranges, self.k_ranges))
        if timed_events:
            e1.record()
            self.attn_events.append((e0, e1))
        self.cross_attend(x_prepared, independent=not timed_events)

    ulysses_overlap_degree = 1

    def plan(self, kind, q, total_k, qr, kr):
        """The work list of a range table is built ONCE per step (= per denoise forward, all layers share the ranges:
        magi_attn_plan) and reused by every layer's launch."""
        key = (kind, q.shape[1], q.shape[0], total_k)
        pl = self._plans.get(key)
        if pl is None:
            pl = self._plans[key] = self.F.attn_plan(qr, kr, num_heads_q=q.shape[1], total_q=q.shape[0], total_k=total_k)
        return pl

    def layer_fn(self):
        return (self.layer_cp_ring_fused if self.ring_fused else self.layer_cp_ring if self.ring else
                self.layer_cp_fused if self.fused else self.layer_cp) if self.cp > 1 else self.layer

    def step(self, timed_events=False):
        fn = self.layer_fn()
        self._plans = {}
        for l in range(self.w["layers"]):
            fn(l, timed_events and l % self.event_stride == 0)

    # Per-launch CUDA events bracket the self-attention kernel on every `event_stride`-th layer of a timed step only:
    # an event record between two launches serialises them, which would switch off the overlap of the cross-attention
    # launch with the self-attention's tail (HotPath.cross_attend) on exactly the launches being measured.
    event_stride = 4

    def event_layers(self):
        return len(range(0, self.w["layers"], self.event_stride))

    def mode(self):
        return "none" if self.cp == 1 else ("ring-fused" if self.ring_fused else "ring-nccl" if self.ring else
                                            "ulysses-fused" if self.fused else "ulysses-nccl")


# ----------------------------------------------------------------------------------------------- whole DiT blocks
def dit_step_leg(w, dev, world, rank, steps, max_over_ranks, barrier, fp8=False):
    """BASELINE.json's second metric, "DiT denoise-step latency": one forward of ALL transformer blocks of the named
    model on the same state - the attention hot path (this repository's kernels) plus everything around it in the
    block (hidden LayerNorms, q/qx/k/v, text-kv, out-projection and MLP GEMMs through cuBLAS, the fused adaLN gating
    chains) via the drop-in TransformerLayer (magi-1_b200/transformer_layer.py); random-init weights of the real
    architecture, synthetic latents / captions / timestep embeddings. N > 1: sequence-sharded blocks, Ulysses
    attention with the all-to-alls fused into the kernels.
    fp8=True: engine_config.fp8_quant (the *_distill_quant configs): every layer but the first and the last holds fp8
    q / qx / k / v (fused tcgen05 kind::f8f6f4 GEMM with the prologue as its epilogue), out-projection, fc1 and fc2
    (activation quantiser kernel + cuBLASLt fp8 matmul, the library GEMM the reference calls through bmm_fp8)."""
    import numpy as np
    import torch
    from magi_b200 import context_parallel as cpm
    from magi_b200.common import (EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCoreAttnParams,
                                  PackedCrossAttnParams)
    from magi_b200.transformer_layer import TransformerLayer, high_precision_promote
    big = w["Hq"] == 48
    mc = (ModelConfig.magi_24b if big else ModelConfig.magi_4p5b)(num_layers=w["layers"])
    ec = EngineConfig(cp_size=world, cp_strategy="none" if world == 1 else "cp_ulysses", fp8_quant=fp8)
    clip, n, sp = w["clip"], len(w["k_chunks"]), w["slice_point"]
    S = clip * n
    torch.manual_seed(1234)
    layers = []
    for l in range(w["layers"]):
        with torch.device(dev):
            layer = high_precision_promote(TransformerLayer(mc, ec, layer_number=l))
        for name, p_ in layer.named_parameters():
            if p_.dtype == torch.float8_e4m3fn:      # a calibrated-checkpoint-like fp8 linear: |w| <= 448 * weight_scale
                p_.data.copy_((torch.randn(p_.shape, device=dev) * 100.0).clamp_(-448.0, 448.0).to(p_.dtype))
            elif name.endswith("weight_scale"):
                p_.data.fill_(0.01 / math.sqrt(dict(layer.named_parameters())[name[:-6]].shape[-1]))
            elif name.endswith("input_scale"):
                p_.data.fill_(6.0 / 448.0)
            elif name.endswith("smooth_scale"):
                p_.data.fill_(1.0)
            elif "layernorm" in name or "post_norm" in name or name.endswith("bias"):
                p_.data.normal_(0.0, 0.02)
            elif name.endswith("layer_norm.weight"):
                p_.data.fill_(1.0)
            else:
                p_.data.normal_(0.0, 1.0 / math.sqrt(p_.shape[-1]))
        layers.append(layer.eval())
    g = torch.Generator(device=dev).manual_seed(99)
    hs = torch.randn(S, 1, mc.hidden_size, device=dev, generator=g).to(torch.bfloat16)
    cond = torch.randn(1, n, int(mc.hidden_size * mc.cond_hidden_ratio), device=dev, generator=g).to(torch.bfloat16)
    cmap = torch.arange(n, device=dev).repeat_interleave(clip).reshape(1, -1).transpose(0, 1).contiguous()
    y = torch.randn(sum(w["text"]), mc.hidden_size, device=dev, generator=g).to(torch.bfloat16)
    ang = torch.rand(S, 48, device=dev, generator=g) * 6.2831853
    rope = torch.cat([torch.sin(ang), torch.cos(ang)], -1)
    i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32), device=dev)
    q_range = np.array([[i * clip, (i + 1) * clip] for i in range(n)], dtype=np.int32)
    k_range = np.array([[s_ * clip, e_ * clip] for s_, e_ in w["k_chunks"]], dtype=np.int32)
    cu_q = np.arange(n + 1, dtype=np.int32) * clip
    cu_k = np.concatenate([[0], np.cumsum(w["text"])]).astype(np.int32)
    core = PackedCoreAttnParams(i32(q_range), i32(k_range), q_range, k_range, clip, int(k_range.max() - k_range.min()))
    cross = PackedCrossAttnParams(i32(np.stack([cu_q[:-1], cu_q[1:]], 1)), i32(np.stack([cu_k[:-1], cu_k[1:]], 1)),
                                  i32(cu_q), i32(cu_k), clip, 800)
    split = None
    if world > 1:
        ardf = dict(denoising_range_num=n, q_range=core.q_range, k_range=core.k_range, max_seqlen_q=clip,
                    max_seqlen_k=core.max_seqlen_k)
        hs, cmap, rope, _, split, core, cross = cpm.cp_pre_process(world, ec.cp_strategy, hs, cmap, rope, None, ardf,
                                                                   core, cross)
    meta = ModelMetaArgs(H=1, W=1, cp_pad_size=0, cp_split_sizes=split, slice_point=sp, denoising_range_num=n,
                         range_num=sp + n, extract_prefix_video_feature=False, fwd_extra_1st_chunk=True,
                         distill_nearly_clean_chunk=False, clip_token_nums=clip, enable_cuda_graph=False,
                         core_attn_params=core, cross_attn_params=cross)
    inf = InferenceParams(max_batch_size=1, max_sequence_length=(sp + n) * clip)
    inf.update_kv_cache = True

    def step():
        x = hs
        for layer in layers:
            x = layer(x, cond, cmap, y, rope, inf, meta)
        return x

    for _ in range(2):
        out = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    # the same stack as ONE CUDA graph per forward signature (EngineConfig.enable_cuda_graph, cuda_graph.py)
    graph_ms = graph_note = None
    if world == 1:
        from magi_b200.cuda_graph import GraphedLayerStack
        ec.enable_cuda_graph = True
        stack = GraphedLayerStack(layers, ec)
        for _ in range(2):
            out_g = stack(hs, cond, cmap, y, rope, inf, meta)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(steps):
            out_g = stack(hs, cond, cmap, y, rope, inf, meta)
        g1.record()
        barrier()
        if stack.disabled_reason is None:
            graph_ms = max_over_ranks(g0.elapsed_time(g1)) / steps
            d_ = (out_g.float() - out.float()).abs().max().item()
            graph_note = "bit-identical to eager" if d_ == 0.0 else (
                f"max |graph - eager| = {d_:.3e} on outputs of magnitude {out.float().abs().max().item():.1f}: cuBLAS "
                f"picks other GEMM algorithms under stream capture (the repository's own kernels replay bit-identically: "
                f"tests/test_gpu_parity.py::test_graphed_layer_stack_replays_bit_identical_to_eager)")
        else:
            graph_note = "capture failed, ran eagerly: " + stack.disabled_reason
        del stack
    h, f = mc.hidden_size, mc.ffn_hidden_size
    per_tok = 2 * h * (2 * w["Hq"] * 128 + 2 * w["Hkv"] * 128) + 2 * (2 * w["Hq"] * 128) * h \
        + 2 * h * f * (3 if mc.gated_linear_unit else 2)
    gemm = (per_tok * S + 2 * sum(w["text"]) * h * 2 * w["Hkv"] * 128) * w["layers"]
    self_f, cross_f = attn_flops(w)
    total = gemm + (self_f + cross_f) * w["layers"]
    assert bool(torch.isfinite(out.float()).all())
    del layers
    torch.cuda.empty_cache()
    return {"ms_per_step": round(ms, 2), "graph_ms_per_step": None if graph_ms is None else round(graph_ms, 2),
            "graph_note": graph_note, "layers": w["layers"], "model": "MAGI-1 24B" if big else "MAGI-1 4.5B",
            "tflops_total": round(total / (ms * 1e-3) / 1e12, 1), "gemm_tflop_per_step": round(gemm / 1e12, 1),
            "attn_tflop_per_step": round((self_f + cross_f) * w["layers"] / 1e12, 1),
            "what": "all transformer blocks of one denoise forward: attention hot path (own kernels) + cuBLAS GEMMs + "
                    "fused adaLN gating chains, random-init weights, " + ("single GPU" if world == 1 else
                                                                         f"cp_ulysses{world} fused all-to-alls")
                    + (", fp8_quant: fp8 linears in every layer but the first and the last" if fp8 else "")}


# ----------------------------------------------------------------------------------------------- GPU reference leg
def gpu_reference_leg(w, dev, iters=3):
    """B-REF-GPU ("the number to beat"): what the unmodified reference executes on this box for
    the self-attention of one layer - its fallback branch, a Python loop of flash_attn_func over the denoising ranges
    on `.contiguous()` K / V slices (dit_module.py:1025-1029, 1046-1065), with the flash-attn wheel of this image
    (library code, used here ONLY as the measured baseline; nothing in the product imports it). MagiAttention, the
    reference's preferred kernel, is not installable here. Same q / k / v shapes as the timed
    workload, CUDA events, L2 not flushed between iterations (q + K/V + out of one layer are several times the 126 MB L2)."""
    import torch
    try:
        from flash_attn import flash_attn_func
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"flash_attn import failed: {e}"[:200]}
    clip, Hq, Hkv = w["clip"], w["Hq"], w["Hkv"]
    n = len(w["k_chunks"])
    k_hi = max(e for _, e in w["k_chunks"])
    g = torch.Generator(device=dev).manual_seed(4242)
    q = torch.randn(1, n * clip, Hq, 128, device=dev, generator=g).to(torch.bfloat16)
    kv = torch.randn(1, k_hi * clip, Hkv, 256, device=dev, generator=g).to(torch.bfloat16)

    def layer():
        # the reference materialises contiguous key / value from the packed tensor on every forward (dit_module.py:945-946)
        k, v = kv[..., :128].contiguous(), kv[..., 128:].contiguous()
        outs = []
        for i, (s_, e_) in enumerate(w["k_chunks"]):
            outs.append(flash_attn_func(q[:, i * clip:(i + 1) * clip], k[:, s_ * clip:e_ * clip],
                                        v[:, s_ * clip:e_ * clip]))
        return torch.cat(outs, dim=1)

    try:
        layer()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            layer()
        e1.record()
        torch.cuda.synchronize()
    except Exception as e:  # pragma: no cover - e.g. a wheel without sm_100 kernels
        return {"unavailable": f"flash_attn_func failed: {e}"[:200]}
    ms = e0.elapsed_time(e1) / iters
    self_f, _ = attn_flops(w)
    import flash_attn
    return {"kind": f"flash-attn {flash_attn.__version__} wheel, the reference's fallback loop of flash_attn_func over "
                    f"{n} ranges (dit_module.py:1046-1065) incl. its K/V .contiguous() copies",
            "self_attn_ms_per_layer": round(ms, 3), "tflops": round(self_f / (ms * 1e-3) / 1e12, 1), "iters": iters}


# ----------------------------------------------------------------------------------------------- CPU baseline
def cpu_baseline_sample(w, budget_s=20.0):
    """The reference's CPU torch path on a bounded sample of the SAME workload: the self-attention of layer 0 of the
    workload state through oracle/torch_cpu_path.py (the reference's fallback loop of flash_attn_func calls,
    dit_module.py:1046-1065, restated with torch fp32 scaled_dot_product_attention on all host threads - flash-attn
    itself is CUDA-only), ranges taken in order until the time budget is spent. Returns (tflops, dict)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch_cpu_path as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    flops, secs, done, threads = T.timed_self_attention_sample(w["clip"], w["Hq"], w["Hkv"], w["k_chunks"], budget_s)
    tf = flops / secs / 1e12
    ks = ",".join(f"({s},{e})" for s, e in w["k_chunks"][:done])
    return tf, {"value": round(tf, 4), "unit": "TFLOP/s", "cores": threads, "kind": "port",
                "sample": f"self-attention of layer 0, ranges 0..{done - 1} of {len(w['k_chunks'])} (k chunks {ks}; "
                          f"{w['clip']} q rows each, all {w['Hq']} q heads): {flops / 1e12:.2f} TFLOP in {secs:.1f} s, "
                          f"torch fp32 CPU port of the reference's flash_attn_func loop (oracle/torch_cpu_path.py)"}


def parallelism_desc(world, mode, overlap=1, multicast=False):
    if world == 1:
        return "none"
    return {
        "ulysses-fused": f"cp_ulysses{world}, all-to-alls fused into the kernels' stores (NVLink peer memory via symmetric "
                         f"memory, no NCCL on the data path), 2 device barriers per layer",
        "ulysses-nccl": f"cp_ulysses{world}, NCCL all_to_all_single (overlap_degree {overlap})",
        "ring-fused": f"cp_ring_kv{world} (every chunk striped across the ranks), K/V all-gather fused into the prologue's "
                      f"stores ({'NVSwitch multicast (multimem.st)' if multicast else 'NVLink peer stores'}) into the "
                      f"replicated cache slot, 1 device barrier per layer",
        "ring-nccl": f"cp_ring_kv{world}, K/V blocks round the ring by NCCL send/recv, log-sum-exp merge, replicated cache",
    }[mode]


def static_config(args, w, world):
    """The `config` object: names the workload. Identical in both arms (ours / --impl reference); everything measured
    lives in `roofline`, `e2e`, `clocks`, `cpu_baseline`."""
    self_f, cross_f = attn_flops(w)
    mode = "none" if world == 1 else {"fused": "ulysses-fused", "nccl": "ulysses-nccl", "ring": "ring-fused",
                                      "ring-nccl": "ring-nccl"}[args.cp_mode]
    return {"workload": w["desc"], "name": args.workload,
            "parallelism": parallelism_desc(world, mode, args.overlap_degree, True),
            "step": "attention hot path of one denoise forward over all layers, three launches per layer: fused LN+RoPE "
                    "prologue (q / k / v with KV-cache append + the cross-attention's qx / text k / v), range-masked "
                    "self-attention, text cross-attention",
            "tokens_q": w["clip"] * len(w["k_chunks"]), "layers": w["layers"],
            "flops_per_step": (self_f + cross_f) * w["layers"], "self_attn_flops_per_layer": self_f,
            "l2": "inputs larger than L2 (>= 600 MB of projection outputs re-read per layer)"}


def run_reference(args, w):
    """--impl reference: the reference's CPU implementation of the path on the host cores, same metric/config: the
    torch fp32 port of its attention fallback loop (oracle/torch_cpu_path.py; the reference is Python whose CUDA-only
    third-party kernels cannot run on a CPU). Each step times a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    self_f, cross_f = attn_flops(w)
    times, tfs = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        tf, info = cpu_baseline_sample(w, budget_s=args.ref_budget)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
            tfs.append(tf)
    tf = sum(tfs) / len(tfs)
    step_flops = (self_f + cross_f) * w["layers"]
    info["value"] = round(tf, 4)
    info["sample"] += "; ms_per_step is extrapolated from the timed sample to the whole step"
    line = {"impl": "reference", "metric": "chunk_causal_attn_tflops", "value": round(tf, 4), "unit": "TFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(step_flops / (tf * 1e12) * 1e3, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config(args, w, args.gpus),
            "cpu_baseline": info, "e2e": {"value": round(tf, 4), "unit": "TFLOP/s", "h2d_bytes_per_step": 0,
                                          "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- measurement legs
class Ctx:
    """Rank / device plumbing shared by the legs."""

    def __init__(self, torch, dist, dev, world, rank, local):
        self.torch, self.dist, self.dev, self.world, self.rank, self.local = torch, dist, dev, world, rank, local

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timed_leg(cx, hp, steps, warm, graph=False):
    """W warm-up steps, then exactly `steps` steps between barrier + synchronize, CUDA events, max over ranks.
    graph=True: the step is captured ONCE into a CUDA graph after the warm-up (every launch of the hot path is
    capturable: no host reads, no syncs) and the timed region replays it `steps` times; the per-launch times of the
    self-attention kernel (roofline) then come from two eager steps with events right after the timed region."""
    from magi_b200 import _lib
    torch = cx.torch
    for _ in range(warm):
        hp.step()
    cx.barrier()
    replay = None
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            hp.step()
        torch.cuda.current_stream().wait_stream(side)
        replay = g.replay
        replay()
        cx.barrier()
    hp.attn_events = []
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(cx.local) as clk:
        e0.record()
        for _ in range(steps):
            if replay is not None:
                replay()
            else:
                hp.step(timed_events=True)
        e1.record()
        cx.barrier()
    launches = _lib.launch_count() - l0
    ms_total = cx.max_over_ranks(e0.elapsed_time(e1))
    events_steps = steps
    if replay is not None:
        l1 = _lib.launch_count()
        events_steps = 2
        for _ in range(events_steps):
            hp.step(timed_events=True)
        cx.barrier()
        launches = (_lib.launch_count() - l1) // events_steps * steps   # what one replayed step launches, x steps
    attn_ms = [a.elapsed_time(b) for a, b in hp.attn_events]
    return dict(ms_step=ms_total / steps, ms_total=ms_total, attn_ms=attn_ms, launches=int(launches),
                clocks=clk.summary(), events_steps=events_steps, graph=replay is not None)


def phase_breakdown_leg(cx, hp, steps=2):
    """Diagnostic, after the timed region: `steps` eager steps with an event after every phase of every layer; average
    microseconds per layer and phase, max over ranks. The events serialise the launches (no overlap of the
    cross-attention with the self-attention's tail), so the sum is an upper bound of the timed step's layer time."""
    torch = cx.torch
    if hp.layer_fn() not in (hp.layer, hp.layer_cp_fused):
        return None
    hp.marks = []
    cx.barrier()
    tot, cnt = {}, {}
    for _ in range(steps):
        hp.marks.append(("start", None))
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        hp.marks[-1] = ("start", e)
        hp.step()
    cx.barrier()
    prev = None
    for name, e in hp.marks:
        if name != "start" and prev is not None:
            tot[name] = tot.get(name, 0.0) + prev.elapsed_time(e) * 1e3
            cnt[name] = cnt.get(name, 0) + 1
        prev = e
    hp.marks = None
    out = {k: round(cx.max_over_ranks(tot[k] / cnt[k]), 1) for k in tot}
    out["sum"] = round(sum(out.values()), 1)
    return out


def e2e_leg(cx, hp, steps):
    """End to end through the public API with HOST buffers: every step copies all of its inputs in from pinned host
    memory and all of its results out, inside the timed region. Inputs and outputs are double-buffered on one copy
    stream at every N: the H2D copy of step t + 1 and the D2H read of step t run under the kernels of the
    neighbouring step. Under context parallelism the attention output lives in the (single) symmetric-memory
    receive buffer, so it is first copied to the step's private output set on the compute stream."""
    torch = cx.torch
    host_in = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in hp.inputs.items()}
    host_out = [torch.empty(hp.core_out.numel(), dtype=torch.bfloat16, pin_memory=True),
                torch.empty(hp.x_out.numel(), dtype=torch.bfloat16, pin_memory=True)]
    copy = torch.cuda.Stream()
    in_sets = [hp.inputs, {k: torch.empty_like(v) for k, v in hp.inputs.items()}]
    out_sets = [(torch.empty_like(hp.core_out), torch.empty_like(hp.x_out)) for _ in range(2)]
    if cx.world == 1:
        out_sets[0] = (hp.core_out, hp.x_out)
    h2d_done, compute_done, d2h_done = [None, None], [None, None], [None, None]

    def issue_h2d(i):
        with torch.cuda.stream(copy):
            if compute_done[i] is not None:
                copy.wait_event(compute_done[i])      # set i was last read by the step before last
            for k, v in host_in.items():
                in_sets[i][k].copy_(v, non_blocking=True)
            h2d_done[i] = torch.cuda.Event()
            h2d_done[i].record(copy)

    def e2e_step(t):
        i = t & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(h2d_done[i])
        if d2h_done[i] is not None:
            cur.wait_event(d2h_done[i])               # output set i has been read back
        hp.inputs = in_sets[i]
        if cx.world == 1:
            hp.core_out, hp.x_out = out_sets[i]       # the kernels write the step's output set directly
            hp.step()
        else:
            hp.x_out = out_sets[i][1]
            hp.step()
            out_sets[i][0].copy_(hp.core_out)         # out of the shared receive buffer (device copy, compute stream)
        compute_done[i] = torch.cuda.Event()
        compute_done[i].record(cur)
        issue_h2d(i ^ 1)
        with torch.cuda.stream(copy):
            copy.wait_event(compute_done[i])
            host_out[0].copy_(out_sets[i][0].reshape(-1), non_blocking=True)
            host_out[1].copy_(out_sets[i][1].reshape(-1), non_blocking=True)
            d2h_done[i] = torch.cuda.Event()
            d2h_done[i].record(copy)

    def e2e_run(n):
        copy.wait_stream(torch.cuda.current_stream())   # the first H2D starts inside the timed region
        issue_h2d(0)
        for t in range(n):
            e2e_step(t)
        torch.cuda.current_stream().wait_stream(copy)   # ... and the last D2H ends inside it

    e2e_run(2)
    cx.barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_run(steps)
    s1.record()
    cx.barrier()
    hp.inputs = in_sets[0]
    return cx.max_over_ranks(s0.elapsed_time(s1)) / steps


def cp_parity_leg(cx, hp, args_w):
    """Multi-GPU correctness under the driver's eyes: ONE layer of the context-parallel path that was just timed vs
    the same layer through the single-GPU path (cp = 1, same kernels) on the all-gathered inputs, every rank checking
    its own rows of the self- and cross-attention outputs; max |diff| over all ranks. The fused paths and NCCL
    Ulysses keep every row's key order, so the bound is 0.0 (bit-identical); the NCCL ring re-associates the softmax
    over key blocks (one extra bf16 rounding per log-sum-exp merge): bound 2.5e-2."""
    torch, dist, w = cx.torch, cx.dist, hp.w
    mode = hp.mode()
    if w["slice_point"] != 0:
        return {"mode": mode, "skipped": "state with a cache prefix: the synthetic prefix differs per rank"}
    S = hp.S
    rows = hp.global_rows
    counts = torch.tensor([rows.numel()], device=cx.dev)
    all_counts = [torch.zeros_like(counts) for _ in range(cx.world)]
    dist.all_gather(all_counts, counts)
    nmax = max(int(c.item()) for c in all_counts)
    pad_rows = torch.full((nmax,), -1, device=cx.dev, dtype=rows.dtype)
    pad_rows[:rows.numel()] = rows
    all_rows = [torch.empty_like(pad_rows) for _ in range(cx.world)]
    dist.all_gather(all_rows, pad_rows)
    ref = HotPath(dict(w), cx.dev, cp=1, rank=0, layers=1)
    ref.ln, ref.lnx = hp.ln, hp.lnx
    for k, v in hp.inputs.items():
        if k == "text_kv":
            ref.inputs[k] = v
            continue
        buf = torch.zeros((nmax,) + tuple(v.shape[1:]), dtype=v.dtype, device=cx.dev)
        buf[:v.shape[0]] = v
        parts = [torch.empty_like(buf) for _ in range(cx.world)]
        dist.all_gather(parts, buf)
        full = torch.empty((S,) + tuple(v.shape[1:]), dtype=v.dtype, device=cx.dev)
        for r_, part in zip(all_rows, parts):
            keep = r_ >= 0
            full[r_[keep]] = part[keep]
        ref.inputs[k] = full
    ref.layer(0)
    hp.layer_fn()(0)
    cx.torch.cuda.synchronize()
    flat = lambda t_: t_.reshape(rows.numel(), -1).float()   # (rows, Hq, 128) or the scheduler's (rows, 1, Hq * 128)
    d_self = (flat(hp.core_out) - flat(ref.core_out[rows])).abs().max().item()
    d_cross = (flat(hp.x_out) - flat(ref.x_out[rows])).abs().max().item()
    finite = bool(torch.isfinite(hp.core_out.float()).all()) and bool(torch.isfinite(hp.x_out.float()).all())
    max_abs = cx.max_over_ranks(max(d_self, d_cross) if finite else float("inf"))
    bound = 2.5e-2 if mode == "ring-nccl" else 0.0
    del ref
    torch.cuda.empty_cache()
    return {"mode": mode, "max_abs": max_abs, "bound": bound, "ok": bool(max_abs <= bound), "rows": S,
            "what": "one layer (self- and cross-attention outputs) of the timed context-parallel path vs the "
                    "single-GPU path on the gathered inputs, max |diff| over all rows of all ranks"}


# ----------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--overlap-degree", type=int, default=1)
    ap.add_argument("--cp-mode", default="fused", choices=["fused", "nccl", "ring", "ring-nccl"],
                    help="N > 1: Ulysses with the all-to-alls fused into the kernels' stores over NVLink peer memory, "
                         "Ulysses with NCCL all_to_all_single, sequence-sharded ring-KV with the K/V all-gather fused "
                         "into the prologue's stores, or the NCCL send/recv ring")
    ap.add_argument("--ref-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-dit-step", action="store_true", help="skip the whole-DiT-block step latency leg")
    ap.add_argument("--no-fp8-dit-step", action="store_true",
                    help="N = 1: skip the fp8_quant variant of the DiT step leg (the *_distill_quant configs)")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the flash-attn (reference fallback) self-attention leg on the GPU")
    ap.add_argument("--no-extra-legs", action="store_true",
                    help="N > 1: skip the extra (not-the-value) legs: ring-KV on the same workload, and configs[3] "
                         "(24B / 720p) under Ulysses and ring-KV")
    ap.add_argument("--graph", default="off", choices=["on", "off"],
                    help="replay the hot-path step from one CUDA graph in the timed region (launch-bound workloads)")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: no e2e / CPU legs, warm-up exactly as given (numbers are not bench values)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.profile:
        args.no_e2e = args.no_cpu_baseline = args.no_dit_step = args.no_gpu_reference = args.no_extra_legs = True
    if args.impl == "reference":
        return run_reference(args, w)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG is left as the caller set it. NCCL carries only plumbing here (barriers, the max-over-ranks
        # all-reduce, rendezvous of the symmetric-memory buffers); the default data plane is the kernels' own NVLink
        # peer stores, so NCCL's channel log says nothing about the bytes that matter.
        dist.init_process_group("nccl", device_id=dev)
        from magi_b200 import parallel_state as mpu
        mpu.initialize_model_parallel(cp_size=world)
    cx = Ctx(torch, dist, dev, world, rank, local)

    hp = HotPath(w, dev, cp=world, rank=rank, fused=args.cp_mode == "fused", ring=args.cp_mode.startswith("ring"),
                 ring_fused=args.cp_mode == "ring")
    hp.ulysses_overlap_degree = args.overlap_degree
    self_f, cross_f = attn_flops(w)
    step_flops = (self_f + cross_f) * w["layers"]

    # ---- device-resident timing (value)
    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    t = timed_leg(cx, hp, args.steps, n_warm, graph=args.graph == "on")
    ms_step, ms_total, attn_ms = t["ms_step"], t["ms_total"], t["attn_ms"]
    value = step_flops / (ms_step * 1e-3) / 1e12
    attn_avg_ms = sum(attn_ms) / len(attn_ms)
    launches_per_layer = len(attn_ms) / (t["events_steps"] * hp.event_layers())   # self-attention launches per layer
    launch_flops = self_f / world / launches_per_layer
    achieved = launch_flops / (attn_avg_ms * 1e-3) / 1e12
    peaks = load_peaks()

    # ---- end to end: host buffers in, host buffers out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e_ms = e2e_leg(cx, hp, args.steps)
        e2e = {"value": round(step_flops / (e2e_ms * 1e-3) / 1e12, 2), "unit": "TFLOP/s",
               "ms_per_step": round(e2e_ms, 3), "h2d_bytes_per_step": hp.h2d_bytes, "d2h_bytes_per_step": hp.d2h_bytes,
               "api": "magi_b200.functional (ctypes -> libmagi_attn.so C-ABI), pinned host buffers, copies "
                      "double-buffered on a copy stream" + (" (per rank: its token shard in, its rows out)"
                                                            if world > 1 else "")}

    phases = phase_breakdown_leg(cx, hp) if not args.profile else None

    # ---- multi-GPU parity of the path that was just timed
    cp_parity = cp_parity_leg(cx, hp, w) if world > 1 and not args.profile else None

    dit = None
    if not args.no_dit_step and not args.cp_mode.startswith("ring"):
        dit = dit_step_leg(w, dev, world, rank, min(args.steps, 3), cx.max_over_ranks, cx.barrier)
        if world == 1 and not args.no_fp8_dit_step:
            try:
                dit["fp8_quant"] = dit_step_leg(w, dev, world, rank, min(args.steps, 3), cx.max_over_ranks, cx.barrier,
                                                fp8=True)
            except Exception as e:  # an extra leg must never cost the headline line
                dit["fp8_quant"] = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- extra legs at N > 1 (never the value): ring-KV on the same workload, configs[3] under both strategies
    extra = None
    if world > 1 and not args.no_extra_legs and args.cp_mode == "fused" and args.workload == "cfg2":
        extra = {}
        del hp
        torch.cuda.empty_cache()
        for name, wl, ring in (("cfg2_ring_kv", "cfg2", True), ("cfg4_ulysses", "cfg4", False), ("cfg4_ring_kv", "cfg4", True)):
            w2 = WORKLOADS[wl]
            try:
                hp2 = HotPath(w2, dev, cp=world, rank=rank, fused=not ring, ring=ring, ring_fused=ring)
                steps2 = max(2, min(args.steps, 3 if wl == "cfg4" else 5))
                t2 = timed_leg(cx, hp2, steps2, 3 if wl == "cfg2" else 1)
                f2 = sum(attn_flops(w2)) * w2["layers"]
                par = cp_parity_leg(cx, hp2, w2)
                extra[name] = {"workload": w2["desc"][:60], "parallelism": hp2.mode(),
                               "value": round(f2 / (t2["ms_step"] * 1e-3) / 1e12, 1), "unit": "TFLOP/s",
                               "ms_per_step": round(t2["ms_step"], 2), "steps": steps2,
                               "cp_parity_max_abs": par.get("max_abs"), "cp_parity_ok": par.get("ok")}
                del hp2
            except Exception as e:  # an extra leg must never cost the headline line
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
            torch.cuda.empty_cache()
        hp = None

    gpu_ref = None
    if world == 1 and not args.no_gpu_reference:
        # last GPU leg on purpose: it runs third-party library kernels, nothing measured above can be affected by it
        gpu_ref = gpu_reference_leg(w, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, cpu = cpu_baseline_sample(w, budget_s=args.ref_budget)

    if rank == 0:
        from magi_b200 import _lib
        traffic = NCU_TRAFFIC.get(args.workload) if world == 1 else None
        line = {
            "metric": "chunk_causal_attn_tflops", "value": round(value, 2), "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": round(ms_step, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": static_config(args, w, world),
            "clocks": t["clocks"],
            "e2e": e2e,
            "gpu_launches": t["launches"],
            "roofline": {"bound": "tensor", "kernel": f"magi::{_lib.attn_kernel_name()} (self-attention launch)",
                         "achieved": round(achieved, 1), "peak": peaks["sustained"], "unit": "TFLOP/s",
                         "frac": round(achieved / peaks["sustained"], 3),
                         "traffic": traffic[0] if traffic else None,
                         "traffic_source": traffic[1] if traffic else "not captured at N > 1 (ncu is single-GPU only)",
                         "peak_source": peaks["source"] + ", sustained (kernel timed inside a long step); burst "
                                        f"{peaks['burst']} -> frac {achieved / peaks['burst']:.3f}",
                         "flops_per_launch": launch_flops, "avg_launch_ms": round(attn_avg_ms, 4),
                         "share_of_step": round(attn_avg_ms * launches_per_layer * w["layers"] / ms_step, 3),
                         "launches_timed": len(attn_ms),
                         "step_launch_mode": "one CUDA graph replay per step; kernel times from 2 eager steps after "
                                             "the timed region" if t["graph"] else "eager launches, events in the "
                                             "timed region",
                         "pct_of_bf16_peak_burst": round(100 * value / world / peaks["burst"], 1),
                         "pct_of_bf16_peak_sustained": round(100 * value / world / peaks["sustained"], 1),
                         "pct_of_bf16_datasheet_2250": round(100 * value / world / 2250.0, 1)},
        }
        # measured companions ride inside `roofline` too: the driver's record keeps that object whole
        if phases is not None:
            line["roofline"]["phases_us_per_layer"] = phases
            if world == 1 and phases.get("prologue"):
                # the path's memory-bound kernel against the measured HBM copy peak. The phase time is an event-to-event
                # interval of the diagnostic steps (launch gap included), i.e. an upper bound of the kernel's duration
                pb = prologue_bytes(w)
                gbs = pb / (phases["prologue"] * 1e-6) / 1e9
                line["roofline"]["prologue_hbm"] = {
                    "bound": "hbm", "kernel": "magi::prologue_ring_kernel (one launch per layer: q / k / v + cache "
                                              "append + cross-attention qx / text kv)",
                    "algorithmic_bytes": pb, "us": phases["prologue"], "achieved": round(gbs, 1), "peak": peaks["hbm"],
                    "unit": "GB/s", "frac": round(gbs / peaks["hbm"], 3)}
        if cp_parity is not None:
            line["cp_parity"] = cp_parity
            line["roofline"]["cp_parity"] = {k: cp_parity.get(k) for k in ("mode", "max_abs", "bound", "ok", "skipped")
                                             if k in cp_parity}
        if extra is not None:
            line["extra_legs"] = extra
            line["roofline"]["extra_legs"] = extra
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if dit is not None:
            line["dit_denoise_step"] = dit
            line["roofline"]["dit_denoise_step_ms"] = dit["ms_per_step"]
            line["roofline"]["dit_denoise_step_tflops"] = dit["tflops_total"]
            if "ms_per_step" in dit.get("fp8_quant", {}):
                line["roofline"]["dit_denoise_step_fp8_quant_ms"] = dit["fp8_quant"]["ms_per_step"]
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
