"""Multi-GPU diagnostic (torchrun): bench.py's cp_parity leg on one layer of several workloads / modes, with the
self- and cross-attention differences reported per rank together with WHERE the differing rows sit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from magi_b200 import parallel_state as mpu
    mpu.initialize_model_parallel(cp_size=world)
    cx = bench.Ctx(torch, dist, dev, world, rank, local)
    cases = [a.split(":") for a in sys.argv[1:]] or [["cfg2", "fused"], ["cfg2", "ring"], ["cfg4", "fused"], ["cfg4", "ring"]]
    for wl, mode in cases:
        w = bench.WORKLOADS[wl]
        hp = bench.HotPath(w, dev, cp=world, rank=rank, fused=mode == "fused", ring=mode.startswith("ring"),
                           ring_fused=mode == "ring", layers=2)
        for rep in range(2):
            hp.step()
        cx.barrier()
        # the parity leg, inlined so that the pieces can be looked at
        res = bench.cp_parity_leg(cx, hp, w)
        # recompute the reference for row-level detail
        ref = bench.HotPath(dict(w), dev, cp=1, rank=0, layers=1)
        ref.ln, ref.lnx = hp.ln, hp.lnx
        rows = hp.global_rows
        S = hp.S
        for k, v in hp.inputs.items():
            if k == "text_kv":
                ref.inputs[k] = v
                continue
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([v.shape[0]], device=dev))
            nmax = max(int(s_.item()) for s_ in sizes)
            buf = torch.zeros((nmax,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
            buf[:v.shape[0]] = v
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf)
            rbuf = torch.full((nmax,), -1, device=dev, dtype=rows.dtype)
            rbuf[:rows.numel()] = rows
            allrows = [torch.empty_like(rbuf) for _ in range(world)]
            dist.all_gather(allrows, rbuf)
            full = torch.empty((S,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
            for r_, part in zip(allrows, parts):
                keep = r_ >= 0
                full[r_[keep]] = part[keep]
            ref.inputs[k] = full
        ref.layer(0)
        torch.cuda.synchronize()
        ref_core, ref_x = ref.core_out[rows].clone(), ref.x_out[rows].clone()
        ref.layer(0)
        torch.cuda.synchronize()
        print(f"[{wl}:{mode} rank {rank}] reference run twice: self diff "
              f"{(ref.core_out[rows].float() - ref_core.float()).abs().max().item():.3e}", flush=True)
        prev = None
        for rep, fence in enumerate((False, False, True, True)):
            if fence:
                torch.cuda.synchronize()
                dist.barrier()
            hp.layer_fn()(rep % 2)   # alternate the layer (= cache slot): ring-KV has one barrier per layer
            torch.cuda.synchronize()
            mine = hp.core_out.reshape(rows.numel(), -1).float()
            d = (mine - ref_core.reshape(rows.numel(), -1).float()).abs()
            bad_rows = (d.amax(dim=1) > 0).nonzero().flatten()
            msg = (f"[{wl}:{mode} rank {rank}] CP call {rep} ({'sync+barrier before' if fence else 'back to back'}): self "
                   f"max_abs {d.max().item():.3e}, {bad_rows.numel()} of {d.shape[0]} rows differ")
            if bad_rows.numel():
                hd = (d.reshape(rows.numel(), -1, 128).amax(dim=(0, 2)) > 0).nonzero().flatten().tolist()
                g = rows[bad_rows]
                msg += f"; global rows {g[:3].tolist()}..{g[-2:].tolist()} heads {hd}"
            if prev is not None:
                msg += f" | vs previous CP call: {(mine - prev).abs().max().item():.3e}"
            prev = mine.clone()
            dx = (hp.x_out.reshape(rows.numel(), -1).float() - ref_x.reshape(rows.numel(), -1).float()).abs().max().item()
            print(msg + f" | cross {dx:.3e}", flush=True)
        print(f"[{wl}:{mode} rank {rank}] cp_parity_leg: {res.get('max_abs')}", flush=True)
        del hp, ref
        torch.cuda.empty_cache()
        cx.barrier()
    dist.destroy_process_group()
