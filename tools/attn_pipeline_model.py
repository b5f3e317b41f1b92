"""Discrete-event model of attn_fwd_kernel's steady state on ONE SM sub-partition, from a SASS listing - no GPU needed.

What is modelled:
  * the tensor pipe, in order:  S0(i) . PV1(i-1)[half 0, half 1] . S1(i) . PV0(i)[half 0, half 1];  an MMA group starts
    when the pipe is free AND its operand is there (S: always; a P half: `arrive_to_mma` cycles after the softmax warp's
    mbarrier.arrive for that half);
  * the two softmax warps that share an SMSP (one of each warpgroup): each waits for its S tile, then runs the
    steady-state instruction stream of the listing (tools/sass_issue_model.py locates it: first half with the running
    max, the vote, publish / second half / publish), instruction by instruction, with ptxas's stall counts and
    scoreboard waits. The warps compete for the issue slot (one instruction per cycle, higher warp id first) and for
    the fma / alu / MUFU pipes (one warp instruction per 2 / 2 / 8 cycles).
Everything else (TMA, the issuer's own instruction stream, the other three SMSPs - which run the same code) is folded
into four latencies taken from the clock64 trace of profiles/r01_trace_halfsplit.log.

    python tools/attn_pipeline_model.py /tmp/attn.sass [/tmp/other.sass ...]
prints the modelled cycles per 128-key step (2 query tiles) and the tensor-pipe utilisation for each listing.

CAVEAT - an exploration aid, not a predictor. Fitted against the four measured MAGI_EMU_MOD builds
(profiles/r01_variants_halfsplit.log: share 0 / 1-in-2 / 1-in-3 / 1-in-4 = 3.322 / 3.326 / 3.176 / 3.224 ms) the best
parameters are 4 cycles per packed f32x2 instruction and 10 per MUFU (ratios 1.08 / 1.10 / 1 / 0.99 against the measured
1.046 / 1.047 / 1 / 1.015), but then the absolute period is 4230 cycles (51 % tensor pipe; ncu measured 68 %), while the
textbook rates (2 and 8) give the right absolute level (2859 cycles, 76 %) and the wrong order of the variants. The two
parameter sets also disagree on the compiled-out candidates (static barriers: -13 % with the textbook rates, +3 % with
the fitted ones). Use it to see WHERE the chain spends its time; measure on a B200 to decide."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_issue_model as M  # noqa: E402


def tile_stream(path):
    """Steady-state per-tile instruction stream of one softmax warp: list of (pipe, stall, wait_mask, wbar, rbar, tag);
    tag is 'arrive0' / 'arrive1' on the two publishing arrives."""
    ins = M.parse(path)
    texts = [t for _, t, _ in ins]
    addr = [a for a, _, _ in ins]
    v = next(i for i, t in enumerate(texts) if t.startswith("VOTE.ANY P6"))
    b0 = max(i for i in range(v) if "BRA" in texts[i] and "BRA.DIV" not in texts[i]) + 1
    g = next(i for i in range(v, len(texts)) if texts[i].startswith("@!P6 BRA"))
    z = int(texts[g].split()[-1], 16)
    zi = addr.index(z)
    arrives = [i for i, t in enumerate(texts) if "SYNCS.ARRIVE.TRANS64.A1T0" in t and addr[i] > z]
    back = next(i for i in range(arrives[1], len(texts)) if texts[i].startswith("@P0 BRA"))
    stream = []
    for lo, hi in ((b0, g + 1), (zi, back + 1)):
        for i in range(lo, hi):
            c, op = M.ctrl(ins[i][2]), M.opname(texts[i])
            tag = "arrive0" if i == arrives[0] else "arrive1" if i == arrives[1] else ""
            stream.append((M.pipe(op), max(c["stall"], 1), c["wait"], c["wbar"], c["rbar"], tag, op))
    return stream


def simulate(stream, steps=40, mma_s=542, mma_pv_half=271, wake=60, ld_head=204, arrive_to_mma=150, lat_mufu=22,
             lat_other=20, rt_packed=2, rt_xu=8):
    """Returns (cycles per step in steady state, tensor-pipe busy fraction)."""
    RT = M.PIPE_RT
    # per warp state
    class W:
        pass
    warps = []
    for t in (0, 1):
        w = W()
        w.t, w.pc, w.ready, w.sb, w.tile, w.active = t, 0, 0, [0] * 6, -1, False
        warps.append(w)
    s_done = [dict(), dict()]          # S_t(i) completion time
    p_ready = [dict(), dict()]         # (i, half) -> time the issuer can use it
    pipe_busy = {"fma": 0, "alu": 0, "xu": 0, "mem": 0}
    # tensor pipe program: list of ops in issue order
    ops = []
    for i in range(steps):
        ops.append(("S", 0, i))
        if i > 0:
            ops += [("PV", 1, i - 1, 0), ("PV", 1, i - 1, 1)]
        ops.append(("S", 1, i))
        ops += [("PV", 0, i, 0), ("PV", 0, i, 1)]
    op_i, tensor_free, tensor_busy_total = 0, 0, 0
    now, step_times = 0, []
    horizon = steps * 6000
    while op_i < len(ops) and now < horizon:
        # --- tensor pipe: start the next op if possible
        while op_i < len(ops) and tensor_free <= now:
            op = ops[op_i]
            if op[0] == "S":
                dur = mma_s
                tensor_busy_total += dur
                tensor_free = now + dur
                s_done[op[1]][op[2]] = tensor_free
                if op[1] == 0:
                    step_times.append(now)
                op_i += 1
            else:
                _, t, i, h = op
                rdy = p_ready[t].get((i, h))
                if rdy is None or rdy > now:
                    break
                tensor_busy_total += mma_pv_half
                tensor_free = now + mma_pv_half
                op_i += 1
        # --- softmax warps: wake up on their S tile
        for w in warps:
            if not w.active:
                nxt = w.tile + 1
                d = s_done[w.t].get(nxt)
                if d is not None and now >= d + wake + ld_head:
                    w.active, w.tile, w.pc, w.ready, w.sb = True, nxt, 0, now, [0] * 6
        # --- issue at most one instruction this cycle, higher warp id (warpgroup 1) first
        for w in (warps[1], warps[0]):
            if not w.active or w.ready > now:
                continue
            pipe, stall, wait, wbar, rbar, tag, op = stream[w.pc]
            if any(((wait >> s) & 1) and w.sb[s] > now for s in range(6)):
                continue
            if pipe_busy[pipe] > now:
                continue
            base = op.split(".")[0]
            rt = rt_packed if base in ("FFMA2", "FADD2", "FMUL2") else rt_xu if pipe == "xu" else RT[pipe]
            pipe_busy[pipe] = now + rt
            lat = lat_mufu if pipe == "xu" else lat_other
            if wbar < 6:
                w.sb[wbar] = max(w.sb[wbar], now + lat)
            if rbar < 6:
                w.sb[rbar] = max(w.sb[rbar], now + 6)
            if tag:
                p_ready[w.t][(w.tile, 0 if tag == "arrive0" else 1)] = now + arrive_to_mma
            w.ready = now + stall
            w.pc += 1
            if w.pc == len(stream):
                w.active = False
            break
        now += 1
    n = len(step_times)
    lo, hi = n // 3, n - 2
    period = (step_times[hi] - step_times[lo]) / (hi - lo)
    return period, (2 * mma_s + 4 * mma_pv_half) / period


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("sass", nargs="+")
    ap.add_argument("--arrive-to-mma", type=int, default=150)
    ap.add_argument("--wake", type=int, default=60)
    ap.add_argument("--ld-head", type=int, default=204, help="cycles from wake-up to the first softmax instruction "
                    "(address arithmetic + tcgen05.ld + wait::ld)")
    ap.add_argument("--rt-packed", type=int, default=2, help="cycles per packed f32x2 instruction on the fma pipe")
    ap.add_argument("--rt-xu", type=int, default=8, help="cycles per MUFU instruction")
    a = ap.parse_args()
    for path in a.sass:
        st = tile_stream(path)
        period, util = simulate(st, wake=a.wake, ld_head=a.ld_head, arrive_to_mma=a.arrive_to_mma, rt_packed=a.rt_packed,
                                rt_xu=a.rt_xu)
        print(f"{os.path.basename(path):24s} {len(st):4d} instructions per tile and warp -> {period:7.0f} cycles per "
              f"128-key step, tensor pipe {100 * util:5.1f} %")
