"""TEST DOUBLE of magi_b200.functional on the CPU, built on the oracle. It exists so that the HOST LOGIC of the drop-in
module under context parallelism (FullyParallelAttention._forward_ulysses / _forward_cso / _forward_ring: cache plans,
send layouts, all-to-all plumbing, head groups, range handling) can run in the CPU suite on gloo, world size > 1,
against the reference's per-rank golden outputs. It mirrors the argument meaning of the real functions; the kernels'
own numerics are what tests/test_gpu_parity.py checks on the B200. Never imported by the product.

Only what the non-fused paths call is implemented (no NVLink peer / multicast routing: gloo has no symmetric memory,
and the module falls back to its NCCL-style paths by itself)."""
import numpy as np
import torch

import magi_oracle as O

HEAD_DIM = 128


def _np(t):
    return t.detach().float().numpy()


def _bf(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16)


def _norm_rope(x, heads, ln, sin, cos, batch, eps, zcg):
    rows = x.shape[0]
    v = _np(x).reshape(rows, heads, HEAD_DIM)
    if ln is not None:
        w, b = ln
        if w.dtype == torch.bfloat16:
            v = O.xattn_layer_norm(v, _np(w), _np(b), eps, zcg)
        else:
            v = O.layer_norm(v, _np(w), _np(b), eps, zcg)
    if sin is not None:
        s, c = np.repeat(_np(sin), batch, axis=0), np.repeat(_np(cos), batch, axis=0)   # row r -> table row r // batch
        v = O.apply_rotary_neox(v, c, s)
    return O.bf16_round(v)


def _send_layout(v, cp, rep=1):
    """"seq (cp hn) hd -> (cp seq) hn hd" with KV-head replication (context_parallel.py:393-396)."""
    if rep > 1:
        v = np.repeat(v, rep, axis=1)
    rows, heads, d = v.shape
    return v.reshape(rows, cp, heads // cp, d).transpose(1, 0, 2, 3).reshape(cp * rows, heads // cp, d)


def qk_norm_rope(x_q, x_k, x_v, num_heads_q, num_heads_kv, ln_q, ln_k, sin, cos, batch=1, eps=1e-6,
                 zero_centered_gamma=True, q_out=None, kv_out=None, cache=None, cache_row_offset=0, cache_rows=0,
                 head_perm_cp=1, cross=None, **unsupported):
    assert not {k for k, v in unsupported.items() if v is not None and v != 0}, "peer / gather routing has no CPU double"
    if cross:   # the cross-attention's tensors riding in the same launch: always plain layout, bf16 LN, no RoPE
        if cross.get("x_qx") is not None:
            cross["qx"], _ = qk_norm_rope(cross["x_qx"], None, None, num_heads_q, num_heads_kv, cross["ln_qx"], None,
                                          None, None, eps=eps, zero_centered_gamma=zero_centered_gamma,
                                          q_out=cross.get("qx_out"))
        if cross.get("text_kv") is not None:
            t = cross["text_kv"]
            _, cross["tkv"] = qk_norm_rope(None, t[..., :HEAD_DIM], t[..., HEAD_DIM:], num_heads_q, num_heads_kv, None,
                                           cross["ln_tk"], None, None, eps=eps,
                                           zero_centered_gamma=zero_centered_gamma, kv_out=cross.get("tkv_out"))
        if x_q is None and x_k is None and x_v is None:
            return None, None
    cp = max(1, head_perm_cp)
    q_res = kv_res = None
    if x_q is not None:
        q = _norm_rope(x_q, num_heads_q, ln_q, sin, cos, batch, eps, zero_centered_gamma)
        if cp > 1:
            q = _send_layout(q, cp)
        q_res = _bf(q)
        if q_out is not None:
            q_out.copy_(q_res)
            q_res = q_out
    if x_k is not None or x_v is not None:
        rows = (x_k if x_k is not None else x_v).shape[0]
        k = _norm_rope(x_k, num_heads_kv, ln_k, sin, cos, batch, eps, zero_centered_gamma)
        v = _np(x_v).reshape(rows, num_heads_kv, HEAD_DIM)
        kv = np.concatenate([k, v], axis=-1)
        if cp > 1:
            kv = _send_layout(kv, cp, cp // num_heads_kv if cp > num_heads_kv else 1)
        kv_res = _bf(kv)
        if kv_out is not None:
            kv_out.copy_(kv_res)
            kv_res = kv_out
        if cache is not None and cache_rows > 0:
            assert cp == 1
            cache[cache_row_offset:cache_row_offset + cache_rows] = kv_res[:cache_rows]
    return q_res, kv_res


def attn_ranges(q, k_new, v_new, q_ranges=None, k_ranges=None, k_cache=None, v_cache=None, cu_seqlens_q=None,
                cu_seqlens_k=None, softmax_scale=None, out=None, return_lse=True, zero_fill=False, num_sms=0,
                proj_input=None, peer_out=None, kernel=0, plan=None, independent_of_prev=False):
    assert peer_out is None, "peer routing has no CPU double"
    if independent_of_prev:   # the real function's preconditions (functional.attn_ranges)
        assert plan is not None and not return_lse and (out is not None or proj_input is not None)
    if plan is not None:   # the real function checks the plan against the call; so does the double
        tk = sum(t.shape[0] for t in (k_cache, k_new) if t is not None)
        assert plan == ("plan", q_ranges.data_ptr(), k_ranges.data_ptr(), q.shape[1], q.shape[0], tk), "stale plan"
    ks = [t for t in (k_cache, k_new) if t is not None and t.shape[0] > 0]
    vs = [t for t in (v_cache, v_new) if t is not None and t.shape[0] > 0]
    if q_ranges is None:
        cq, ck = cu_seqlens_q.numpy().astype(np.int64), cu_seqlens_k.numpy().astype(np.int64)
        qr, kr = np.stack([cq[:-1], cq[1:]], 1), np.stack([ck[:-1], ck[1:]], 1)
    else:
        qr, kr = q_ranges.numpy().astype(np.int64).reshape(-1, 2), k_ranges.numpy().astype(np.int64).reshape(-1, 2)
    hkv = (k_new if k_new is not None else k_cache).shape[1]
    k = np.concatenate([_np(t) for t in ks], 0) if ks else np.zeros((0, hkv, HEAD_DIM), np.float32)
    v = np.concatenate([_np(t) for t in vs], 0) if vs else np.zeros((0, hkv, HEAD_DIM), np.float32)
    kr = np.clip(kr, 0, k.shape[0])
    o, lse = O.attn_ranges(_np(q), k, v, qr, kr, softmax_scale)
    covered = np.zeros(q.shape[0], dtype=bool)
    for a, b in qr:
        covered[a:b] = True
    rows = torch.from_numpy(np.nonzero(covered)[0])
    res = _bf(o)
    if proj_input is not None:
        buf, n_src, hn = proj_input
        tq, hq = q.shape[0], q.shape[1]
        view = buf.view(tq, hn, 2, hq // hn, HEAD_DIM)
        view[rows, :, n_src] = res.view(tq, hn, hq // hn, HEAD_DIM)[rows]
        out = buf
    elif out is None:
        out = res                                  # uncovered rows of a fresh tensor are unspecified: zeros here
    else:
        out[rows] = res[rows]                      # rows covered by no range are left untouched
    return out, (torch.from_numpy(lse) if return_lse else None)


def stack_qkv_weights(wq, wk, wv, wqx=None):
    ws = [t for t in (wq, wk, wv, wqx) if t is not None]
    if ws[0].dtype == torch.float8_e4m3fn:
        ws = [t.view(torch.uint8) for t in ws]
    return torch.cat(ws, dim=0).contiguous()


def fp8_quantize(x, scale):
    """div_clamp_to through the oracle, as a float8_e4m3fn tensor."""
    b = O.e4m3_encode(O.div_clamp_to(_np(x), _np(scale)))
    return torch.from_numpy(np.ascontiguousarray(b)).view(torch.float8_e4m3fn).reshape(x.shape)


def fp8_scaled_linear(xq, weight, input_scale, weight_scale, out_dtype=torch.bfloat16):
    """bmm_fp8's published math on the already quantised input (see magi_oracle.linear_fp8)."""
    acc = O.e4m3_decode(xq.view(torch.uint8).numpy()).reshape(-1, xq.shape[-1]) @ \
        O.e4m3_decode(weight.view(torch.uint8).numpy()).T
    s = np.float32(_np(input_scale).reshape(-1)[0]) * np.float32(_np(weight_scale).reshape(-1)[0])
    return _bf(O.bf16_round(acc * s)).to(out_dtype).view(*xq.shape[:-1], weight.shape[0])


def qkv_proj_norm_rope(x, w, num_heads_q, num_heads_kv, ln_q, ln_k, sin, cos, ln_qx=None, batch=1, eps=1e-6,
                       zero_centered_gamma=True, cache=None, cache_row_offset=0, cache_rows=0, num_sms=0,
                       peer_q_out=None, peer_kv_out=None, peer_row_offset=0, out_scale=None):
    """The fused projection GEMM as the composition it replaces: nn.Linear in bf16 (or, with out_scale, the fp8 GEMM
    scaled per projection), then the prologue double."""
    nq, nk = num_heads_q * HEAD_DIM, num_heads_kv * HEAD_DIM
    if out_scale is not None:
        assert x.dtype in (torch.float8_e4m3fn, torch.uint8) and w.dtype in (torch.float8_e4m3fn, torch.uint8)
        acc = O.e4m3_decode(x.view(torch.uint8).numpy()) @ O.e4m3_decode(w.view(torch.uint8).numpy()).T
        col_scale = np.repeat(_np(out_scale), [nq, nk, nk, nq])[:acc.shape[1]]
        y = _bf(O.bf16_round(acc * col_scale))
    else:
        y = _bf(O.linear_bf16(_np(x), _np(w)))
    if peer_q_out is not None:    # the Ulysses input all-to-all fused into the epilogue: the prologue double's routing
        cp = len(peer_q_out)
        hkv_loc = num_heads_kv // cp if num_heads_kv >= cp else 1
        q, kv = qk_norm_rope(y[:, :nq], y[:, nq:nq + nk], y[:, nq + nk:nq + 2 * nk], num_heads_q, num_heads_kv, ln_q,
                             ln_k, sin, cos, batch=batch, eps=eps, zero_centered_gamma=zero_centered_gamma,
                             head_perm_cp=cp, peer_q_out=peer_q_out, peer_kv_out=peer_kv_out,
                             peer_row_offset=peer_row_offset, peer_heads=(num_heads_q // cp, hkv_loc))
    else:
        q, kv = qk_norm_rope(y[:, :nq], y[:, nq:nq + nk], y[:, nq + nk:nq + 2 * nk], num_heads_q, num_heads_kv, ln_q,
                             ln_k, sin, cos, batch=batch, eps=eps, zero_centered_gamma=zero_centered_gamma, cache=cache,
                             cache_row_offset=cache_row_offset, cache_rows=cache_rows)
    qx = None
    if ln_qx is not None:
        qx, _ = qk_norm_rope(y[:, nq + 2 * nk:], None, None, num_heads_q, num_heads_kv, ln_qx, None, None, None,
                             batch=batch, eps=eps, zero_centered_gamma=zero_centered_gamma)
    return q, kv, qx


def attn_plan(q_ranges=None, k_ranges=None, *, cu_seqlens_q=None, cu_seqlens_k=None, num_heads_q, total_q, total_k):
    return ("plan", q_ranges.data_ptr(), k_ranges.data_ptr(), num_heads_q, total_q, total_k)


def kv_cache_append(cache, kv, row_offset, rows=None):
    rows = kv.shape[0] if rows is None else rows
    cache[row_offset:row_offset + rows] = kv[:rows]


def cso_unshuffle(recv, cp, dn, clip, out=None):
    """dst[d * clip + t] = src[((t // spb) * dn + d) * spb + t % spb] (dit_module.py:1215-1224)."""
    spb = recv.shape[0] // (cp * dn)
    t = np.arange(clip)
    idx = np.concatenate([((t // spb) * dn + d) * spb + t % spb for d in range(dn)])
    res = recv[torch.from_numpy(idx)]
    if out is not None:
        out.copy_(res)
        return out
    return res.contiguous()


def ulysses_relayout(src, cp, mode, kv_repeat=1, out=None, dst_heads_total=0, dst_head_offset=0):
    t_, h_, e_ = src.shape
    if mode == 0:
        res = _bf(_send_layout(_np(src), cp, kv_repeat))
        if out is not None:
            out.copy_(res)
            return out
        return res
    seq = t_ // cp
    total = dst_heads_total if dst_heads_total > 0 else h_
    if out is None:
        out = torch.zeros((seq, cp * total, e_), dtype=src.dtype)
    out.view(seq, cp, total, e_)[:, :, dst_head_offset:dst_head_offset + h_] = src.view(cp, seq, h_, e_).permute(1, 0, 2, 3)
    return out


def lse_merge(out_a, lse_a, out_b, lse_b):
    m, l = O.lse_merge(_np(out_a), lse_a.numpy(), _np(out_b), lse_b.numpy())
    out_a.copy_(_bf(m))
    lse_a.copy_(torch.from_numpy(l))


def bias_modulate_add(x, residual, condition_map, gate, post_norm, out=None):
    if isinstance(post_norm, (tuple, list)):
        w, b, eps, zcg = post_norm
    else:
        w, b, eps, zcg = post_norm.weight, post_norm.bias, post_norm.eps, post_norm.zero_centered_gamma
    res = _bf(O.bias_modulate_add(_np(x), _np(residual), condition_map.numpy(), _np(gate), _np(w), _np(b), eps, zcg))
    if out is not None:
        out.copy_(res)
        return out
    return res


# ----------------------------------------------------------------------------------------------- peer-memory emulation
# The fused context-parallel paths hand the kernels raw device addresses of OTHER ranks' buffers (NVLink peer memory).
# For the CPU emulation (tests/test_cp_fused_emulated.py: virtual ranks = threads of one process) a "pointer" is
# (registry id << 40) + byte offset, so the byte arithmetic the module performs on such pointers keeps working.
_PEER_TENSORS = []


def mlp_activation(x, gated, out=None):
    r = _bf(O.mlp_activation(_np(x), gated))
    if out is not None:
        out.copy_(r)
        return out
    return r


def fake_ptr(t: torch.Tensor) -> int:
    assert t.is_contiguous() and t.dtype == torch.bfloat16
    _PEER_TENSORS.append(t.view(-1))
    return len(_PEER_TENSORS) << 40


def _peer(ptr: int, shape):
    """View of the registered buffer behind `ptr` from its byte offset on, shaped (-1, *shape[1:])."""
    flat = _PEER_TENSORS[(ptr >> 40) - 1]
    off = (ptr & ((1 << 40) - 1)) // 2
    row = int(np.prod(shape[1:]))
    n = (flat.numel() - off) // row
    return flat[off:off + n * row].view((n,) + tuple(shape[1:]))


_plain_qk_norm_rope, _plain_attn_ranges = qk_norm_rope, attn_ranges


def qk_norm_rope(x_q, x_k, x_v, num_heads_q, num_heads_kv, ln_q, ln_k, sin, cos, *, peer_q_out=None, peer_kv_out=None,
                 peer_row_offset=0, peer_heads=None, gather_kv_out=None, gather_geometry=None, cross=None, **kw):  # noqa: F811
    """Adds the routing of include/magi_attn.h's peer_* / gather_* fields on top of the plain double."""
    if peer_q_out is None and peer_kv_out is None and gather_kv_out is None:
        return _plain_qk_norm_rope(x_q, x_k, x_v, num_heads_q, num_heads_kv, ln_q, ln_k, sin, cos, cross=cross, **kw)
    if cross:   # the cross-attention's tensors are local whatever the routing of q / k / v
        _plain_qk_norm_rope(None, None, None, num_heads_q, num_heads_kv, None, None, None, None, cross=cross,
                            eps=kw.get("eps", 1e-6), zero_centered_gamma=kw.get("zero_centered_gamma", True))
    if gather_kv_out is not None:
        # cp_ring_kv: q stays local; K/V row (chunk j, stripe row t) -> row j * chunk_stride + row_offset + t everywhere
        q, kv = _plain_qk_norm_rope(x_q, x_k, x_v, num_heads_q, num_heads_kv, ln_q, ln_k, sin, cos, **kw)
        rpc, cstride, roff = gather_geometry
        r = torch.arange(kv.shape[0])
        dst = (r // rpc) * cstride + roff + r % rpc
        for ptr in gather_kv_out:
            _peer(ptr, (0, num_heads_kv, 2 * HEAD_DIM))[dst] = kv
        return q, None
    cp = kw["head_perm_cp"]
    q, kv = _plain_qk_norm_rope(x_q, x_k, x_v, num_heads_q, num_heads_kv, ln_q, ln_k, sin, cos, **kw)  # send layout
    rows = (x_q if x_q is not None else x_k).shape[0]
    for c in range(cp):     # head slice c -> rank c's receive buffer, at this rank's row offset
        if q is not None:
            _peer(peer_q_out[c], (0, peer_heads[0], HEAD_DIM))[peer_row_offset:peer_row_offset + rows] = \
                q[c * rows:(c + 1) * rows]
        if kv is not None:
            _peer(peer_kv_out[c], (0, peer_heads[1], 2 * HEAD_DIM))[peer_row_offset:peer_row_offset + rows] = \
                kv[c * rows:(c + 1) * rows]
    return None, None


def attn_ranges(q, k_new, v_new, q_ranges=None, k_ranges=None, *, peer_out=None, peer_proj=None, **kw):  # noqa: F811
    if peer_out is None:
        return _plain_attn_ranges(q, k_new, v_new, q_ranges, k_ranges, **kw)
    ptrs, bounds, head_off, heads_total = peer_out[:4]
    tstride = peer_out[4] if len(peer_out) > 4 else heads_total * HEAD_DIM      # batched rows: b * heads_total * 128
    kw.pop("out", None)
    kw.pop("independent_of_prev", None)
    res, _ = _plain_attn_ranges(q, k_new, v_new, q_ranges, k_ranges, **dict(kw, return_lse=False))
    covered = np.zeros(q.shape[0], dtype=bool)
    for a, b in q_ranges.numpy().reshape(-1, 2):
        covered[a:b] = True
    hq = q.shape[1]
    for rho, ptr in enumerate(ptrs):    # row g of local head h -> rank rho's output, row g - bounds[rho], head slot
        lo, hi = int(bounds[rho]), int(bounds[rho + 1])
        rows = torch.from_numpy(np.nonzero(covered[lo:hi])[0])
        if not len(rows):
            continue
        if peer_proj is None:
            # flat element indexing: with batched rows the pointer is advanced to the batch row, so the buffer does not
            # hold a whole number of (tstride)-wide rows from there on
            flat = _PEER_TENSORS[(ptr >> 40) - 1]
            off = (ptr & ((1 << 40) - 1)) // 2
            idx = off + rows[:, None] * tstride + head_off * HEAD_DIM + torch.arange(hq * HEAD_DIM)[None, :]
            flat[idx] = res[lo:hi][rows].reshape(len(rows), hq * HEAD_DIM)
        else:    # the owner's out-projection input "(hn n hd)": destination head H -> group H // gh, slot n, H % gh
            n_src, hn = peer_proj
            gh = heads_total // hn
            dst = _peer(ptr, (0, hn, 2, gh, HEAD_DIM))
            for h in range(hq):
                H = head_off + h
                dst[rows, H // gh, n_src, H % gh] = res[lo:hi][rows, h]
    return None, None
