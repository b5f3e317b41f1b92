"""CPU-only checks of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/magi_attn.h declares, mirrors the header's struct layouts, and fails loudly (status + message, no crash)
on bad arguments. No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import magi_b200  # noqa: F401
from magi_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "magi_attn.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MAGI_API\s+[\w\s\*]+?\b(magi_\w+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} is declared in include/magi_attn.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared_symbols()) <= exported
    # nothing but the C-ABI leaks out of the library
    assert all(s.startswith("magi_") for s in exported), exported


def test_struct_layout_matches_header():
    """A C translation unit compiled against the header must agree with the ctypes mirrors on size and offsets."""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "magi_attn.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(magi_view), sizeof(magi_attn_args), offsetof(magi_attn_args, out),
             offsetof(magi_attn_args, n_ranges), offsetof(magi_attn_args, workspace), offsetof(magi_attn_args, num_sms));
      printf("%zu %zu %zu %zu %zu\n", sizeof(magi_prologue_args), offsetof(magi_prologue_args, num_heads_q),
             offsetof(magi_prologue_args, ln_dtype), offsetof(magi_prologue_args, q_out),
             offsetof(magi_prologue_args, head_perm_cp));
      printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(magi_attn_args, peer_out), offsetof(magi_attn_args, peer_row_bounds),
             offsetof(magi_prologue_args, peer_kv_out), offsetof(magi_prologue_args, peer_row_offset),
             sizeof(MagiGatingArgs), offsetof(MagiGatingArgs, ln_w), offsetof(MagiGatingArgs, hidden));
      printf("%zu %zu %zu\n", offsetof(magi_prologue_args, gather_peers), offsetof(magi_prologue_args, gather_kv_out),
             offsetof(magi_prologue_args, gather_row_offset));
      printf("%zu %zu %zu %zu %zu\n", sizeof(magi_rotary_args), offsetof(magi_rotary_args, batch),
             offsetof(magi_rotary_args, heads), offsetof(magi_rotary_args, x_batch_stride),
             offsetof(magi_rotary_args, cos));
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(magi_qkv_proj_args), offsetof(magi_qkv_proj_args, w),
             offsetof(magi_qkv_proj_args, num_sms), offsetof(magi_qkv_proj_args, peer_row_offset),
             offsetof(magi_qkv_proj_args, fp8), offsetof(magi_qkv_proj_args, out_scale));
      return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        a, b, c_, d_, e_, f_ = subprocess.run([exe], capture_output=True, text=True,
                                              check=True).stdout.strip().splitlines()
    Q = _lib.MagiQkvProjArgs
    assert [int(x) for x in f_.split()] == [C.sizeof(Q), Q.w.offset, Q.num_sms.offset, Q.peer_row_offset.offset,
                                            Q.fp8.offset, Q.out_scale.offset]
    A, P, G, R = _lib.MagiAttnArgs, _lib.MagiPrologueArgs, _lib.MagiGatingArgs, _lib.MagiRotaryArgs
    assert [int(x) for x in e_.split()] == [C.sizeof(R), R.batch.offset, R.heads.offset, R.x_batch_stride.offset,
                                            R.cos.offset]
    assert [int(x) for x in d_.split()] == [P.gather_peers.offset, P.gather_kv_out.offset, P.gather_row_offset.offset]
    assert [int(x) for x in c_.split()] == [A.peer_out.offset, A.peer_row_bounds.offset, P.peer_kv_out.offset,
                                            P.peer_row_offset.offset, C.sizeof(G), G.ln_w.offset, G.hidden.offset]
    assert [int(x) for x in a.split()] == [C.sizeof(_lib.MagiView), C.sizeof(A), A.out.offset, A.n_ranges.offset,
                                           A.workspace.offset, A.num_sms.offset]
    assert [int(x) for x in b.split()] == [C.sizeof(P), P.num_heads_q.offset, P.ln_dtype.offset, P.q_out.offset,
                                           P.head_perm_cp.offset]


def test_argument_errors_are_reported_not_crashed():
    """Precondition failures come back as a status code with a message (no exception crosses the ABI); these checks
    happen before any CUDA call, so they run without a GPU."""
    L = _lib.lib()
    a = _lib.MagiAttnArgs()
    a.head_dim = 64
    assert L.magi_attn_fwd(C.byref(a), None) != 0
    assert b"head_dim must be 128" in L.magi_last_error()
    a.head_dim = 128
    a.num_heads_q, a.num_heads_kv = 6, 4
    assert L.magi_attn_fwd(C.byref(a), None) != 0
    assert b"multiple of num_heads_kv" in L.magi_last_error()
    a.num_heads_kv = 2
    assert L.magi_attn_fwd(C.byref(a), None) != 0      # neither ranges nor cu_seqlens
    assert b"exactly one of" in L.magi_last_error()
    assert L.magi_attn_fwd(None, None) != 0
    p = _lib.MagiPrologueArgs()
    assert L.magi_qk_norm_rope_fwd(C.byref(p), None) != 0
    assert b"no input tensor" in L.magi_last_error()
    assert L.magi_kv_cache_append(None, 256, None, 256, 0, 4, 12, None) != 0
    assert L.magi_ulysses_relayout(None, None, 4, 3, 128, 2, 1, 0, 0, 0, None) != 0   # 3 heads not divisible by cp=2
    r = _lib.MagiRotaryArgs()
    r.dtype, r.batch, r.seqlen, r.heads, r.head_dim, r.rotary_dim = 0, 1, 4, 2, 128, 100
    assert L.magi_apply_rotary_emb(C.byref(r), None) != 0
    assert b"multiples of 4" in L.magi_last_error()
    r.rotary_dim = 160
    assert L.magi_apply_rotary_emb(C.byref(r), None) != 0 and b"<= head_dim" in L.magi_last_error()
    r.rotary_dim = 96
    assert L.magi_apply_rotary_emb(C.byref(r), None) != 0 and b"NULL pointer" in L.magi_last_error()
    assert L.magi_fp8_quantize(None, 64, None, 1, None, 64, 4, 60, None) != 0 and b"multiple of 16" in L.magi_last_error()
    assert L.magi_fp8_quantize(None, 64, None, 7, None, 64, 4, 64, None) != 0 and b"scale_len" in L.magi_last_error()
    assert L.magi_fp8_quantize(None, 64, None, 1, None, 64, 4, 64, None) != 0 and b"NULL" in L.magi_last_error()
    assert L.magi_fp8_quantize(None, 64, None, 1, None, 64, 0, 64, None) == 0          # nothing to do
    qp = _lib.MagiQkvProjArgs()
    qp.rows, qp.hidden, qp.num_heads_q, qp.num_heads_kv, qp.batch, qp.fp8 = 4, 192, 2, 2, 1, 1
    assert L.magi_qkv_proj_norm_rope_fwd(C.byref(qp), None) != 0 and b"128 (fp8)" in L.magi_last_error()
    qp.hidden = 256
    assert L.magi_qkv_proj_norm_rope_fwd(C.byref(qp), None) != 0 and b"out_scale" in L.magi_last_error()
    assert L.magi_attn_workspace_bytes(5) >= 4 * (4 + 6 + 20)
    assert b"sm_100a" in L.magi_build_info()


def test_python_shim_refuses_cpu_tensors():
    """The product has no CPU path: the shim raises instead of silently computing elsewhere."""
    import torch
    import magi_b200.functional as F
    q = torch.zeros(4, 2, 128, dtype=torch.bfloat16)
    r = torch.tensor([[0, 4]], dtype=torch.int32)
    with pytest.raises(ValueError, match="CUDA"):
        F.attn_ranges(q, q, q, r, r)
    with pytest.raises(ValueError, match="CUDA"):
        F.qk_norm_rope(torch.zeros(4, 256, dtype=torch.bfloat16), None, None, 2, 2,
                       (torch.zeros(128), torch.zeros(128)), None, None, None)
    with pytest.raises(ValueError, match="CUDA"):
        F.apply_rotary_emb(torch.zeros(1, 4, 2, 128), torch.zeros(4, 48), torch.zeros(4, 48))


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under magi-1_b200/ may import or reference it."""
    pkg = os.path.join(ROOT, "magi-1_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "magi_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
