"""ctypes binding of libmagi_attn.so (the C-ABI declared in include/magi_attn.h).

The shared library is the product; this module only mirrors its structs. There is no fallback: if the library is
missing or was not built, importing the compute entry points raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# MAGI_ATTN_LIB selects another build of the same library (kernel tuning variants, tools/bench_variants.py)
LIB_PATH = os.environ.get("MAGI_ATTN_LIB") or os.path.join(_HERE, "libmagi_attn.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

# Every symbol include/magi_attn.h declares (checked by tests/test_abi.py against the header text).
EXPORTS = (
    "magi_attn_workspace_bytes",
    "magi_attn_fwd",
    "magi_attn_plan",
    "magi_qk_norm_rope_fwd",
    "magi_kv_cache_append",
    "magi_lse_merge",
    "magi_ulysses_relayout",
    "magi_cso_unshuffle",
    "magi_bias_modulate_add",
    "magi_apply_rotary_emb",
    "magi_fp8_quantize",
    "magi_mlp_activation",
    "magi_qkv_proj_norm_rope_fwd",
    "magi_last_error",
    "magi_build_info",
    "magi_launch_count",
)


MAX_PEERS = 8   # MAGI_MAX_PEERS


class MagiView(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("tokens", C.c_int64),
        ("token_stride", C.c_int64),
        ("head_stride", C.c_int64),
    ]


class MagiAttnArgs(C.Structure):
    _fields_ = [
        ("q", MagiView),
        ("k_a", MagiView),
        ("v_a", MagiView),
        ("k_b", MagiView),
        ("v_b", MagiView),
        ("out", C.c_void_p),
        ("out_token_stride", C.c_int64),
        ("out_head_stride", C.c_int64),
        ("lse", C.c_void_p),
        ("q_ranges", C.c_void_p),
        ("k_ranges", C.c_void_p),
        ("cu_seqlens_q", C.c_void_p),
        ("cu_seqlens_k", C.c_void_p),
        ("n_ranges", C.c_int32),
        ("num_heads_q", C.c_int32),
        ("num_heads_kv", C.c_int32),
        ("head_dim", C.c_int32),
        ("softmax_scale", C.c_float),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("num_sms", C.c_int32),
        ("out_group_heads", C.c_int32),
        ("out_group_stride", C.c_int64),
        ("num_peers", C.c_int32),
        ("peer_head_offset", C.c_int32),
        ("peer_out", C.c_void_p * MAX_PEERS),
        ("peer_row_bounds", C.c_int64 * (MAX_PEERS + 1)),
        ("kernel", C.c_int32),
        ("plan_ready", C.c_int32),
        ("launch_dependency", C.c_int32),
        ("fill_uncovered_rows", C.c_int32),
    ]


class MagiPrologueArgs(C.Structure):
    _fields_ = [
        ("x_q", C.c_void_p),
        ("x_k", C.c_void_p),
        ("x_v", C.c_void_p),
        ("rows", C.c_int64),
        ("xq_row_stride", C.c_int64),
        ("xk_row_stride", C.c_int64),
        ("xv_row_stride", C.c_int64),
        ("xq_head_stride", C.c_int64),
        ("xk_head_stride", C.c_int64),
        ("xv_head_stride", C.c_int64),
        ("num_heads_q", C.c_int32),
        ("num_heads_kv", C.c_int32),
        ("batch", C.c_int32),
        ("ln_q_w", C.c_void_p),
        ("ln_q_b", C.c_void_p),
        ("ln_k_w", C.c_void_p),
        ("ln_k_b", C.c_void_p),
        ("ln_dtype", C.c_int32),
        ("eps", C.c_float),
        ("zero_centered_gamma", C.c_int32),
        ("sin", C.c_void_p),
        ("cos", C.c_void_p),
        ("q_out", C.c_void_p),
        ("q_out_token_stride", C.c_int64),
        ("q_out_head_stride", C.c_int64),
        ("kv_out", C.c_void_p),
        ("kv_out_token_stride", C.c_int64),
        ("kv_out_head_stride", C.c_int64),
        ("cache", C.c_void_p),
        ("cache_token_stride", C.c_int64),
        ("cache_head_stride", C.c_int64),
        ("cache_row_offset", C.c_int64),
        ("cache_rows", C.c_int64),
        ("head_perm_cp", C.c_int32),
        ("num_peers", C.c_int32),
        ("peer_q_out", C.c_void_p * MAX_PEERS),
        ("peer_kv_out", C.c_void_p * MAX_PEERS),
        ("peer_row_offset", C.c_int64),
        ("gather_peers", C.c_int32),
        ("gather_kv_out", C.c_void_p * MAX_PEERS),
        ("gather_rows_per_chunk", C.c_int64),
        ("gather_chunk_stride", C.c_int64),
        ("gather_row_offset", C.c_int64),
        ("gather_multicast", C.c_int32),
        ("x_qx", C.c_void_p),
        ("xqx_row_stride", C.c_int64),
        ("xqx_head_stride", C.c_int64),
        ("ln_qx_w", C.c_void_p),
        ("ln_qx_b", C.c_void_p),
        ("qx_out", C.c_void_p),
        ("qx_out_token_stride", C.c_int64),
        ("qx_out_head_stride", C.c_int64),
        ("x_tk", C.c_void_p),
        ("x_tv", C.c_void_p),
        ("rows_t", C.c_int64),
        ("xt_row_stride", C.c_int64),
        ("xt_head_stride", C.c_int64),
        ("ln_tk_w", C.c_void_p),
        ("ln_tk_b", C.c_void_p),
        ("tkv_out", C.c_void_p),
        ("tkv_out_token_stride", C.c_int64),
        ("tkv_out_head_stride", C.c_int64),
    ]


class MagiQkvProjArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("rows", C.c_int64),
        ("x_row_stride", C.c_int64),
        ("hidden", C.c_int32),
        ("w", C.c_void_p),
        ("w_row_stride", C.c_int64),
        ("num_heads_q", C.c_int32),
        ("num_heads_kv", C.c_int32),
        ("num_heads_qx", C.c_int32),
        ("batch", C.c_int32),
        ("ln_q_w", C.c_void_p),
        ("ln_q_b", C.c_void_p),
        ("ln_k_w", C.c_void_p),
        ("ln_k_b", C.c_void_p),
        ("ln_dtype", C.c_int32),
        ("ln_qx_w", C.c_void_p),
        ("ln_qx_b", C.c_void_p),
        ("eps", C.c_float),
        ("zero_centered_gamma", C.c_int32),
        ("sin", C.c_void_p),
        ("cos", C.c_void_p),
        ("q_out", C.c_void_p),
        ("q_out_token_stride", C.c_int64),
        ("q_out_head_stride", C.c_int64),
        ("kv_out", C.c_void_p),
        ("kv_out_token_stride", C.c_int64),
        ("kv_out_head_stride", C.c_int64),
        ("qx_out", C.c_void_p),
        ("qx_out_token_stride", C.c_int64),
        ("qx_out_head_stride", C.c_int64),
        ("cache", C.c_void_p),
        ("cache_token_stride", C.c_int64),
        ("cache_head_stride", C.c_int64),
        ("cache_row_offset", C.c_int64),
        ("cache_rows", C.c_int64),
        ("num_sms", C.c_int32),
        ("num_peers", C.c_int32),
        ("peer_q_out", C.c_void_p * MAX_PEERS),
        ("peer_kv_out", C.c_void_p * MAX_PEERS),
        ("peer_row_offset", C.c_int64),
        ("fp8", C.c_int32),
        ("out_scale", C.c_void_p),
    ]


class MagiGatingArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_dtype", C.c_int32), ("x_row_stride", C.c_int64),
        ("residual", C.c_void_p), ("residual_row_stride", C.c_int64),
        ("cond_map", C.c_void_p), ("cond_map_is_int64", C.c_int32), ("cond_map_stride", C.c_int64),
        ("gate", C.c_void_p), ("gate_dtype", C.c_int32), ("gate_row_stride", C.c_int64), ("num_gates", C.c_int64),
        ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
        ("eps", C.c_float), ("zero_centered_gamma", C.c_int32),
        ("out", C.c_void_p), ("out_row_stride", C.c_int64),
        ("rows", C.c_int64), ("hidden", C.c_int32),
    ]


class MagiRotaryArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("out", C.c_void_p), ("dtype", C.c_int32),
        ("batch", C.c_int64), ("seqlen", C.c_int64),
        ("heads", C.c_int32), ("head_dim", C.c_int32), ("rotary_dim", C.c_int32),
        ("x_batch_stride", C.c_int64), ("x_seq_stride", C.c_int64), ("x_head_stride", C.c_int64),
        ("out_batch_stride", C.c_int64), ("out_seq_stride", C.c_int64), ("out_head_stride", C.c_int64),
        ("cos", C.c_void_p), ("sin", C.c_void_p),
    ]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libmagi_attn.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j4"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libmagi_attn.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built: there is no CPU or eager fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C magi-1_b200/csrc`). This package has no fallback path."
        )
    L = C.CDLL(LIB_PATH)
    L.magi_attn_workspace_bytes.restype = C.c_size_t
    L.magi_attn_workspace_bytes.argtypes = [C.c_int32]
    L.magi_attn_fwd.restype = C.c_int
    L.magi_attn_fwd.argtypes = [C.POINTER(MagiAttnArgs), C.c_void_p]
    L.magi_attn_plan.restype = C.c_int
    L.magi_attn_plan.argtypes = [C.POINTER(MagiAttnArgs), C.c_void_p]
    L.magi_qk_norm_rope_fwd.restype = C.c_int
    L.magi_qk_norm_rope_fwd.argtypes = [C.POINTER(MagiPrologueArgs), C.c_void_p]
    L.magi_kv_cache_append.restype = C.c_int
    L.magi_kv_cache_append.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                       C.c_void_p]
    L.magi_lse_merge.restype = C.c_int
    L.magi_lse_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64,
                                 C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    L.magi_ulysses_relayout.restype = C.c_int
    L.magi_ulysses_relayout.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.magi_cso_unshuffle.restype = C.c_int
    L.magi_cso_unshuffle.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64,
                                     C.c_int64, C.c_int64, C.c_void_p]
    L.magi_bias_modulate_add.restype = C.c_int
    L.magi_bias_modulate_add.argtypes = [C.POINTER(MagiGatingArgs), C.c_void_p]
    L.magi_apply_rotary_emb.restype = C.c_int
    L.magi_apply_rotary_emb.argtypes = [C.POINTER(MagiRotaryArgs), C.c_void_p]
    L.magi_qkv_proj_norm_rope_fwd.restype = C.c_int
    L.magi_qkv_proj_norm_rope_fwd.argtypes = [C.POINTER(MagiQkvProjArgs), C.c_void_p]
    L.magi_fp8_quantize.restype = C.c_int
    L.magi_fp8_quantize.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_int32, C.c_void_p]
    L.magi_mlp_activation.restype = C.c_int
    L.magi_mlp_activation.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                      C.c_void_p]
    L.magi_last_error.restype = C.c_char_p
    L.magi_build_info.restype = C.c_char_p
    L.magi_launch_count.restype = C.c_uint64
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().magi_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what}: {msg}")


def source_hash() -> str:
    """sha256 (first 16 hex digits) over the library's sources, computed the way csrc/Makefile stamps it into
    magi_build_info(): sorted *.cu and *.cuh of csrc/, then include/magi_attn.h."""
    import glob
    import hashlib
    files = sorted(glob.glob(os.path.join(CSRC_DIR, "*.cu")) + glob.glob(os.path.join(CSRC_DIR, "*.cuh")),
                   key=os.path.basename)
    files.append(os.path.join(os.path.dirname(_HERE), "include", "magi_attn.h"))
    h = hashlib.sha256()
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def binary_is_current() -> bool:
    """True when the loaded library was built from the sources in the tree (variant builds selected through
    MAGI_ATTN_LIB are exempt: they are built with other flags on purpose)."""
    if os.environ.get("MAGI_ATTN_LIB"):
        return True
    return ("source " + source_hash()) in lib().magi_build_info().decode("utf-8", "replace")


def launch_count() -> int:
    return int(lib().magi_launch_count())


def attn_kernel_name() -> str:
    """Which attention kernel magi_attn_fwd launches in this process (MAGI_ATTN_KERNEL=single|pair overrides the
    library's built-in default, reported by magi_build_info)."""
    e = os.environ.get("MAGI_ATTN_KERNEL", "")
    if e == "single":
        return "attn_fwd_kernel"
    if e == "pair":
        return "attn_pair_kernel"
    info = lib().magi_build_info().decode("utf-8", "replace")
    return "attn_pair_kernel" if "default attention kernel: pair" in info else "attn_fwd_kernel"
