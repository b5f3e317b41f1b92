"""Pair-kernel tuning variants (magi-1_b200/variants/lib_<name>.so given on the command line) on the cfg2 / cross
microbenchmarks, plus two correctness cases each. GPU box only."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from pair_bringup import run  # noqa: E402

if __name__ == "__main__":
    for name in sys.argv[1:]:
        kern = "pair"
        if ":" in name:
            name, kern = name.split(":")
        lib = os.path.join(ROOT, "magi-1_b200", "libmagi_attn.so" if name == "default" else f"variants/lib_{name}.so")
        env = dict(os.environ, MAGI_ATTN_LIB=lib, MAGI_ATTN_KERNEL=kern)
        print(f"===== {name} [{kern}]", flush=True)
        for c in ["g_ragged_ranges", "j_many_items", "p_bench_cfg2", "p2_bench_cfg2_sustained", "s_bench_cross"]:
            run(c, env, timeout=240)
