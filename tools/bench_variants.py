"""Runs a correctness check + the cfg2 attention microbench for every library variant in magi-1_b200/variants/
(each in its own process, selected through MAGI_ATTN_LIB). GPU box only."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["g_ragged_ranges", "h_two_segments", "j_many_items", "p_bench_cfg2", "p2_bench_cfg2_sustained", "s_bench_cross"]

if __name__ == "__main__":
    libs = sorted(glob.glob(os.path.join(ROOT, "magi-1_b200", "variants", "lib_*.so")))
    sel = sys.argv[1:]
    for lib in libs:
        name = os.path.basename(lib)[4:-3]
        if sel and name not in sel:
            continue
        env = dict(os.environ, MAGI_ATTN_LIB=lib)
        print(f"===== variant {name}", flush=True)
        for c in CASES:
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_probe.py"), f"--run={c}"], env=env,
                                   capture_output=True, text=True, timeout=240)
                out = [l for l in (r.stdout + r.stderr).strip().splitlines() if l.startswith("[") or "FAIL" in l]
                print("\n".join(out[-4:]) if out else f"[{c}] no output, rc={r.returncode}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"[{c}] TIMEOUT", flush=True)
