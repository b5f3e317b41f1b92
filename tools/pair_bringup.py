"""Bring-up of the CTA-pair attention kernel on a B200 box: TMEM fragment-layout probe, the diagnostic attention cases
of tools/gpu_probe.py with MAGI_ATTN_KERNEL=pair (debug build: 1 s mbarrier watchdog), then - only if all pass - the
cfg2 / cross-attention microbenchmarks for both kernels."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEBUG = os.path.join(ROOT, "magi-1_b200", "variants", "lib_debug.so")
PROBE = os.path.join(ROOT, "tools", "gpu_probe.py")


def run(case, env, timeout=75):
    try:
        r = subprocess.run([sys.executable, PROBE, f"--run={case}"], env=env, capture_output=True, text=True,
                           timeout=timeout)
        out = [l for l in (r.stdout + r.stderr).strip().splitlines()
               if l.startswith("[") or "FAIL" in l or "magi" in l or "rror" in l or l.startswith("    ")]
        print("\n".join(out[-12:]) if out else f"[{case}] no output, rc={r.returncode}", flush=True)
        return r.returncode == 0
    except subprocess.TimeoutExpired:
        print(f"[{case}] TIMEOUT", flush=True)
        return False


if __name__ == "__main__":
    if "--layout" in sys.argv or len(sys.argv) == 1:
        lib = C.CDLL(DEBUG)
        err = (C.c_int * 3)()
        rc = lib.magi_debug_tmem_layout(err)
        print(f"[tmem layout probe] rc={rc} mismatches: ld16x256b={err[0]} st16x128b={err[1]} ldst16x256b.x4={err[2]}",
              flush=True)
    env = dict(os.environ, MAGI_ATTN_LIB=DEBUG, MAGI_ATTN_KERNEL="pair")
    ok = True
    for c in ["a_veye_1tile", "b_kzero_1tile", "c_onehot_1tile", "d_rand_1tile", "e_rand_4tiles", "f_rand_packed_gqa",
              "g_ragged_ranges", "h_two_segments", "i_empty_k_and_gap", "j_many_items"]:
        ok = run(c, env) and ok
    if ok or "--bench" in sys.argv:
        for kern in ("pair", "single"):
            env = dict(os.environ, MAGI_ATTN_KERNEL=kern)
            print(f"===== kernel {kern}", flush=True)
            for c in ["p_bench_cfg2", "p2_bench_cfg2_sustained", "s_bench_cross"]:
                run(c, env, timeout=240)
