"""Vendors the part of the UNMODIFIED reference that the same-box parity test needs ("O2": the
reference's own FullyParallelAttention on the B200 with its flash-attn fallback path) into the git-ignored
`baseline/_ref/` - the one place this repository is allowed to hold reference files: never tracked, never imported by
the product, shipped to the GPU box with the gpurun snapshot (`/root/reference` does not exist there).

    python tools/vendor_reference.py            # /root/reference/inference/{common,infra,model/dit} -> baseline/_ref/

`__graft_entry__.build()` calls this when /root/reference is present. tests/test_reference_gpu.py (`-m gpu`) imports
the copy and is skipped when it is absent."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/inference"
DST = os.path.join(ROOT, "baseline", "_ref", "inference")
PARTS = ("common", "infra", os.path.join("model", "dit"))


def vendor(verbose=True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: nothing vendored")
        return False
    for part in PARTS:
        dst = os.path.join(DST, part)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, part), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"vendored {n} reference files into {os.path.relpath(DST, ROOT)} (git-ignored)")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
