#!/bin/bash
# GPU box: prologue bandwidth for every pro_* variant + the default build
cd "$(dirname "$0")/.."
echo "== default"; python tools/gpu_probe.py --run=r_bench_prologue 2>&1 | grep "^\["
for f in magi-1_b200/variants/lib_pro_*.so; do echo "== $f"; MAGI_ATTN_LIB=$f python tools/gpu_probe.py --run=r_bench_prologue 2>&1 | grep "^\["; done
