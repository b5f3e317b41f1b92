#!/bin/bash
# Builds (HERE, before the gpurun call: the .so files travel with the snapshot) the compiled-out attention-kernel
# candidates of DESIGN.md section 4.1, and `base` = the default flags for the A/B on the same box. On the box:
#   gpurun --timeout 900 -- 'python tools/bench_variants.py > gpurun_out/r02_variants.log 2>&1; tail -60 gpurun_out/r02_variants.log'
# (correctness cases first, then the cfg2 / single-range / cross-attention microbenchmarks, per variant).
cd "$(dirname "$0")/.."
tools/build_variants.sh \
  base="" \
  static="-DMAGI_STATIC_BARRIERS=1" \
  staticearly="-DMAGI_STATIC_BARRIERS=1 -DMAGI_EARLY_PUBLISH=1" \
  early="-DMAGI_EARLY_PUBLISH=1" \
  r216="-DMAGI_SOFTMAX_REGS=216 -DMAGI_AUX_REGS=72" \
  static216="-DMAGI_STATIC_BARRIERS=1 -DMAGI_SOFTMAX_REGS=216 -DMAGI_AUX_REGS=72"
