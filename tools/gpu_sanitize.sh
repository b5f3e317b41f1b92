#!/bin/bash
# ONE gpurun call, ONE compute-sanitizer tool (the profiling recipe allows one tool per call, and only on a program
# that has just exited 0 without it): the smoke pass - prologue with cache append, range-masked attention over
# cache ++ new tokens, varlen cross-attention with an empty segment - under memcheck (default), racecheck, initcheck or
# synccheck. Run it in a call of its own, never together with ncu.
# NOTE (round 2): this pool refuses compute-sanitizer ("closed on this pool: runs under it have left GPUs needing a
# reset"); the write-canary / NaN-padding / bit-identity tests of tests/test_gpu_parity.py stand in for it.
#
#   gpurun --timeout 900 -- 'bash tools/gpu_sanitize.sh memcheck r02a'
tool="${1:-memcheck}"; tag="${2:-rXX}"
out=gpurun_out
mkdir -p $out
cd "$(dirname "$0")/.."
cmd='import __graft_entry__ as g; g.build(); g.smoke()'
timeout 300 python -c "$cmd" > $out/${tag}_smoke_plain.log 2>&1 || { echo "plain smoke failed: not sanitizing"; tail -5 $out/${tag}_smoke_plain.log; exit 1; }
timeout 800 compute-sanitizer --tool "$tool" --print-limit 20 python -c "$cmd" > $out/${tag}_sanitizer_${tool}.log 2>&1
echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|smoke:" $out/${tag}_sanitizer_${tool}.log | tail -3
