#!/bin/bash
# ONE gpurun call that produces everything a round's profiles/ entry needs, in the order the profiling recipe asks for
# (plain run exits 0 first, then the same command under ncu; one GPU; nothing multi-rank under ncu).
#
#   gpurun --timeout 1700 -- 'bash tools/gpu_round.sh r02a'
#
# writes gpurun_out/<tag>_pytest_gpu.log, _smoke.log, _bench_n1.json, _bench_n1_graph.json, _bench_cfg1[_graph].json,
# _bench_reference.json, _launches.csv, _ncu_raw.csv (raw page of the `--set full` capture of one layer's kernels),
# _sass_attn_fwd.txt - copy what is to be judged to profiles/.
tag="${1:-rXX}"
out=gpurun_out
mkdir -p $out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { echo "build failed"; tail -20 $out/${tag}_build.log; exit 1; }
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?"; cut -c1-400 $out/${tag}_bench_n1.json
timeout 600 python bench.py --graph on --no-cpu-baseline --no-dit-step --no-gpu-reference > $out/${tag}_bench_n1_graph.json 2> $out/${tag}_bench_n1_graph.err; echo "bench --graph on rc=$?"; cut -c1-200 $out/${tag}_bench_n1_graph.json
for g in off on; do
  timeout 300 python bench.py --workload cfg1 --graph $g --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-reference \
      > $out/${tag}_bench_cfg1_graph_$g.json 2> $out/${tag}_bench_cfg1_graph_$g.err
  echo "cfg1 graph $g rc=$?"; cut -c1-160 $out/${tag}_bench_cfg1_graph_$g.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --ref-budget 5 > $out/${tag}_bench_reference.json 2>&1; echo "reference arm rc=$?"
cuobjdump -sass magi-1_b200/libmagi_attn.so | awk '/Function : /{f=0} /Function : .*attn_fwd_kernel/{f=1} f' | cut -c1-110 > $out/${tag}_sass_attn_fwd.txt
# launch list of one profiled step (cold-cache, serialised: shares, not absolutes)
P="python bench.py --profile --steps 1 --warmup 1"
timeout 600 $P > $out/${tag}_profile_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $P \
    > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# full capture of one layer's kernels after the first (warm-up) step: self-attention + prologues + cross-attention
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_kernel|prologue_ring_kernel' \
    -s 6 -c 5 -o $out/${tag}_prof $P > $out/${tag}_ncu_full.log 2>&1
echo "full capture rc=$?"
[ -f $out/${tag}_prof.ncu-rep ] && ncu -i $out/${tag}_prof.ncu-rep --page raw --csv > $out/${tag}_ncu_raw.csv 2>/dev/null
# fp8_quant mode: bench probe (plain, then one full capture of the quantiser and the fp8 GEMM at cfg2 shapes)
timeout 200 python tools/gpu_probe.py --run=v_bench_fp8_qkv > $out/${tag}_fp8_bench.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fp8_quant_kernel|qkv_gemm_kernel' -s 8 -c 2 \
    -o $out/${tag}_fp8prof python tools/gpu_probe.py --run=v_bench_fp8_qkv > $out/${tag}_ncu_fp8.log 2>&1
echo "fp8 probe + capture rc=$?"; grep "quantiser alone\|fused fp8" $out/${tag}_fp8_bench.log | cut -c1-160
[ -f $out/${tag}_fp8prof.ncu-rep ] && ncu -i $out/${tag}_fp8prof.ncu-rep --page raw --csv > $out/${tag}_ncu_fp8_raw.csv 2>/dev/null
ls -la $out | tail -20
