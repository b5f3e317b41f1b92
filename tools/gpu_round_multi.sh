#!/bin/bash
# ONE multi-GPU gpurun call (charged N x box time): context-parallel parity with the real kernels, then the scaling bench.
#
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_round_multi.sh r02c 2'
#
# writes gpurun_out/<tag>_cp_check_n<N>.log and <tag>_bench_n<N>[_<mode>].json. Nothing here runs under ncu.
tag="${1:-rXX}"; n="${2:-2}"; quick="${3:-}"
out=gpurun_out
mkdir -p $out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { echo "build failed"; exit 1; }
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
# the default line the driver records (value + e2e + cp_parity + extra legs: ring-KV, configs[3])
timeout 1200 $run --master-port 29710 bench.py --gpus $n > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
echo "bench (default: Ulysses fused + parity + extra legs) rc=$?"; tail -c 1800 $out/${tag}_bench_n${n}.json; tail -3 $out/${tag}_bench_n${n}.err
timeout 600 $run --master-port 29715 bench.py --gpus $n --graph on --no-extra-legs --no-dit-step \
    > $out/${tag}_bench_n${n}_graph.json 2> $out/${tag}_bench_n${n}_graph.err
echo "bench --graph on rc=$?"; cut -c1-200 $out/${tag}_bench_n${n}_graph.json; tail -3 $out/${tag}_bench_n${n}_graph.err
[ -n "$quick" ] && exit 0
CP_CHECK_KV_OFFLOAD=1 timeout 900 $run --master-port 29700 tools/cp_check.py > $out/${tag}_cp_check_n${n}.log 2>&1
echo "cp_check rc=$?"; grep "cp_check\|CP_CHECK" $out/${tag}_cp_check_n${n}.log | tail -12
for mode in nccl ring-nccl; do
  timeout 900 $run --master-port 29720 bench.py --gpus $n --cp-mode $mode --no-dit-step --no-extra-legs \
      > $out/${tag}_bench_n${n}_${mode}.json 2> $out/${tag}_bench_n${n}_${mode}.err
  echo "bench --cp-mode $mode rc=$?"; cut -c1-200 $out/${tag}_bench_n${n}_${mode}.json
done
