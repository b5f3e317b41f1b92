tag=r02zg; n=2; out=gpurun_out; mkdir -p $out
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
CP_CHECK_KV_OFFLOAD=1 timeout 900 $run --master-port 29700 tools/cp_check.py > $out/${tag}_cp_check_n${n}.log 2>&1
echo "cp_check rc=$?"; grep "cp_check\|CP_CHECK\|Error\|error" $out/${tag}_cp_check_n${n}.log | tail -16
