tag=r02zq; n=2; out=gpurun_out; mkdir -p $out
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
CP_CHECK_FUSED_GEMM=1 timeout 600 $run --master-port 29700 tools/cp_check.py > $out/${tag}_cp_check_gemm_n${n}.log 2>&1
echo "cp_check (fused gemm) rc=$?"; grep "cp_check\|CP_CHECK\|Error\|error" $out/${tag}_cp_check_gemm_n${n}.log | tail -6
timeout 600 $run --master-port 29701 tools/cp_check.py > $out/${tag}_cp_check_n${n}.log 2>&1
echo "cp_check rc=$?"; grep "CP_CHECK\|FAIL" $out/${tag}_cp_check_n${n}.log | tail -4
