tag=r02zo; out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 600 python -m pytest tests -m gpu -q -k "projection or transformer_layer or two_block or graphed or full_attention_block" > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python tools/gpu_probe.py --run=u_bench_qkv_gemm 2>&1 | grep "^\[\|fused" | tail -12 | tee $out/${tag}_gemm_bench.log
