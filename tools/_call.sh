tag=r02o; out=gpurun_out; mkdir -p $out
timeout 900 python tools/ab_variants.py --rounds 5 2>&1 | tee $out/${tag}_ab.log
