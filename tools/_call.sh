tag=r02zm; out=gpurun_out; mkdir -p $out
timeout 600 python tools/ab_variants.py --rounds 4 --launches 30 2>&1 | tee $out/${tag}_ab.log
