tag=r02zb; out=gpurun_out; mkdir -p $out
for c in g_ragged_ranges h_two_segments j_many_items; do MAGI_ATTN_LIB=magi-1_b200/variants/lib_auto.so timeout 240 python tools/gpu_probe.py --run=$c 2>&1 | grep "^\[" | tail -1; done | tee $out/${tag}_probe.log
timeout 600 python tools/ab_variants.py --rounds 5 --launches 30 2>&1 | tee $out/${tag}_ab.log
