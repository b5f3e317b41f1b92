tag=r02q; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
timeout 600 python tools/ab_variants.py --rounds 4 2>&1 | tee $out/${tag}_ab.log
Q="--no-cpu-baseline --no-gpu-reference --no-dit-step"
timeout 600 python bench.py $Q > $out/${tag}_bench_a.json 2> $out/${tag}_bench_a.err; echo "bench rc=$?"
MAGI_BENCH_PDL_OVERLAP=0 timeout 600 python bench.py $Q > $out/${tag}_bench_nooverlap.json 2> $out/${tag}_bench_nooverlap.err; echo "bench no-overlap rc=$?"
MAGI_ATTN_LIB=magi-1_b200/variants/lib_static.so timeout 600 python bench.py $Q > $out/${tag}_bench_static.json 2> $out/${tag}_bench_static.err; echo "bench static rc=$?"
timeout 600 python bench.py $Q > $out/${tag}_bench_b.json 2> $out/${tag}_bench_b.err; echo "bench rc=$?"
for f in a nooverlap static b; do python - <<PY
import json
for l in open("$out/${tag}_bench_$f.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("$f", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "achieved", r["achieved"], r["avg_launch_ms"], "share", r["share_of_step"], d["clocks"]["sm_mhz"], r.get("phases_us_per_layer"))
PY
done
