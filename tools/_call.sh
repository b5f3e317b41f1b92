tag=r02u; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("$out/${tag}_bench_n1.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "achieved", r["achieved"], r["avg_launch_ms"], "frac", r["frac"], "share", r["share_of_step"], d["clocks"], r.get("phases_us_per_layer"), d.get("dit_denoise_step"), d.get("cpu_baseline"), d.get("gpu_reference"))
PY
