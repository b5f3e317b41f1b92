tag=r02ze; out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $out/${tag}_pytest_gpu.log
Q="--no-cpu-baseline --no-gpu-reference"
timeout 600 python bench.py $Q > $out/${tag}_bench_a.json 2> $out/${tag}_bench_a.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_a.err
python - <<PY
import json
for l in open("$out/${tag}_bench_a.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "achieved", r["achieved"], r["avg_launch_ms"], "share", r["share_of_step"], d["clocks"]["sm_mhz"], "launches", d["gpu_launches"], r.get("phases_us_per_layer"), (d.get("dit_denoise_step") or {}).get("ms_per_step"))
PY
