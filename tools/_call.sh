tag=r02zi; n=8; out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { echo "build failed"; exit 1; }
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 900 $run --master-port 29710 bench.py --gpus $n > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("$out/${tag}_bench_n${n}.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("n8", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "achieved", r["achieved"], "share", r["share_of_step"], "parity", d.get("cp_parity",{}).get("max_abs"), r.get("phases_us_per_layer"), (d.get("dit_denoise_step") or {}).get("ms_per_step"), {k:(v.get("value"),v.get("cp_parity_ok")) for k,v in (d.get("extra_legs") or {}).items()}, "launches", d["gpu_launches"], d["clocks"])
PY
tail -3 $out/${tag}_bench_n${n}.err
