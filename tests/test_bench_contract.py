"""bench.py contract checks that run without a GPU: the reference arm (the torch fp32 CPU port of the reference's attention loop, timed on the host cores) prints
exactly one JSON line with the keys the driver reads, and the FLOP accounting follows SURVEY.md section 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "0", "--ref-budget", "0.2"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "chunk_causal_attn_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_flop_accounting_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS["cfg2"]
    self_f, cross_f = bench.attn_flops(w)
    # 4 * hd * Hq * sum_i Lq_i * Lk_i with k = 1..4 chunks of 9360: 10.77 TFLOP per layer
    assert self_f == 4 * 128 * 24 * 9360 * 9360 * (1 + 2 + 3 + 4)
    assert abs(self_f / 1e12 - 10.77) < 0.01
    assert cross_f == 4 * 128 * 24 * 9360 * (50 + 512 + 512 + 512)
    w4 = bench.WORKLOADS["cfg4"]
    assert abs(bench.attn_flops(w4)[0] / 1e12 - 172.0) < 0.1


def test_both_arms_name_the_same_config():
    """`config` names the workload and is identical in both arms (ours / --impl reference), so the driver's
    same_config check can hold; everything measured lives elsewhere in the line."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(cp_mode="fused", overlap_degree=1, workload="cfg2", gpus=1)
    c1 = bench.static_config(a, bench.WORKLOADS["cfg2"], 1)
    assert c1 == bench.static_config(a, bench.WORKLOADS["cfg2"], 1)
    assert {"workload", "name", "parallelism", "step", "flops_per_step"} <= set(c1) and "model" not in c1
    assert "ulysses" in bench.static_config(a, bench.WORKLOADS["cfg2"], 8)["parallelism"]
    a.cp_mode = "ring"
    assert "multimem" in bench.static_config(a, bench.WORKLOADS["cfg2"], 2)["parallelism"]


def test_prologue_byte_accounting():
    """roofline.prologue_hbm: the algorithmic bytes of the layer's one prologue launch at cfg2 - projection outputs in
    (q, k, v, qx, text [K|V], sin / cos rows), q / packed [K|V] / its cache copy / qx / text [K|V] out."""
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS["cfg2"]
    S, T = 4 * 9360, 50 + 3 * 512
    q, kv = S * 24 * 128 * 2, S * 8 * 256 * 2
    expect = (2 * q + kv + S * 96 * 4 + T * 8 * 256 * 2) + (2 * q + 2 * kv + T * 8 * 256 * 2)
    assert bench.prologue_bytes(w) == expect
    assert abs(expect / 1e6 - 1407.6) < 0.1
