"""One DiT denoise forward (all blocks, bench.py's dit_step_leg) for an `ncu --metrics gpu__time_duration.sum` launch
list: which kernels make up the step besides the attention hot path.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/dit_step_profile.py
    python tools/dit_step_profile.py --summarise out.csv"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if "--summarise" in sys.argv:
    path = sys.argv[sys.argv.index("--summarise") + 1]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[hdr.index("Metric Unit")]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        name = r[ik].split("(")[0][-70:]
        tot[name] += us
        cnt[name] += 1
    total = sum(tot.values())
    print(f"total {total / 1e3:.1f} ms over {sum(cnt.values())} launches")
    for name, us in tot.most_common(25):
        print(f"{us / 1e3:9.2f} ms {100 * us / total:5.1f} %  x{cnt[name]:4d}  {name}")
    sys.exit(0)

import torch  # noqa: E402

import bench  # noqa: E402

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg2"])
w["layers"] = int(os.environ.get("DIT_LAYERS", "4"))
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
r = bench.dit_step_leg(w, dev, 1, 0, 1, lambda v: v, torch.cuda.synchronize)
print(r)
