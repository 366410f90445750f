"""A/B of the onesweep pass-kernel variants (K3) on one GPU.  Writes gpurun_out/sort_variants.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openge_b200 import dedup  # noqa: E402

out = []
variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,2".split(","))]
for n, lo, hi in [(25_000_000, 42, 112), (50_000_000, 43, 79)]:
    for mode in (0, 1):
        for v in variants:
            r = dedup.debug_sort_bench(n, lo, hi, variant=v, mode=mode, reps=3)
            out.append(r)
            print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "sort_variants.json"), "w") as f:
    json.dump(out, f, indent=1)
