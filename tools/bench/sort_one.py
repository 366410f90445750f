"""One sort-bench configuration (for ncu captures): python tools/bench/sort_one.py VARIANT [N] [LO] [HI] [MODE] [REPS]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openge_b200 import dedup  # noqa: E402

a = sys.argv[1:]
variant = int(a[0]) if a else 0
n = int(a[1]) if len(a) > 1 else 25_000_000
lo = int(a[2]) if len(a) > 2 else 42
hi = int(a[3]) if len(a) > 3 else 112
mode = int(a[4]) if len(a) > 4 else 0
reps = int(a[5]) if len(a) > 5 else 3
print(json.dumps(dedup.debug_sort_bench(n, lo, hi, variant=variant, mode=mode, reps=reps)))
