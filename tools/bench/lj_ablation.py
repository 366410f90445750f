"""Measurement only: end-build and mate join stage times, windowed join against the whole-file hash join (legacy),
on one workload.  One JSON line per form.

    python tools/bench/lj_ablation.py [--scale 0.4] [--steps 3] [--dbg 0,1,2,4,8,12]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from openge_b200 import dedup, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--scale", type=float, default=0.4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dbg", default="0,1,2,4,8,12")
    ap.add_argument("--product", action="store_true", help="the product library only")
    ap.add_argument("--fused-only", action="store_true", help="skip the legacy (whole-file hash join) form")
    ap.add_argument("--warm", type=int, default=2)
    a = ap.parse_args()
    bam = synth.make(a.workload, a.scale)

    def run_all(label):
        for legacy in ((False, True) if label == "product" and not a.fused_only else (False,)):
            with dedup.context_for(bam, legacy_join=legacy) as ctx:
                ctx.push(bam.records, bam.offsets)
                ms = []
                for i in range(a.steps + a.warm):
                    ctx.run()
                    st = ctx.stats()
                    if i >= a.warm:
                        ms.append((st["ms_endbuild"], st["ms_join"], st["ms_total"]))
                m = np.mean(np.asarray(ms), axis=0)
                print(json.dumps({"workload": a.workload, "reads": bam.n, "lib": label, "legacy": legacy, "ms_endbuild": float(m[0]),
                                  "ms_join": float(m[1]), "ms_total": float(m[2]), "dups": int(st["n_duplicates"]),
                                  "local_pairs": int(st["n_local_pairs"]), "leftovers": int(st["n_join_leftovers"]),
                                  "local_retracted": int(st["n_local_retracted"]), "complex": int(st["n_complex_names"])}), flush=True)

    run_all("product")
    if not a.product:
        with dedup.testing_library():
            run_all("testing")


if __name__ == "__main__":
    main()
