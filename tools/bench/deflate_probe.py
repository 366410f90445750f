"""The device deflate on its own: records pushed raw, dedup run, oge_gpu_dedup_deflate timed (CUDA events inside the
library), the members inflated on the host and compared with the flag-patched records.  Measurement tool (also the
program ncu captures bgzf_deflate_warps from).

    python tools/bench/deflate_probe.py --config C2 --scale 0.05 [--reps 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-verify", action="store_true")
    a = ap.parse_args()
    import numpy as np
    from openge_b200 import bamhost, dedup, synth
    bam = synth.make(a.config, a.scale, seed=2)
    with dedup.context_for(bam, device=0) as ctx:
        ctx.push(bam.records, bam.offsets)
        ctx.run()
        ms = []
        for _ in range(a.reps):
            members, blocks, nrec = ctx.deflate()
            ms.append(ctx.stats()["ms_deflate"])
        st = ctx.stats()
        ok = None
        if not a.no_verify:
            rec, off = ctx.pull()      # flag-patched, bins recomputed in place by deflate
            ok = bamhost.bgzf_decompress(members.tobytes()) == rec.tobytes()
    print(json.dumps({"config": a.config, "scale": a.scale, "records": int(nrec), "blocks": int(blocks), "bytes_in": st["deflate_bytes_in"],
                      "bytes_out": st["deflate_bytes_out"], "ratio": st["deflate_bytes_in"] / max(1, st["deflate_bytes_out"]), "ms_deflate": ms,
                      "in_GBps": [st["deflate_bytes_in"] / 1e6 / m for m in ms], "inflates_back_to_the_records": ok}))


if __name__ == "__main__":
    main()
