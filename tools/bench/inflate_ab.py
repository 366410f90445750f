"""The BGZF decoder on its own: CUDA-event time of oge_gpu_dedup_push_bgzf's inflate for one decoder, the file uploaded
in ONE piece first (no overlap with the upload, so the time is the decoder's).

    python tools/bench/inflate_ab.py --config C2 --scale 0.1 --level 1 --mode engine|threads|warp [--reps 3]

Run once per mode (the form is chosen by OGE_INFLATE_KERNEL when the library first launches it).  Prints one JSON
line; the inflated bytes are verified against the generator's stream every time."""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--mode", default="threads", choices=["engine", "threads", "warp"])
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    os.environ["OGE_INFLATE_KERNEL"] = a.mode
    import numpy as np
    from openge_b200 import bamhost, bamio, dedup, synth
    dedup.set_bgzf_chunk_bytes(2 ** 64 - 1)
    assert dedup.inflate_kernel() == a.mode, "decoder %s is not available here" % a.mode
    bam = synth.make(a.config, a.scale, seed=2)
    raw = bamio.serialize_bam_stream(bam)
    with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
        p = os.path.join(d, "in.bam")
        with open(p, "wb") as f:
            f.write(bamhost.bgzf_compress(raw, a.level))
        fsize = os.path.getsize(p)
        ms = []
        for rep in range(1 + a.reps):
            with bamhost.HostBam(p, defer_inflate=True) as h, dedup.DedupContext(n_ref=len(h.refs), max_ref_len=max(l for _, l in h.refs)) as ctx:
                ix = h.bgzf_index()
                ctx.push_bgzf(ix["comp"], ix["comp_bytes"], ix["in_off"], ix["csize"], ix["isize"], ix["n_blocks"], ix["header_bytes"],
                              h.records_buffer() if rep == 0 else None)
                st = ctx.stats()
                if rep == 0:
                    h.frame_records()
                    assert np.array_equal(h.records, bam.records), "inflated bytes differ"
                else:
                    ms.append(st["ms_inflate"])
        out_bytes = len(raw)
        print(json.dumps({"mode": a.mode, "config": a.config, "scale": a.scale, "level": a.level, "blocks": ix["n_blocks"],
                          "bytes_in": fsize, "bytes_out": out_bytes, "ms_kernel": ms, "out_GBps": [out_bytes / 1e6 / m for m in ms],
                          "in_GBps": [fsize / 1e6 / m for m in ms], "verified": True}))


if __name__ == "__main__":
    main()
