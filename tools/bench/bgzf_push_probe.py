"""push_bgzf on its own, with the trace of the library's -DOGE_TESTING build (OGE_TRACE_PUSH): how the pieces of the upload
and their inflates line up on the device, for a few piece sizes.  Measurement tool.

    OGE_TRACE_PUSH=1 python tools/bench/bgzf_push_probe.py --scale 0.2
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scale", type=float, default=0.2)
    ap.add_argument("--pageable-table", action="store_true")
    a = ap.parse_args()
    import numpy as np
    from openge_b200 import bamhost, bamio, dedup, synth
    bam = synth.make(a.config, a.scale, seed=2)
    raw = bamio.serialize_bam_stream(bam)
    z = bamhost.bgzf_compress(raw, 1)
    pin = dedup.PinnedBuffer(len(z) + 64)
    comp = pin.array[:len(z)]
    comp[:] = np.frombuffer(z, dtype=np.uint8)
    in_off, csize, isize = [], [], []
    pos = 0
    while pos < len(z):
        bs = int(comp[pos + 16]) + (int(comp[pos + 17]) << 8) + 1
        in_off.append(pos); csize.append(bs); isize.append(int.from_bytes(z[pos + bs - 4: pos + bs], "little"))
        pos += bs
    in_off = np.asarray(in_off, dtype=np.uint64); csize = np.asarray(csize, dtype=np.uint32); isize = np.asarray(isize, dtype=np.uint32)
    head = len(raw) - bam.records.nbytes
    with dedup.testing_library(), dedup.context_for(bam, device=0) as ctx:      # the product library carries no trace hook
        for piece in (0, 16 << 20, 2 ** 64 - 1, 0):
            dedup.set_bgzf_chunk_bytes(piece)
            for rep in range(2):
                ctx.reset()
                ctx.sync()
                t0 = time.perf_counter()
                ctx.push_bgzf(comp.ctypes.data, comp.nbytes, in_off.ctypes.data, csize.ctypes.data, isize.ctypes.data, len(in_off), head, None)
                t1 = time.perf_counter()
                st = ctx.stats()
                print(json.dumps({"piece_bytes": piece, "rep": rep, "wall_ms": (t1 - t0) * 1e3, "ms_push_bgzf": st["ms_push_bgzf"], "ms_upload": st["ms_inflate_h2d"],
                                  "ms_inflate_span": st["ms_inflate"], "first_inflate_at": st["ms_inflate_start"], "pieces": st["inflate_pieces"],
                                  "bytes_in": st["inflate_bytes_in"], "bytes_out": st["inflate_bytes_out"]}), flush=True)
        n = ctx.frame(bam.records.nbytes)
        ctx.run()
        print(json.dumps({"records": n, "duplicates": ctx.stats()["n_duplicates"]}))
    pin.free()


if __name__ == "__main__":
    main()
