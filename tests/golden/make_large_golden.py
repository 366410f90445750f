"""Regenerates tests/golden/large_pins.npz: the COMPILED REFERENCE's own flags (oracle/_ref/oge_ref_dedup --mem -v
--flags: MarkDuplicates::runInternal with the records preloaded in RAM) on seeded synthetic inputs far larger than the
file-to-file goldens of make_golden.py -- C1 at its full size (1 M reads), a 5 M-read C2 slice, 2 M-record C3/C4 slices and a
C5 slice.  Stored per case: the duplicate bit of every record (packed, 1 bit each), the sha256 of the reference's whole
flag array and the sha256 of the generated records (guards generator drift).  Run in the build container only:

    python tests/golden/make_large_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle  # noqa: E402
from openge_b200 import synth  # noqa: E402

LARGE = {"C1": 1.0, "C2": 0.1, "C3": 0.2, "C4": 0.1, "C5": 0.0025}


def main():
    assert oracle.ref_available(), "reference not built"
    out = {}
    for name, scale in LARGE.items():
        bam = synth.make(name, scale)
        f = oracle.ref_flags_mem(bam)
        out[name + "_scale"] = np.float64(scale)
        out[name + "_n"] = np.int64(bam.n)
        out[name + "_dupbits"] = np.packbits((f & 0x400) != 0)
        out[name + "_flags_sha256"] = np.array(hashlib.sha256(f.tobytes()).hexdigest())
        out[name + "_records_sha256"] = np.array(hashlib.sha256(bam.records.tobytes()).hexdigest())
        print("%s scale %g: n=%d dup=%d" % (name, scale, bam.n, int(((f & 0x400) != 0).sum())), flush=True)
    np.savez_compressed(os.path.join(HERE, "large_pins.npz"), **out)


if __name__ == "__main__":
    main()
