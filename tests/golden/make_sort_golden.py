"""Regenerates tests/golden/sort_order.npz with the COMPILED REFERENCE's ReadSorter chain (`openge mergesort`,
algorithms/read_sorter.cpp; openge_b200/host/refcli/ref_driver.cpp --sort).  Run in the build container only:

    python tests/golden/make_sort_golden.py

Inputs are seeded shuffles of the synthetic configs (tests/fixtures.py: shuffled).  The reference's order is defined only
outside groups that tie on (refID, pos, strand, name, flag) and outside the unplaced tail (its comparator ends on object
addresses), so the golden holds, per case: a digest of the records at the DEFINED output positions in order, the number of
such positions, and an order-independent digest of all records.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import fixtures  # noqa: E402
import oracle  # noqa: E402
from openge_b200 import synth  # noqa: E402

CASES = [("C3", 0.01, 5, 20000), ("C4", 0.003, 6, 7000), ("C1", 0.02, 7, 200000), ("C3", 0.004, 8, 1000)]


def digests(records, offsets, order, tied):
    """order[k] = record ordinal (into records/offsets) at output position k."""
    o = offsets.astype(np.int64)
    h_def = hashlib.sha256()
    per = []
    for k, i in enumerate(order):
        b = bytearray(records[o[i]:o[i + 1]].tobytes())
        b[14:16] = b"\0\0"      # the bin is recomputed by the reference's writer
        per.append(hashlib.sha256(bytes(b)).digest())
        if not tied[k]:
            h_def.update(bytes(b))
    return h_def.hexdigest(), int((~tied).sum()), hashlib.sha256(b"".join(sorted(per))).hexdigest()


def main():
    assert oracle.ref_available(), "reference not built"
    out = {}
    for name, scale, seed, per_tempfile in CASES:
        bam = fixtures.shuffled(synth.make(name, scale, seed=seed), seed)
        ref = oracle.ref_sort(bam, per_tempfile=per_tempfile)
        _, tied = oracle.coordinate_order(bam.records, bam.offsets)      # which positions are defined depends on the keys only
        d, n_def, ms = digests(ref.records, ref.offsets, np.arange(ref.n), tied)
        key = "%s_%g_%d" % (name, scale, seed)
        out[key + "_defined"] = np.array(d)
        out[key + "_n_defined"] = np.int64(n_def)
        out[key + "_multiset"] = np.array(ms)
        assert "SO:coordinate" in ref.text
        # `openge mergesort -M`: the same chain with MarkDuplicates behind the sorter; flag words in output order
        # (comparable position by position only where the order is defined)
        out[key + "_dedup_flags"] = oracle.ref_sort(bam, dedup=True, per_tempfile=per_tempfile).flags()
        out[key + "_tied"] = tied
        print(key, ref.n, "defined", n_def, d[:16], ms[:16])
    np.savez_compressed(os.path.join(HERE, "sort_order.npz"), **out)


if __name__ == "__main__":
    main()
