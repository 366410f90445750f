"""Regenerates tests/golden/flagstats.npz with the COMPILED REFERENCE's own Statistics module
(algorithms/statistics.cpp) chained behind its MarkDuplicates: `oge_ref_dedup --nosplit -v --stats`
(openge_b200/host/refcli/ref_driver.cpp).  Run in the build container only:

    python tests/golden/make_flagstats_golden.py

One row of 13 numbers per case (oracle.FLAGSTAT_FIELDS order): the golden dedup cases and the
sortedness fixtures of tests/fixtures.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import fixtures  # noqa: E402
import oracle  # noqa: E402
from conftest import GOLDEN_CASES, load_golden  # noqa: E402


def main():
    assert oracle.ref_available(), "reference not built"
    out = {}
    for case in GOLDEN_CASES:
        bam, _ = load_golden(case)
        r = oracle.ref_stats(bam)
        out[case] = np.array([r[k] for k in oracle.FLAGSTAT_FIELDS], dtype=np.uint64)
        print(case, r)
    for name, bam in fixtures.sortedness_cases().items():
        r = oracle.ref_stats(bam)
        out["sortedness_" + name] = np.array([r[k] for k in oracle.FLAGSTAT_FIELDS], dtype=np.uint64)
        print(name, "sorted =", r["sorted"], "reads =", r["reads"])
    np.savez_compressed(os.path.join(HERE, "flagstats.npz"), **out)


if __name__ == "__main__":
    main()
