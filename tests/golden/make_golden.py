"""Regenerates tests/golden/*.npz by running the COMPILED REFERENCE (oracle/_ref/oge_ref_dedup,
built in place from /root/reference) -- run in the build container only:

    python tests/golden/make_golden.py

Each .npz holds the reference's output flag words (canonical `--nosplit -v`, plus the compat
modes where noted).  Inputs are either stored (fixtures, 208.yhet) or regenerated from a seed
(synthetic configs; a sha256 of the generated records guards generator drift).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixtures  # noqa: E402
import oracle  # noqa: E402
from openge_b200 import bamio, synth  # noqa: E402

SYNTH_SMALL = {"C1": 0.04, "C2": 0.002, "C3": 0.01, "C4": 0.005, "C5": 0.0001}


def save_case(name, bam, store_input, **extra):
    ref = oracle.ref_dedup(bam, nosplit=True, verbose=True)
    assert np.array_equal(ref.offsets, bam.offsets)
    out = dict(flags_nosplit_v=ref.flags(), sha256=hashlib.sha256(bam.records.tobytes()).hexdigest(), **extra)
    out["flags_quiet"] = oracle.ref_dedup(bam, nosplit=True, verbose=False).flags()
    out["flags_split_t4_v"] = oracle.ref_dedup(bam, nosplit=False, verbose=True, threads=4).flags()
    rem = oracle.ref_dedup(bam, nosplit=True, verbose=True, remove=True)
    out["removed_n"] = np.int64(rem.n)
    out["removed_sha256"] = hashlib.sha256(rem.records.tobytes()).hexdigest()
    if store_input:
        out.update(records=bam.records, offsets=bam.offsets, text=np.array(bam.text),
                   ref_names=np.array([r[0] for r in bam.refs]), ref_lens=np.array([r[1] for r in bam.refs]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    f = ref.flags()
    print("%-16s n=%d dup=%d quiet=%d split=%d removed_n=%d" % (
        name, bam.n, int(((f & 0x400) != 0).sum()), int(((out["flags_quiet"] & 0x400) != 0).sum()),
        int(((out["flags_split_t4_v"] & 0x400) != 0).sum()), rem.n))


def main():
    assert oracle.ref_available(), "reference not built"
    b1, exp1 = fixtures.fixture1()
    save_case("a3_fixture1", b1, True, expected_dup=exp1)
    b2, exp2 = fixtures.fixture2()
    save_case("a3_fixture2", b2, True, expected_flags=exp2)
    save_case("edge_cases", fixtures.edge_cases(), True)
    yhet = "/root/reference/openge/test/data/208.yhet.bam"
    if os.path.exists(yhet):
        save_case("yhet208", bamio.read_bam(yhet), True)
    for name, scale in SYNTH_SMALL.items():
        save_case("synth_%s" % name, synth.make(name, scale), False, scale=np.float64(scale))


if __name__ == "__main__":
    main()
