"""CPU suite: the oracle (oracle/markdup_oracle.c) against the reference's golden vectors.

The golden .npz files hold flag words produced by the compiled reference itself
(tests/golden/make_golden.py); where /root/reference is present the live reference
binary is compared as well.
"""
import hashlib

import numpy as np
import pytest

import fixtures
import oracle
from conftest import GOLDEN_CASES, LARGE_PIN_CASES, load_golden, load_large_pin


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_reference_golden(case):
    bam, g = load_golden(case)
    assert hashlib.sha256(bam.records.tobytes()).hexdigest() == str(g["sha256"]), "input drifted from golden"
    flags = oracle.markdup(bam.records, bam.offsets, bam.text)
    assert np.array_equal(flags, g["flags_nosplit_v"])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_compat_quiet_matches_reference(case):
    bam, g = load_golden(case)
    flags = oracle.markdup(bam.records, bam.offsets, bam.text, compat_quiet=True)
    assert np.array_equal(flags, g["flags_quiet"])


@pytest.mark.parametrize("case", ["a3_fixture1", "synth_C3", "synth_C5"])
def test_oracle_compat_split_matches_reference(case):
    bam, g = load_golden(case)
    # -t 4 -> min(12, 4/2) = 2 chains (command_dedup.cpp:46)
    flags = oracle.markdup_split(bam, 2)
    assert np.array_equal(flags, g["flags_split_t4_v"])


def test_survey_a3_expected_values():
    b1, exp1 = fixtures.fixture1()
    f1 = oracle.markdup(b1.records, b1.offsets, b1.text)
    assert np.array_equal(((f1 & 0x400) != 0).astype(np.uint8), exp1)
    b2, exp2 = fixtures.fixture2()
    f2 = oracle.markdup(b2.records, b2.offsets, b2.text)
    assert np.array_equal(f2, exp2)


def test_remove_duplicates_matches_reference():
    bam, g = load_golden("synth_C3")
    flags = oracle.markdup(bam.records, bam.offsets, bam.text)
    keep = (flags & 0x400) == 0          # mark_duplicates.cpp:456: drops anything flagged after the rewrite
    assert int(keep.sum()) == int(g["removed_n"])
    rec = bam.records.copy()
    off = bam.offsets[:-1].astype(np.int64)
    rec[off + 18] = (flags & 0xFF).astype(np.uint8)
    rec[off + 19] = (flags >> 8).astype(np.uint8)
    sizes = np.diff(bam.offsets.astype(np.int64))
    mask = np.repeat(keep, sizes)
    assert hashlib.sha256(rec[mask].tobytes()).hexdigest() == str(g["removed_sha256"])


def test_oracle_ends_fields():
    b1, _ = fixtures.fixture1()
    _, ends, stats = oracle.markdup(b1.records, b1.offsets, b1.text, want_ends=True)
    # G_softclip 5S95M at 1-based 8005 -> unclipped start 0-based 7999, same as F_noclip
    assert ends["coord"][11] == 7999 and ends["coord"][12] == 7999
    # reverse read 100M at 1-based 1300 -> unclipped end 0-based 1398
    assert ends["coord"][4] == 1398 and ends["orientation"][4] == 2
    assert ends["score"][0] == 4000 and ends["lib"][3] != ends["lib"][0]
    assert stats[0] == 25 and stats[1] == 9      # 25 frag entries (2 unmapped), 9 pairs


@pytest.mark.skipif(not oracle.ref_available(), reason="compiled reference not present")
@pytest.mark.parametrize("name,scale,seed", [("C1", 0.01, 11), ("C3", 0.004, 12), ("C4", 0.002, 13)])
def test_oracle_vs_live_reference(name, scale, seed):
    from openge_b200 import synth
    bam = synth.make(name, scale, seed=seed)
    try:
        ref = oracle.ref_dedup(bam, timeout=30)
    except oracle.RefHang as e:      # the reference's own pipeline race; the golden files pin the same thing
        pytest.skip(str(e))
    flags = oracle.markdup(bam.records, bam.offsets, bam.text)
    assert np.array_equal(flags, ref.flags())


def test_empty_input():
    flags = oracle.markdup(np.zeros(0, np.uint8), np.zeros(1, np.uint64), "@HD\tVN:1.4\tSO:coordinate\n")
    assert len(flags) == 0


# ---- flag statistics: the oracle restatement of Statistics::runInternal against the
# numbers printed by the compiled reference's own Statistics module (tests/golden/make_flagstats_golden.py)
def _flagstats_golden():
    import os
    from conftest import GOLDEN
    return dict(np.load(os.path.join(GOLDEN, "flagstats.npz")))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_flagstats_match_reference_statistics(case):
    bam, g = load_golden(case)
    want = _flagstats_golden()[case]
    got = oracle.flagstats(bam.records, bam.offsets, g["flags_nosplit_v"])
    assert [got[k] for k in oracle.FLAGSTAT_FIELDS] == [int(x) for x in want]


def test_oracle_sorted_verdict_matches_reference_statistics():
    gold = _flagstats_golden()
    for name, bam in fixtures.sortedness_cases().items():
        flags = oracle.markdup(bam.records, bam.offsets, bam.text)
        got = oracle.flagstats(bam.records, bam.offsets, flags)
        assert [got[k] for k in oracle.FLAGSTAT_FIELDS] == [int(x) for x in gold["sortedness_" + name]], name


# ---- coordinate order: the oracle restatement of ReadSorter + Sort::ByPosition against the compiled
# reference's own sorter (tests/golden/make_sort_golden.py).  Only the positions whose order the reference defines are compared
# in order (its comparator ends on object addresses); all records are compared as a multiset.
@pytest.mark.parametrize("name,scale,seed", [("C3", 0.01, 5), ("C4", 0.003, 6), ("C1", 0.02, 7), ("C3", 0.004, 8)])
def test_oracle_coordinate_order_matches_reference_sorter(name, scale, seed):
    import os
    import sys
    from conftest import GOLDEN
    from openge_b200 import synth
    sys.path.insert(0, GOLDEN)
    from make_sort_golden import digests
    gold = dict(np.load(os.path.join(GOLDEN, "sort_order.npz")))
    bam = fixtures.shuffled(synth.make(name, scale, seed=seed), seed)
    perm, tied = oracle.coordinate_order(bam.records, bam.offsets)
    assert sorted(perm.tolist()) == list(range(bam.n))
    d, n_def, ms = digests(bam.records, bam.offsets, perm, tied)
    key = "%s_%g_%d" % (name, scale, seed)
    assert n_def == int(gold[key + "_n_defined"]) and ms == str(gold[key + "_multiset"])
    assert d == str(gold[key + "_defined"])
    # feeding the sorted stream to MarkDuplicates is `openge mergesort -M`: the dedup oracle runs on the oracle's order
    o = bam.offsets.astype(np.int64)
    recs = [bam.records[o[i]:o[i + 1]].tobytes() for i in perm]
    from openge_b200 import bamio
    r, off = bamio.concat_records(recs)
    flags = oracle.markdup(r, off, bam.text)
    want = gold[key + "_dedup_flags"]
    assert np.array_equal(gold[key + "_tied"], tied)
    assert np.array_equal(flags[~tied], want[~tied])
    if not tied.any():
        assert np.array_equal(flags, want)


@pytest.mark.parametrize("name", LARGE_PIN_CASES)
def test_oracle_matches_reference_at_millions_of_records(name):
    """The restatement against the compiled reference's own flags (`--mem -v --flags`) on 1-5 M-record inputs:
    C1 at its full size, a 5 M-read C2 slice, 2 M-record C3/C4/C5 slices (tests/golden/large_pins.npz)."""
    import hashlib
    bam, dup, sha = load_large_pin(name)
    got = oracle.markdup(bam.records, bam.offsets, bam.text)
    assert np.array_equal((got & 0x400) != 0, dup)
    assert hashlib.sha256(got.tobytes()).hexdigest() == sha
