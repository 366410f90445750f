"""GPU suite (-m gpu): f3, the coordinate sort on the device in front of the duplicate marking (`openge mergesort -M`).

oge_gpu_dedup_sort against the oracle's restatement of ReadSorter + Sort::ByPosition (oracle.coordinate_order, itself
pinned to the compiled reference's sorter: tests/golden/sort_order.npz) -- the permutation must be IDENTICAL, ties
and the unplaced tail in input order included -- and against the reference's own sorter output wherever the reference
defines the order (its comparator ends on object addresses).  Then the whole chain sort -> dedup against the oracle
and the reference's `mergesort -M` flags."""
import os
import sys

import numpy as np
import pytest

import fixtures
import oracle
from conftest import GOLDEN
from openge_b200 import bamio, dedup, synth

pytestmark = pytest.mark.gpu


def gpu_sort(bam, run=False, **kw):
    with dedup.context_for(bam, **kw) as ctx:
        ctx.push(bam.records, bam.offsets)
        ctx.sort()
        perm = ctx.sort_order()
        rec, off = ctx_records(ctx)
        flags = None
        if run:
            ctx.run()
            flags = ctx.flags()
        return perm, rec, off, flags, ctx.sort_stats()


def ctx_records(ctx):
    # pull() needs a completed run; the sorted records themselves are read through a run-less path: run on them is cheap
    ctx.run()
    rec, off = ctx.pull()
    return rec, off


def reorder(bam, perm):
    o = bam.offsets.astype(np.int64)
    recs = [bam.records[o[i]:o[i + 1]].tobytes() for i in perm]
    return bamio.concat_records(recs)


CASES = [("C3", 0.01, 5), ("C4", 0.003, 6), ("C1", 0.02, 7), ("C3", 0.004, 8)]


@pytest.mark.parametrize("name,scale,seed", CASES)
def test_sort_matches_oracle_and_reference_sorter(name, scale, seed):
    sys.path.insert(0, GOLDEN)
    from make_sort_golden import digests
    gold = dict(np.load(os.path.join(GOLDEN, "sort_order.npz")))
    bam = fixtures.shuffled(synth.make(name, scale, seed=seed), seed)
    want_perm, tied = oracle.coordinate_order(bam.records, bam.offsets)
    perm, rec, off, flags, st = gpu_sort(bam, run=True)
    assert np.array_equal(perm, want_perm)
    # the records really moved: byte-identical to the input records taken in that order, except the duplicate bit the run set
    r_want, o_want = reorder(bam, want_perm)
    assert np.array_equal(off, o_want)
    diff = np.nonzero(rec != r_want)[0]
    assert all(int(d - o_want[np.searchsorted(o_want, d, side="right") - 1]) == 19 for d in diff)      # high byte of the flag word
    # against the reference's own sorter, where it defines the order
    d, n_def, ms = digests(bam.records, bam.offsets, perm, tied)
    key = "%s_%g_%d" % (name, scale, seed)
    assert n_def == int(gold[key + "_n_defined"]) and ms == str(gold[key + "_multiset"]) and d == str(gold[key + "_defined"])
    # sort -> dedup = `openge mergesort -M`
    assert np.array_equal(flags, oracle.markdup(r_want, o_want, bam.text))
    want = gold[key + "_dedup_flags"]
    assert np.array_equal(flags[~tied], want[~tied])


@pytest.mark.parametrize("seed", [1, 2])
def test_sort_long_tie_runs_and_long_names(seed):
    """Few distinct positions (runs of hundreds of records tying on refID, position and strand), names that differ only
    behind a 34-byte common prefix, exact copies, unplaced records in between."""
    bam = fixtures.shuffled(fixtures.name_soup(n=40000, seed=seed, pool_div=2.0, n_pos=60), seed)
    want_perm, _ = oracle.coordinate_order(bam.records, bam.offsets)
    perm, _, _, flags, st = gpu_sort(bam, run=True)
    assert np.array_equal(perm, want_perm)
    assert st["tied_records"] > 30000 and st["refinement_rounds"] >= 5
    r, o = reorder(bam, want_perm)
    assert np.array_equal(flags, oracle.markdup(r, o, bam.text))


def test_sort_unplaced_tail_keeps_input_order():
    base = synth.make("C3", 0.005, seed=31)      # unmapped mates and secondary records in the mix
    o = base.offsets.astype(np.int64)
    recs = []
    for i in range(base.n):
        r = bytearray(base.records[o[i]:o[i + 1]].tobytes())
        if i % 7 == 0:
            r[4:8] = (-1).to_bytes(4, "little", signed=True)      # no reference: Sort::ByPosition treats these as equivalent
        recs.append(bytes(r))
    rr, oo = bamio.concat_records(recs)
    bam = fixtures.shuffled(bamio.BamFile(text=base.text, refs=list(base.refs), records=rr, offsets=oo), 3)
    want_perm, _ = oracle.coordinate_order(bam.records, bam.offsets)
    perm, _, _, _, _ = gpu_sort(bam)
    assert np.array_equal(perm, want_perm)


def test_sort_of_sorted_input_is_the_identity_and_small_cases():
    bam = synth.make("C1", 0.02, seed=5)
    want_perm, _ = oracle.coordinate_order(bam.records, bam.offsets)
    perm, _, _, _, _ = gpu_sort(bam)
    assert np.array_equal(perm, want_perm)
    for n in (1, 2, 3):
        o = bam.offsets.astype(np.int64)
        r, off = bamio.concat_records([bam.records[o[i]:o[i + 1]].tobytes() for i in range(n)][::-1])
        small = bamio.BamFile(text=bam.text, refs=list(bam.refs), records=r, offsets=off)
        w, _ = oracle.coordinate_order(small.records, small.offsets)
        p, _, _, _, _ = gpu_sort(small)
        assert np.array_equal(p, w)


def test_sort_at_scale_then_dedup():
    """2 M shuffled C2 reads: permutation against the oracle, sort -> dedup against the oracle's chain."""
    bam = synth.make("C2", 0.04, seed=77)
    rng = np.random.default_rng(7)
    p0 = rng.permutation(bam.n)
    o = bam.offsets.astype(np.int64)
    sizes = np.diff(o)
    new_off = np.zeros(bam.n + 1, dtype=np.uint64)
    new_off[1:] = np.cumsum(sizes[p0])
    rec = np.empty(int(new_off[-1]), dtype=np.uint8)
    for k, i in enumerate(p0):      # python loop over 2 M records: a few seconds
        rec[int(new_off[k]):int(new_off[k + 1])] = bam.records[o[i]:o[i + 1]]
    sh = bamio.BamFile(text=bam.text, refs=list(bam.refs), records=rec, offsets=new_off)
    want_perm, _ = oracle.coordinate_order(sh.records, sh.offsets)
    with dedup.context_for(sh) as ctx:
        ctx.push(sh.records, sh.offsets)
        ctx.sort()
        perm = ctx.sort_order()
        ctx.run()
        flags = ctx.flags()
    assert np.array_equal(perm, want_perm)
    r, off = reorder(sh, want_perm)
    assert np.array_equal(flags, oracle.markdup(r, off, sh.text))


@pytest.mark.parametrize("cpu_inflate", [False, True])
def test_fused_binary_sort_then_dedup_is_mergesort_M(cpu_inflate):
    """`oge_dedup_fused --sort` = `openge mergesort -M`: file in any order -> coordinate-sorted, duplicate-marked file.
    Records and flags against the oracle's chain, the reference's `mergesort -M` flags where its order is defined,
    @HD SO:coordinate in the stored header (read_sorter.cpp:256-258)."""
    import subprocess
    import tempfile
    from openge_b200 import _build, bamhost
    exe = _build.ensure_fused()
    if not exe or not os.path.exists(exe):
        pytest.skip("openge_b200/host/_build/oge_dedup_fused was not built")
    gold = dict(np.load(os.path.join(GOLDEN, "sort_order.npz")))
    name, scale, seed = "C3", 0.01, 5
    bam = fixtures.shuffled(synth.make(name, scale, seed=seed), seed)
    want_perm, tied = oracle.coordinate_order(bam.records, bam.offsets)
    r_want, o_want = reorder(bam, want_perm)
    f_want = oracle.markdup(r_want, o_want, bam.text)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        inp, out = os.path.join(d, "in.bam"), os.path.join(d, "out.rawbam")
        bamio.write_bam(inp, bam)
        cmd = [exe, inp, "-o", out, "-F", "rawbam", "-v", "--nopg", "--sort"] + (["--cpu-inflate"] if cpu_inflate else [])
        r = subprocess.run(cmd, capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        assert b"by coordinate on the GPU" in r.stderr
        got = bamio.read_bam(out)
    assert "SO:coordinate" in got.text and got.text == bamhost.header_render(bam.text.replace("SO:unsorted", "SO:coordinate"))
    assert np.array_equal(got.offsets, o_want)
    assert np.array_equal(got.flags(), f_want)
    # every byte but the flag word's duplicate bit (and the bin the writer recomputes) is the input record's
    a, b = got.records.copy(), r_want.copy()
    for arr in (a, b):
        starts = o_want[:-1].astype(np.int64)
        arr[starts + 14] = 0; arr[starts + 15] = 0; arr[starts + 19] &= 0xFB
    assert np.array_equal(a, b)
    key = "%s_%g_%d" % (name, scale, seed)
    assert np.array_equal(got.flags()[~tied], gold[key + "_dedup_flags"][~tied])
