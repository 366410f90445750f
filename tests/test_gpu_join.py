"""GPU suite (-m gpu): the mate join in both of its forms -- the end-build fused with the in-CTA join plus the global
join over its leftovers and the check pass (default), and the separate whole-file hash join (debug_legacy_join, the form
the range-sharded path runs) -- on inputs built to break it: names seen one to eight times anywhere in the file, mates
far apart, names that differ only behind the 29 bytes of a name tag, read groups the header does not list, stripped names.
The oracle pairs in file order like the reference's map (util/picard_structures.h:87-96)."""
import numpy as np
import pytest

import fixtures
import oracle
from openge_b200 import dedup, synth

pytestmark = pytest.mark.gpu


def gpu_flags(bam, **kw):
    with dedup.context_for(bam, **kw) as ctx:
        ctx.push(bam.records, bam.offsets)
        ctx.run()
        return ctx.flags(), ctx.stats()


@pytest.mark.parametrize("legacy", [False, True])
@pytest.mark.parametrize("n,seed,pool_div,n_pos", [(3000, 1, 2.5, 40), (60000, 2, 2.5, 400), (60000, 3, 1.2, 3000),
                                                  (150000, 4, 4.0, 200), (150000, 5, 2.0, 100000)])
def test_name_soup_vs_oracle(legacy, n, seed, pool_div, n_pos):
    bam = fixtures.name_soup(n=n, seed=seed, pool_div=pool_div, n_pos=n_pos)
    want, _, ostats = oracle.markdup(bam.records, bam.offsets, bam.text, want_ends=True)
    got, st = gpu_flags(bam, legacy_join=legacy)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, "%d of %d flag words differ, first at record %d" % (len(bad), bam.n, int(bad[0]))
    assert st["n_pair_entries"] == int(ostats[1])
    assert st["n_complex_names"] > 0


@pytest.mark.parametrize("legacy", [False, True])
@pytest.mark.parametrize("n", [2, 3, 257, 3000, 100001])
def test_one_name_for_every_record(legacy, n):
    """Stripped read names: every record toggles the same map key, so sightings pair (1,2), (3,4), ... in file order.  At 100 001
    records the whole file is ONE segment of the exact path: it must not be replayed quadratically (a CTA establishes that all
    keys are equal and emits the pairs in parallel)."""
    bam = fixtures.one_name(n=n)
    want = oracle.markdup(bam.records, bam.offsets, bam.text)
    got, st = gpu_flags(bam, legacy_join=legacy)
    assert np.array_equal(got, want)
    assert st["n_pair_entries"] == n // 2


@pytest.mark.parametrize("name,scale,seed", [("C1", 0.3, 11), ("C2", 0.02, 12), ("C3", 0.1, 13), ("C4", 0.05, 14), ("C5", 0.001, 15)])
def test_fused_and_legacy_join_agree(name, scale, seed):
    bam = synth.make(name, scale, seed=seed)
    a, sa = gpu_flags(bam)
    b, sb = gpu_flags(bam, legacy_join=True)
    assert np.array_equal(a, b)
    assert sa["n_pair_entries"] == sb["n_pair_entries"] and sa["n_duplicates"] == sb["n_duplicates"]


def test_all_singleton_names_do_not_overfill_the_table():
    """Every read claims a mate that is not in the file (a region extract): as many distinct keys as records."""
    bam = fixtures.name_soup(n=40000, seed=9, pool_div=0.01, n_pos=2000, long_names=False)
    want = oracle.markdup(bam.records, bam.offsets, bam.text)
    for legacy in (False, True):
        got, _ = gpu_flags(bam, legacy_join=legacy)
        assert np.array_equal(got, want)
