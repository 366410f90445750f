"""GPU suite (-m gpu), full sizes: the CUDA path through the C ABI against the oracle, record by record, at the sizes
BASELINE.json states for the single-GPU configs -- C1 1 M, C2 50 M, C3 10 M records, C4 20 M reads -- and against the
compiled reference's own flags on the 1-5 M-record pins (tests/golden/large_pins.npz).  Index fields, table load,
32-bit counters and the exact path all sit in other regimes at these sizes than in the small cases of test_gpu_parity.py.
The oracle runs about 1.4 s per million reads on one host core, so the C2 case takes about two minutes."""
import hashlib

import numpy as np
import pytest

import oracle
from conftest import LARGE_PIN_CASES, load_large_pin
from openge_b200 import dedup, synth

pytestmark = pytest.mark.gpu


def gpu_flags(bam, **kw):
    with dedup.context_for(bam, **kw) as ctx:
        ctx.push(bam.records, bam.offsets)
        ctx.run()
        return ctx.flags(), ctx.stats()


@pytest.mark.parametrize("name", LARGE_PIN_CASES)
def test_flags_match_reference_at_millions_of_records(name):
    bam, dup, sha = load_large_pin(name)
    got, _ = gpu_flags(bam)
    assert np.array_equal((got & 0x400) != 0, dup)
    assert hashlib.sha256(got.tobytes()).hexdigest() == sha


@pytest.mark.parametrize("name", ["C1", "C3", "C4", "C2"])
def test_full_size_config_vs_oracle(name):
    bam = synth.make(name, 1.0)
    want, _, ostats = oracle.markdup(bam.records, bam.offsets, bam.text, want_ends=True)
    got, st = gpu_flags(bam)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, "%s: %d of %d flag words differ, first at record %d" % (name, len(bad), bam.n, int(bad[0]))
    assert st["n_frag_entries"] == int(ostats[0]) and st["n_pair_entries"] == int(ostats[1])
    assert st["n_duplicates"] == int(((want & 0x400) != 0).sum() - ((want & 0x500) == 0x500).sum())
