"""GPU suite: the drop-in `openge dedup` binary -- the reference's own pipeline classes with this
repo's MarkDuplicates (openge_b200/host/mark_duplicates_gpu.cpp) linked in place of the
reference's -- against the unmodified reference binary and the golden flags, file to file."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import load_golden
from openge_b200 import _build, bamio, synth

pytestmark = pytest.mark.gpu


def run_dedup(exe, bam, *extra, timeout=120):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        inp, out = os.path.join(d, "in.bam"), os.path.join(d, "out.rawbam")
        bamio.write_bam(inp, bam)      # BGZF in, as `openge dedup in.bam` gets it
        cmd = [exe, "-T", d, "-F", "rawbam", "--nosplit", "-v", *extra, inp, out]
        for _ in range(4):      # the reference's pipeline now and then never terminates
            try:
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
            except subprocess.TimeoutExpired:
                continue
            assert r.returncode == 0, r.stderr.decode()[-2000:]
            return bamio.read_bam(out), r.stderr.decode()
        pytest.skip("pipeline did not terminate in %d s (4 attempts)" % timeout)


@pytest.fixture(scope="module")
def gpu_exe():
    exe = _build.ensure_host()
    if not exe:
        pytest.skip("openge_b200/host/_build/oge_dedup_gpu was not built (needs the reference sources at build time)")
    return exe


@pytest.mark.parametrize("case", ["a3_fixture1", "a3_fixture2", "edge_cases", "synth_C3"])
def test_dropin_binary_matches_golden(gpu_exe, case):
    bam, g = load_golden(case)
    out, log = run_dedup(gpu_exe, bam)
    assert "on the GPU" in log
    assert np.array_equal(out.flags(), g["flags_nosplit_v"])
    assert np.array_equal(out.offsets, bam.offsets)
    # everything but the flag word is untouched -- and the bin, which the reference's writer recomputes
    # for every record (util/bam_serializer.h:108-126)
    keep = np.ones(len(bam.records), dtype=bool)
    pos = bam.offsets[:-1].astype(np.int64)
    for o in (14, 15, 18, 19):
        keep[pos + o] = False
    assert np.array_equal(out.records[keep], bam.records[keep])


def test_dropin_binary_equals_reference_binary_file_to_file(gpu_exe):
    ref = _build.REF_BIN if os.path.exists(_build.REF_BIN) else None
    if not ref:
        pytest.skip("oracle/_ref/oge_ref_dedup not built")
    bam = synth.make("C3", 0.02, seed=7)
    a, _ = run_dedup(gpu_exe, bam)
    b, _ = run_dedup(ref, bam)
    assert a.text == b.text and a.refs == b.refs
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.records, b.records)


def test_dropin_binary_remove_duplicates(gpu_exe):
    bam, g = load_golden("synth_C3")
    out, _ = run_dedup(gpu_exe, bam, "-r")
    assert out.n == int(g["removed_n"])
    import hashlib
    assert hashlib.sha256(out.records.tobytes()).hexdigest() == str(g["removed_sha256"])


def test_dropin_class_behind_the_reference_sorter_is_mergesort_M(gpu_exe):
    """`openge mergesort -M` (commands/command_mergesort.cpp:70-113), the other caller of MarkDuplicates: the reference's own
    ReadSorter in front of this repo's drop-in class.  Flags compared with the compiled reference's run of the same chain
    (tests/golden/sort_order.npz) wherever the reference defines the record order."""
    import fixtures
    from conftest import GOLDEN
    gold = dict(np.load(os.path.join(GOLDEN, "sort_order.npz")))
    for name, scale, seed, per_tempfile in (("C4", 0.003, 6, 7000), ("C3", 0.01, 5, 20000)):
        bam = fixtures.shuffled(synth.make(name, scale, seed=seed), seed)
        key = "%s_%g_%d" % (name, scale, seed)
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        with tempfile.TemporaryDirectory(dir=base) as d:
            inp, out = os.path.join(d, "in.rawbam"), os.path.join(d, "out.rawbam")
            bamio.write_bam(inp, bam, raw=True)
            cmd = [gpu_exe, "-T", d, "--sort", "-v", "-F", "rawbam", "-n", str(per_tempfile), inp, out]
            got = None
            for _ in range(4):
                try:
                    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
                except subprocess.TimeoutExpired:
                    continue
                assert r.returncode == 0, r.stderr.decode()[-2000:]
                assert "on the GPU" in r.stderr.decode()
                got = bamio.read_bam(out)
                break
            if got is None:
                pytest.skip("pipeline did not terminate")
        tied = gold[key + "_tied"]
        assert got.n == bam.n
        assert np.array_equal(got.flags()[~tied], gold[key + "_dedup_flags"][~tied])


@pytest.fixture(scope="module")
def gpu_sort_exe():
    _build.ensure_host()
    if not os.path.exists(_build.HOST_SORT_BIN):
        pytest.skip("openge_b200/host/_build/oge_mergesort_gpu was not built (needs the reference sources at build time)")
    return _build.HOST_SORT_BIN


def test_dropin_read_sorter_is_mergesort(gpu_sort_exe):
    """`openge mergesort [-M]` (commands/command_mergesort.cpp:68-100) with BOTH stages replaced: this repo's ReadSorter
    (read_sorter_gpu.cpp -> oge_gpu_dedup_sort) and MarkDuplicates inside the reference's own pipeline.  The sorter's output
    must be the oracle's restatement of ReadSorter + Sort::ByPosition record for record (which is pinned to the compiled
    reference's sorter wherever the reference defines the order), the chain's flags the compiled reference's."""
    import fixtures
    import oracle
    from conftest import GOLDEN
    gold = dict(np.load(os.path.join(GOLDEN, "sort_order.npz")))
    for name, scale, seed in (("C4", 0.003, 6), ("C3", 0.01, 5)):
        bam = fixtures.shuffled(synth.make(name, scale, seed=seed), seed)
        key = "%s_%g_%d" % (name, scale, seed)
        want_perm, _ = oracle.coordinate_order(bam.records, bam.offsets)
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        with tempfile.TemporaryDirectory(dir=base) as d:
            inp = os.path.join(d, "in.rawbam")
            bamio.write_bam(inp, bam, raw=True)
            outs = {}
            for mode, extra in (("sort", ["--nodedup"]), ("chain", [])):
                out = os.path.join(d, mode + ".rawbam")
                cmd = [gpu_sort_exe, "-T", d, "--sort", "-v", "-F", "rawbam"] + extra + [inp, out]
                got = None
                for _ in range(4):
                    try:
                        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
                    except subprocess.TimeoutExpired:
                        continue
                    assert r.returncode == 0, r.stderr.decode()[-2000:]
                    assert "by coordinate on the GPU" in r.stderr.decode()
                    got = bamio.read_bam(out)
                    break
                if got is None:
                    pytest.skip("pipeline did not terminate")
                outs[mode] = got
        # the sorter alone: the records in the oracle's order, byte for byte (the writer recomputes the bin: bytes 14-15)
        s = outs["sort"]
        assert s.n == bam.n and "SO:coordinate" in s.text
        sizes = np.diff(bam.offsets.astype(np.int64))
        assert np.array_equal(np.diff(s.offsets.astype(np.int64)), sizes[want_perm])
        starts = bam.offsets[:-1].astype(np.int64)[want_perm]
        idx = np.concatenate([np.arange(a, a + l) for a, l in zip(starts, sizes[want_perm])])
        keep = np.ones(len(s.records), dtype=bool)
        pos = s.offsets[:-1].astype(np.int64)
        for o in (14, 15):
            keep[pos + o] = False
        assert np.array_equal(s.records[keep], bam.records[idx][keep])
        # the chain: the compiled reference's flags wherever its order is defined
        tied = gold[key + "_tied"]
        assert outs["chain"].n == bam.n
        assert np.array_equal(outs["chain"].flags()[~tied], gold[key + "_dedup_flags"][~tied])
