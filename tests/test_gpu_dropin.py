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
