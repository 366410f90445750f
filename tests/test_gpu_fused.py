"""GPU suite: `openge dedup in.bam -o out.bam` as one fused path (oge_bam_load -> CUDA dedup -> oge_bam_store), through
the Python binding and through the native binary, against the output FILES of the compiled reference
(tests/golden/bamfile.npz holds their sha256) and the golden flags."""
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, load_golden
from openge_b200 import _build, bamhost, bamio, dedup, synth

pytestmark = pytest.mark.gpu

GOLD = dict(np.load(os.path.join(GOLDEN, "bamfile.npz")))


@pytest.fixture()
def tmp():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        yield d


@pytest.fixture(scope="module")
def fused_exe():
    exe = _build.ensure_fused()
    if not exe or not os.path.exists(exe):
        pytest.skip("openge_b200/host/_build/oge_dedup_fused was not built")
    return exe


@pytest.mark.parametrize("name,scale,seed,level,remove", [("C3", 0.01, 99, 6, False), ("C3", 0.01, 99, 1, False), ("C3", 0.01, 99, 6, True),
                                                          ("C1", 0.02, 5, 6, False), ("C4", 0.004, 6, 9, False)])
def test_dedup_file_is_byte_identical_to_the_reference_output(tmp, name, scale, seed, level, remove):
    bam = synth.make(name, scale, seed=seed)
    inp, out = os.path.join(tmp, "in.bam"), os.path.join(tmp, "out.bam")
    bamio.write_bam(inp, bam)
    st = bamhost.dedup_file(inp, out, remove_duplicates=remove, level=level)
    key = "%s_%g_%d_c%d%s" % (name, scale, seed, level, "_r" if remove else "")
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == str(GOLD[key])
    assert st["dedup"]["launches"] > 0 and st["flagstats"]["reads"] == bam.n


@pytest.mark.parametrize("case,level", [("yhet208", 6), ("edge_cases", 1)])
def test_dedup_file_on_real_data_is_byte_identical_to_the_reference_output(tmp, case, level):
    bam, _ = load_golden(case)
    inp, out = os.path.join(tmp, "in.bam"), os.path.join(tmp, "out.bam")
    bamio.write_bam(inp, bam)
    bamhost.dedup_file(inp, out, level=level)
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == str(GOLD["%s_c%d" % (case, level)])


@pytest.mark.parametrize("case", ["a3_fixture1", "a3_fixture2", "edge_cases", "yhet208", "synth_C2", "synth_C5"])
def test_fused_binary_matches_golden_flags(tmp, fused_exe, case):
    bam, g = load_golden(case)
    inp, out = os.path.join(tmp, "in.bam"), os.path.join(tmp, "out.rawbam")
    bamio.write_bam(inp, bam)
    r = subprocess.run([fused_exe, "dedup", inp, "-o", out, "-F", "rawbam", "-v", "--nopg", "--nosplit", "-T", tmp],
                       capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert b"on the GPU" in r.stderr
    got = bamio.read_bam(out)
    assert np.array_equal(got.flags(), g["flags_nosplit_v"])
    assert np.array_equal(got.offsets, bam.offsets)
    assert got.text == bamhost.header_render(bam.text)


def test_fused_binary_remove_stats_and_pg(tmp, fused_exe):
    bam, g = load_golden("synth_C3")
    inp, out = os.path.join(tmp, "in.bam"), os.path.join(tmp, "out.bam")
    bamio.write_bam(inp, bam)
    r = subprocess.run([fused_exe, inp, "-o", out, "-r", "--stats", "-c", "1", "-t", "4"], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    got = bamio.read_bam(out)
    assert got.n == int(g["removed_n"])
    assert hashlib.sha256(got.records.tobytes()).hexdigest() == str(g["removed_sha256"])
    assert got.text.endswith("@PG\tID:openge\tCL:openge %s -o %s -r --stats -c 1 -t 4 \tVN:0.3-b200\n" % (inp, out))
    # --stats: the reference's Statistics report, counted on the device (before the -r filter)
    want = dict(np.load(os.path.join(GOLDEN, "flagstats.npz")))["synth_C3"]
    lines = dict(l.split(":", 1) for l in r.stdout.decode().splitlines())
    assert int(lines["Total reads"].split()[0]) == int(want[0])
    assert int(lines["Duplicates"].split()[0]) == int(want[5])
    assert int(lines["Singletons"].split()[0]) == int(want[11])
    assert lines["Sorted"].split()[0] == "Yes"
    if oracle.ref_available():      # the text itself, character for character
        exe = _build.REF_BIN
        o2 = os.path.join(tmp, "ref.rawbam")
        for _ in range(4):
            try:
                rr = subprocess.run([exe, "-T", tmp, "--nosplit", "-v", "--stats", "-F", "rawbam", inp, o2], capture_output=True, timeout=120)
                break
            except subprocess.TimeoutExpired:
                rr = None
        if rr is not None and rr.returncode == 0:
            assert rr.stdout.decode() == r.stdout.decode()


def test_fused_binary_reports_errors_like_the_reference(tmp, fused_exe):
    p = os.path.join(tmp, "bad.bam")
    open(p, "wb").write(b"\x1f\x8b\x08\x04" + b"\0" * 60)
    r = subprocess.run([fused_exe, p, "-o", os.path.join(tmp, "o.bam")], capture_output=True, timeout=60)
    assert r.returncode != 0 and b"Aborting." in r.stderr


def test_fused_binary_over_several_gpus_writes_the_same_file(tmp):
    """`oge_dedup_fused --gpus N`: the records range-sharded over N GPUs of the node (one host thread and one context per GPU,
    the exchanges done by NCCL inside the library) must write the file of the single-GPU run, byte for byte -- cuts inside
    contigs, mates and duplicate sets straddling them."""
    n_dev = dedup.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs on the box")
    exe = _build.FUSED_BIN
    for name, scale, seed in (("C5", 0.0005, 4), ("C3", 0.02, 8), ("C4", 0.01, 3)):
        bam = synth.make(name, scale, seed=seed)
        inp = os.path.join(tmp, "in_%s.bam" % name)
        bamio.write_bam(inp, bam)
        outs = []
        for gpus in [1, 2] + ([n_dev] if n_dev > 2 else []):
            for extra in ([], ["-r"]):
                out = os.path.join(tmp, "o_%s_%d_%d.bam" % (name, gpus, len(extra)))
                r = subprocess.run([exe, "dedup", inp, "-o", out, "-v", "--nopg", "-c", "1", "--gpus", str(gpus)] + extra,
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
                assert r.returncode == 0, r.stderr.decode()[-2000:]
                assert ("Range-sharded over %d GPUs" % gpus in r.stderr.decode()) == (gpus > 1)
                outs.append((gpus, len(extra), hashlib.sha256(open(out, "rb").read()).hexdigest()))
        for remove in (0, 1):
            assert len({h for g, e, h in outs if e == remove}) == 1, outs
