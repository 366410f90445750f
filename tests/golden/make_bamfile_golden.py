"""Regenerates tests/golden/bamfile.npz: sha256 of the output FILES the COMPILED REFERENCE writes
(`oge_ref_dedup --nosplit -v -c <level> in.bam out.bam`, i.e. FileReader -> MarkDuplicates -> FileWriter with its own
BgzfOutputStream) and the header texts it renders for awkward headers.  Run in the build container only:

    python tests/golden/make_bamfile_golden.py

tests/test_bamhost.py rebuilds each input from its seed, runs the host layer (load -> oracle flags -> apply -> store)
and compares file hashes: byte-identical output, compressed blocks included (same zlib in this image and on the GPU box).
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import fixtures  # noqa: E402
import oracle  # noqa: E402
from openge_b200 import _build, bamio, synth  # noqa: E402

CASES = [("C3", 0.01, 99, 6, False), ("C3", 0.01, 99, 1, False), ("C3", 0.01, 99, 0, False), ("C3", 0.01, 99, 6, True),
         ("C1", 0.02, 5, 6, False), ("C4", 0.004, 6, 9, False)]


def ref_file(bam, level, remove, d, fmt=None):
    inp, out = os.path.join(d, "in.bam"), os.path.join(d, "out.bam")
    bamio.write_bam(inp, bam)
    cmd = [_build.REF_BIN, "-T", d, "--nosplit", "-v", "-c", str(level)] + (["-r"] if remove else []) + (["-F", fmt] if fmt else []) + [inp, out]
    for _ in range(4):
        try:
            r = subprocess.run(cmd, capture_output=True, timeout=120)
        except subprocess.TimeoutExpired:
            continue
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        return open(out, "rb").read()
    raise RuntimeError("reference did not terminate")


def main():
    assert oracle.ref_available(), "reference not built"
    out = {}
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        for name, scale, seed, level, remove in CASES:
            bam = synth.make(name, scale, seed=seed)
            data = ref_file(bam, level, remove, d)
            key = "%s_%g_%d_c%d%s" % (name, scale, seed, level, "_r" if remove else "")
            out[key] = np.array(hashlib.sha256(data).hexdigest())
            print(key, len(data), out[key])
        # real data: the reference's own test file (test/data/208.yhet.bam, stored in yhet208.npz), whose header is not canonical
        from conftest import load_golden
        for case in ("yhet208", "edge_cases"):
            bam, _ = load_golden(case)
            for level in (6, 1):
                data = ref_file(bam, level, False, d)
                out["%s_c%d" % (case, level)] = np.array(hashlib.sha256(data).hexdigest())
                print(case, level, len(data))
        for name, bam in fixtures.header_cases().items():
            data = ref_file(bam, 6, False, d, fmt="rawbam")
            got = bamio.parse_bam_stream(data)
            out["header_" + name] = np.array(got.text)
            out["headerfile_" + name] = np.array(hashlib.sha256(ref_file(bam, 6, False, d)).hexdigest())
            print("header", name, repr(got.text)[:200])
    np.savez_compressed(os.path.join(HERE, "bamfile.npz"), **out)


if __name__ == "__main__":
    main()
