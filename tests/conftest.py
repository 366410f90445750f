import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """-> (BamFile or None, dict of golden arrays)"""
    from openge_b200 import bamio, synth
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    if "records" in g:
        refs = list(zip([str(x) for x in g["ref_names"]], [int(x) for x in g["ref_lens"]]))
        bam = bamio.BamFile(text=str(g["text"]), refs=refs, records=g["records"], offsets=g["offsets"])
    else:
        bam = synth.make(name.split("_")[1], float(g["scale"]))
    return bam, g


GOLDEN_CASES = ["a3_fixture1", "a3_fixture2", "edge_cases", "yhet208",
                "synth_C1", "synth_C2", "synth_C3", "synth_C4", "synth_C5"]


@pytest.fixture(scope="session")
def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


LARGE_PIN_CASES = ["C1", "C2", "C3", "C4", "C5"]


def load_large_pin(name):
    """-> (BamFile regenerated from its seed, duplicate bit per record, sha256 of the reference's flag array):
    the compiled reference's own run on inputs of 1-5 M records (tests/golden/make_large_golden.py)."""
    import hashlib
    from openge_b200 import synth
    g = np.load(os.path.join(GOLDEN, "large_pins.npz"), allow_pickle=False)
    bam = synth.make(name, float(g[name + "_scale"]))
    assert bam.n == int(g[name + "_n"])
    assert hashlib.sha256(bam.records.tobytes()).hexdigest() == str(g[name + "_records_sha256"]), "synthetic generator drifted"
    dup = np.unpackbits(g[name + "_dupbits"])[: bam.n].astype(bool)
    return bam, dup, str(g[name + "_flags_sha256"])
