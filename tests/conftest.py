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
