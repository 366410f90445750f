"""CPU suite: the C-ABI library builds, loads and exports everything include/oge_gpu_dedup.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from openge_b200 import dedup


def declared_functions():
    text = open(os.path.join(ROOT, "include", "oge_gpu_dedup.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oge_gpu_\w+)\s*\(", text)))


def test_header_symbols_exported():
    L = dedup.lib()
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(L, n), "libopenge_b200.so does not export %s" % n
    assert sorted(dedup.EXPORTS) == names


def test_abi_version_and_struct_sizes():
    L = dedup.lib()
    assert L.oge_gpu_abi_version() == dedup.ABI_VERSION
    assert C.sizeof(dedup.Config) == 80 == L.oge_gpu_sizeof(0)
    assert C.sizeof(dedup.Stats) == L.oge_gpu_sizeof(1)
    assert dedup.END_DTYPE.itemsize == 28 == L.oge_gpu_sizeof(2)
    assert L.oge_gpu_sizeof(3) == 13 * 8


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every context creation must fail loudly (never a CPU path)."""
    if dedup.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(dedup.DedupError) as e:
        dedup.DedupContext()
    assert e.value.code == -2
    with pytest.raises(dedup.DedupError):
        import numpy as np
        dedup.debug_sort128(np.zeros((4, 2), np.uint64), 0, 8)


def test_product_path_does_not_use_oracle():
    """The oracle is test infrastructure: nothing under openge_b200/ may import, load or link it
    (openge_b200/_build.py only knows how to BUILD the checker)."""
    bad = re.compile(r"import\s+oracle|from\s+oracle|liboge_oracle|markdup_oracle|oge_oracle_|oracle\.markdup|oge_ref_dedup")
    # the reference CLI shim names the binary it becomes in the oracle's build in a comment; nothing else may
    ref_cli = os.path.join("host", "refcli", "ref_driver.cpp")
    for d, _, files in os.walk(os.path.join(ROOT, "openge_b200")):
        for f in files:
            path = os.path.join(d, f)
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) and f != "_build.py" and not path.endswith(ref_cli):
                assert not bad.search(open(path).read()), path
            if f == "Makefile":      # the product's builds take no source or header from oracle/
                assert "oracle/" not in re.sub(r"#.*", "", open(path).read()), path
