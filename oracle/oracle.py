"""TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle/markdup_oracle.c and of the
compiled reference (oracle/_ref/oge_ref_dedup).  Import only from tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from openge_b200 import _build, bamio  # noqa: E402  (build recipes and file I/O only: no product logic)


class OracleEnd(C.Structure):
    _fields_ = [("eligible", C.c_int32), ("pair_eligible", C.c_int32), ("ref", C.c_int32),
                ("coord", C.c_int32), ("orientation", C.c_int32), ("read2Sequence", C.c_int32),
                ("score", C.c_int16), ("lib", C.c_int16)]


END_DTYPE = np.dtype([("eligible", "<i4"), ("pair_eligible", "<i4"), ("ref", "<i4"), ("coord", "<i4"),
                      ("orientation", "<i4"), ("read2Sequence", "<i4"), ("score", "<i2"), ("lib", "<i2")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.ensure_oracle())
        L.oge_oracle_markdup.restype = C.c_int
        L.oge_oracle_markdup.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_char_p), C.c_void_p,
                                         C.c_int32, C.c_int16, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oge_oracle_markdup_text.restype = C.c_int
        L.oge_oracle_markdup_text.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oge_oracle_coordinate_order.restype = C.c_int
        L.oge_oracle_coordinate_order.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.oge_oracle_flagstats.restype = C.c_int
        L.oge_oracle_flagstats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def coordinate_order(records: np.ndarray, offsets: np.ndarray):
    """The order `openge mergesort` gives the records (read_sorter.cpp + Sort::ByPosition), ties and the unplaced tail in
    input order.  -> (perm, tied): perm[k] = input ordinal at output position k; tied[k] = the reference does not define
    the order inside the group position k belongs to."""
    L = lib()
    records = np.ascontiguousarray(records)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    perm = np.zeros(max(1, n), dtype=np.uint32)
    tied = np.zeros(max(1, n), dtype=np.uint8)
    if L.oge_oracle_coordinate_order(records.ctypes.data, offsets.ctypes.data, n, perm.ctypes.data, tied.ctypes.data) != 0:
        raise RuntimeError("oracle failed")
    return perm[:n], tied[:n].astype(bool)


def ref_sort(bam, dedup=False, per_tempfile=200000, tmpdir=None, timeout=120):
    """The compiled reference's ReadSorter chain (`openge mergesort [-M]`, command_mergesort.cpp:70-113) -> output BamFile."""
    exe = _build.ensure_ref()
    if exe is None:
        raise RuntimeError("oracle/_ref/oge_ref_dedup not built")
    base = tmpdir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    with tempfile.TemporaryDirectory(dir=base) as d:
        inp, out = os.path.join(d, "in.rawbam"), os.path.join(d, "out.rawbam")
        bamio.write_bam(inp, bam, raw=True)
        cmd = [exe, "-T", d, "--sort", "-v", "-F", "rawbam", "-n", str(per_tempfile)] + ([] if dedup else ["--nodedup"]) + [inp, out]
        for attempt in range(4):
            try:
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
            except subprocess.TimeoutExpired:
                continue
            if r.returncode != 0:
                raise RuntimeError("reference failed: %s" % r.stderr.decode()[-2000:])
            return bamio.read_bam(out)
        raise RefHang("reference did not terminate in %d s (4 attempts)" % timeout)


FLAGSTAT_FIELDS = ("reads", "mapped", "forward", "reverse", "failed_qc", "duplicates", "paired", "proper_pair",
                   "both_mapped", "first_mate", "second_mate", "singletons", "sorted")


def flagstats(records: np.ndarray, offsets: np.ndarray, flags: np.ndarray | None = None) -> dict:
    """Statistics::runInternal's counters (statistics.cpp:77-162) over framed records; `flags` overrides the
    records' flag words (the dedup output)."""
    L = lib()
    records = np.ascontiguousarray(records)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    out = np.zeros(13, dtype=np.uint64)
    fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint16)
    rc = L.oge_oracle_flagstats(records.ctypes.data, offsets.ctypes.data, n, None if fl is None else fl.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle failed")
    return dict(zip(FLAGSTAT_FIELDS, (int(x) for x in out)))


def markdup(records: np.ndarray, offsets: np.ndarray, text: str, compat_quiet: bool = False, want_ends: bool = False):
    """CPU oracle over framed records -> flags (u16 per record) [, ends, stats]."""
    L = lib()
    n = len(offsets) - 1
    flags = np.zeros(n, dtype=np.uint16)
    ends = np.zeros(n, dtype=END_DTYPE) if want_ends else None
    stats = np.zeros(4, dtype=np.uint64)
    records = np.ascontiguousarray(records)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    # the @RG -> LB -> library id resolution is the oracle's own (parse_read_groups in markdup_oracle.c)
    rc = L.oge_oracle_markdup_text(records.ctypes.data, offsets.ctypes.data, n, text.encode("latin-1"), int(compat_quiet),
                                   flags.ctypes.data, ends.ctypes.data if want_ends else None, stats.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle failed")
    if want_ends:
        return flags, ends, stats
    return flags


def markdup_split(bam, n_chains: int):
    """Compat F2: independent runs over refID % n sub-streams (split_by_chromosome.cpp:45-50)."""
    off = bam.offsets[:-1].astype(np.int64)
    ref = bam.records[off[:, None] + np.arange(4, 8)].copy().view("<i4").ravel()
    chain = np.where(ref < 0, 0, ref % n_chains)
    flags = np.zeros(bam.n, dtype=np.uint16)
    sizes = np.diff(bam.offsets.astype(np.int64))
    for c in range(n_chains):
        idx = np.nonzero(chain == c)[0]
        if len(idx) == 0:
            continue
        recs = [bam.records[off[i]: off[i] + sizes[i]].tobytes() for i in idx]
        r, o = bamio.concat_records(recs)
        flags[idx] = markdup(r, o, bam.text)
    return flags


def ref_available() -> bool:
    return _build.ensure_ref() is not None


class RefHang(RuntimeError):
    pass


def ref_dedup(bam, nosplit=True, verbose=True, remove=False, threads=None, tmpdir=None, timeout=60):
    """Run the compiled reference (file -> file, rawbam both ways) -> output BamFile."""
    exe = _build.ensure_ref()
    if exe is None:
        raise RuntimeError("oracle/_ref/oge_ref_dedup not built")
    base = tmpdir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    with tempfile.TemporaryDirectory(dir=base) as d:
        inp, out = os.path.join(d, "in.rawbam"), os.path.join(d, "out.rawbam")
        bamio.write_bam(inp, bam, raw=True)
        cmd = [exe, "-T", d, "-F", "rawbam"]
        if verbose:
            cmd.append("-v")
        if nosplit:
            cmd.append("--nosplit")
        if remove:
            cmd.append("-r")
        if threads:
            cmd += ["-t", str(threads)]
        cmd += [inp, out]
        # the reference's pipeline is racy (getInputAlignment tail-drop window) and
        # now and then never terminates: bound every run and retry
        for attempt in range(4):
            try:
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
            except subprocess.TimeoutExpired:
                continue
            if r.returncode != 0:
                raise RuntimeError("reference failed: %s" % r.stderr.decode()[-2000:])
            return bamio.read_bam(out)
        raise RefHang("reference did not terminate in %d s (4 attempts)" % timeout)


def ref_time_mem(bam, reps=1, threads=None, tmpdir=None, timeout=900):
    """Time MarkDuplicates::runInternal in the compiled reference with records preloaded in RAM.
    -> dict(records, threads, seconds[list], duplicates)"""
    exe = _build.ensure_ref()
    if exe is None:
        raise RuntimeError("oracle/_ref/oge_ref_dedup not built")
    base = tmpdir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    with tempfile.TemporaryDirectory(dir=base) as d:
        inp = os.path.join(d, "in.rawbam")
        bamio.write_bam(inp, bam, raw=True)
        cmd = [exe, "--mem", "-v", "-T", d, "--reps", str(reps)]
        if threads:
            cmd += ["-t", str(threads)]
        cmd.append(inp)
        try:
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=timeout)
        except subprocess.TimeoutExpired:
            raise RefHang("reference did not terminate in %d s" % timeout)
        if r.returncode != 0:
            raise RuntimeError("reference failed")
        return json.loads(r.stdout.decode().strip().splitlines()[-1])


def ref_flags_mem(bam, threads=None, tmpdir=None, timeout=1800):
    """The compiled reference's MarkDuplicates::runInternal over records preloaded in RAM (`--mem -v --flags F`):
    -> the output flag word of every record.  No file writer behind it, so it scales to millions of records."""
    exe = _build.ensure_ref()
    if exe is None:
        raise RuntimeError("oracle/_ref/oge_ref_dedup not built")
    base = tmpdir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    with tempfile.TemporaryDirectory(dir=base) as d:
        inp, out = os.path.join(d, "in.rawbam"), os.path.join(d, "flags.u16")
        bamio.write_bam(inp, bam, raw=True)
        cmd = [exe, "--mem", "-v", "-T", d, "--reps", "1", "--flags", out]
        if threads:
            cmd += ["-t", str(threads)]
        cmd.append(inp)
        for attempt in range(3):
            try:
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=timeout)
            except subprocess.TimeoutExpired:
                continue
            if r.returncode != 0:
                raise RuntimeError("reference failed")
            f = np.fromfile(out, dtype=np.uint16)
            if len(f) != bam.n:
                raise RuntimeError("reference returned %d flag words for %d records" % (len(f), bam.n))
            return f
        raise RefHang("reference did not terminate in %d s (3 attempts)" % timeout)


_STAT_LABELS = {"Total reads": "reads", "Mapped reads": "mapped", "Forward strand": "forward", "Reverse strand": "reverse",
                "Failed QC": "failed_qc", "Duplicates": "duplicates", "Paired-end reads": "paired", "'Proper-pairs'": "proper_pair",
                "Both pairs mapped": "both_mapped", "Read 1": "first_mate", "Read 2": "second_mate", "Singletons": "singletons",
                "Sorted": "sorted"}


def ref_stats(bam, tmpdir=None, timeout=60) -> dict:
    """The compiled reference's own Statistics module (algorithms/statistics.cpp) run behind its MarkDuplicates
    (`--nosplit -v --stats`): the printed report parsed into the FLAGSTAT_FIELDS dict.  Lines the reference omits
    when there are no paired reads (statistics.cpp:165-171) come back as 0."""
    exe = _build.ensure_ref()
    if exe is None:
        raise RuntimeError("oracle/_ref/oge_ref_dedup not built")
    base = tmpdir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    with tempfile.TemporaryDirectory(dir=base) as d:
        inp, out = os.path.join(d, "in.rawbam"), os.path.join(d, "out.rawbam")
        bamio.write_bam(inp, bam, raw=True)
        cmd = [exe, "-T", d, "-F", "rawbam", "-v", "--nosplit", "--stats", inp, out]
        for attempt in range(4):
            try:
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
            except subprocess.TimeoutExpired:
                continue
            if r.returncode != 0:
                raise RuntimeError("reference failed: %s" % r.stderr.decode()[-2000:])
            res = {k: 0 for k in FLAGSTAT_FIELDS}
            for line in r.stdout.decode().splitlines():
                label, _, rest = line.partition(":")
                if label in _STAT_LABELS and rest.strip():
                    tok = rest.split()[0]
                    res[_STAT_LABELS[label]] = {"Yes": 1, "No": 0}.get(tok, None) if tok in ("Yes", "No") else int(tok)
            return res
        raise RefHang("reference did not terminate in %d s (4 attempts)" % timeout)
