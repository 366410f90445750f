"""ctypes front-end of tools/synth/oge_synth.c: synthetic workloads C1..C5.

Host tooling for tests and bench.py -- produces raw BAM records + offsets, the same
in-memory form bamio.read_bam returns.  ``scale`` shrinks a config to test size while
keeping its shape (contig count, fractions, read length).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

_HUMAN = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
          138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
          83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]


class SynthCfg(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_templates", C.c_uint64), ("n_contigs", C.c_int32),
        ("contig_len", C.c_int32 * 256), ("read_len", C.c_int32),
        ("insert_lo", C.c_int32), ("insert_hi", C.c_int32),
        ("dup_frac", C.c_double), ("softclip_frac", C.c_double), ("hardclip_frac", C.c_double),
        ("indel_frac", C.c_double), ("rf_frac", C.c_double), ("ff_frac", C.c_double),
        ("single_frac", C.c_double), ("mate_unmapped_frac", C.c_double),
        ("cross_contig_frac", C.c_double), ("secondary_frac", C.c_double),
        ("supplementary_frac", C.c_double), ("unmapped_pair_frac", C.c_double),
        ("predup_frac", C.c_double), ("no_rg_frac", C.c_double), ("unknown_rg_frac", C.c_double),
        ("const_qual_frac", C.c_double), ("extra_tag_frac", C.c_double),
        ("n_rg", C.c_int32), ("hot_loci", C.c_int32), ("dup_same_rg", C.c_int32),
        ("contig_lo", C.c_int32), ("contig_hi", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.ensure_synth()
        L = C.CDLL(path)
        L.oge_synth_plan.restype = C.c_void_p
        L.oge_synth_plan.argtypes = [C.POINTER(SynthCfg), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oge_synth_emit.restype = C.c_int
        L.oge_synth_emit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oge_synth_free.argtypes = [C.c_void_p]
        L.oge_merge_sorted.restype = C.c_uint64
        L.oge_merge_sorted.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oge_frame_records.restype = C.c_int64
        L.oge_frame_records.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


def frame_records_fast(raw: bytes, start: int = 0) -> np.ndarray:
    """offsets (u64, n+1, relative to ``start``) of the record chain in ``raw[start:]``."""
    L = lib()
    buf = np.frombuffer(raw, dtype=np.uint8)[start:]
    n = L.oge_frame_records(buf.ctypes.data, len(buf), None, 0)
    if n < 0:
        raise ValueError("malformed BAM record chain")
    offs = np.zeros(n + 1, dtype=np.uint64)
    L.oge_frame_records(buf.ctypes.data, len(buf), offs.ctypes.data, n + 1)
    return offs


def _cfg(seed, n_templates, contigs, read_len, **kw):
    c = SynthCfg()
    c.seed, c.n_templates, c.n_contigs, c.read_len = seed, n_templates, len(contigs), read_len
    for i, l in enumerate(contigs):
        c.contig_len[i] = l
    c.insert_lo, c.insert_hi = kw.pop("insert", (120, 500))
    c.n_rg = kw.pop("n_rg", 1)
    c.dup_same_rg = kw.pop("dup_same_rg", 1)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def header_text(contigs, rgs, sort_order="coordinate"):
    """Canonical-form header (bam_header.cpp:184-214): @HD, @SQ..., @RG..."""
    lines = ["@HD\tVN:1.4\tSO:%s" % sort_order]
    lines += ["@SQ\tSN:%s\tLN:%d" % (n, l) for n, l in contigs]
    for rid, lb in rgs:
        lines.append("@RG\tID:%s%s\tSM:s" % (rid, ("\tLB:" + lb) if lb else ""))
    return "\n".join(lines) + "\n"


def config(name: str, scale: float = 1.0):
    """-> (SynthCfg, contigs [(name,len)], read groups [(id, LB)]) for 'C1'..'C5'."""
    name = name.upper()
    if name == "C1":      # 1 M reads 2x100, 4 contigs x 60 Mb, 10 % duplicate pairs, one library
        contigs = [60_000_000] * 4
        c = _cfg(1, int(500_000 * scale), contigs, 100, dup_frac=0.10)
        rgs = [("rg1", "lib1")]
    elif name == "C2":    # 50 M reads 2x150, human-like, 10 % dups, 10 % soft-clipped ends, FR and RF
        contigs = _HUMAN
        c = _cfg(2, int(25_000_000 * scale), contigs, 150, insert=(200, 600), dup_frac=0.10,
                 softclip_frac=0.10, rf_frac=0.05)
        rgs = [("rg1", "lib1")]
    elif name == "C3":    # fragment-heavy mixed SE/PE, unmapped mates, clips, indels, several libraries
        contigs = [5_000_000, 3_000_000, 2_000_000, 1_000_000, 500_000, 200_000]
        c = _cfg(3, int(6_400_000 * scale), contigs, 100, insert=(120, 500), dup_frac=0.15,
                 softclip_frac=0.25, hardclip_frac=0.3, indel_frac=0.15, rf_frac=0.05, ff_frac=0.03,
                 single_frac=0.50, mate_unmapped_frac=0.10, cross_contig_frac=0.03,
                 secondary_frac=0.08, supplementary_frac=0.02, unmapped_pair_frac=0.01,
                 predup_frac=0.02, no_rg_frac=0.05, unknown_rg_frac=0.02, const_qual_frac=0.3,
                 extra_tag_frac=0.5, n_rg=5, dup_same_rg=0)
        rgs = [("rg1", "libA"), ("rg2", "libA"), ("rg3", "libB"), ("rg4", "libC"), ("rg5", "")]
    elif name == "C4":    # exome-like: 40 % duplicates from 1 % of loci, long equal-key runs, score ties
        contigs = _HUMAN[:8]
        n_t = int(10_000_000 * scale)
        c = _cfg(4, n_t, contigs, 100, dup_frac=0.40, const_qual_frac=0.6,
                 hot_loci=max(1, int(n_t * 0.6 * 0.01)))
        rgs = [("rg1", "lib1")]
    elif name == "C5":    # 30x WGS shape, 10 % dups, 0.5 % cross-contig pairs (per-shard generation)
        contigs = _HUMAN
        c = _cfg(5, int(400_000_000 * scale), contigs, 150, insert=(200, 600), dup_frac=0.10,
                 softclip_frac=0.05, cross_contig_frac=0.005)
        rgs = [("rg1", "lib1")]
    else:
        raise KeyError(name)
    cl = [("chr%d" % (i + 1), l) for i, l in enumerate(contigs)]
    return c, cl, rgs


def generate(cfg: SynthCfg, records_out=None, nthreads: int = 0):
    """Run the generator -> (records uint8 array, offsets uint64 n+1).

    ``records_out``: optional callable ``nbytes -> writable uint8 numpy array`` (e.g. pinned
    host memory); default allocates a numpy array.
    """
    L = lib()
    n, nb = C.c_uint64(), C.c_uint64()
    h = L.oge_synth_plan(C.byref(cfg), C.byref(n), C.byref(nb))
    if not h:
        raise MemoryError("oge_synth_plan failed")
    try:
        rec = records_out(nb.value) if records_out else np.empty(nb.value + 16, dtype=np.uint8)[: nb.value]
        offs = np.empty(n.value + 1, dtype=np.uint64)
        L.oge_synth_emit(h, rec.ctypes.data, offs.ctypes.data, nthreads or min(32, os.cpu_count() or 1))
    finally:
        L.oge_synth_free(h)
    return rec, offs


def restrict(cfg: SynthCfg, all_contigs, contig_lo: int, contig_hi: int, seed: int, name_base: int = 0) -> SynthCfg:
    """A copy of ``cfg`` over the contig table ``all_contigs`` [(name, len)] that draws its templates on
    contigs [contig_lo, contig_hi) only (range shards of a multi-GPU run).  Names are a hash of
    (template ordinal, seed): different seeds give disjoint name sets."""
    c = SynthCfg.from_buffer_copy(bytes(cfg))
    c.seed = seed
    c.n_contigs = len(all_contigs)
    for i, (_, ln) in enumerate(all_contigs):
        c.contig_len[i] = ln
    c.contig_lo, c.contig_hi = contig_lo, contig_hi
    return c


def remap_pieces(rec, offs, piece_ref, piece_start, nthreads: int = 0):
    """Records drawn on "pieces" (stretches of real contigs, each generated as a contig of its own) onto the real contigs, in
    place: refID piece -> piece_ref[piece], pos += piece_start[piece], same for the mate fields."""
    L = lib()
    pr = np.ascontiguousarray(piece_ref, dtype=np.int32)
    ps = np.ascontiguousarray(piece_start, dtype=np.int32)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    L.oge_synth_remap.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int]
    rc = L.oge_synth_remap(rec.ctypes.data, offs.ctypes.data, len(offs) - 1, pr.ctypes.data, ps.ctypes.data, len(pr), nthreads or min(32, os.cpu_count() or 1))
    if rc != 0:
        raise RuntimeError("oge_synth_remap failed")


def records_in_range(rec, offs, lo, hi) -> np.ndarray:
    """uint8 mask over the records: (refID, pos) in [lo, hi), both (ref, pos) tuples; records without a reference never."""
    o = offs[:-1].astype(np.int64)
    if not len(o):
        return np.zeros(0, np.uint8)
    rp = rec[o[:, None] + np.arange(4, 12)].copy().view("<i4").reshape(-1, 2)
    key = rp[:, 0].astype(np.int64) * (1 << 32) + rp[:, 1].astype(np.int64)
    klo, khi = lo[0] * (1 << 32) + lo[1], hi[0] * (1 << 32) + hi[1]
    return ((rp[:, 0] >= 0) & (key >= klo) & (key < khi)).astype(np.uint8)


def records_on_contigs(rec, offs, contig_lo: int, contig_hi: int) -> np.ndarray:
    """uint8 mask over the records: refID in [contig_lo, contig_hi)."""
    o = offs[:-1].astype(np.int64)
    ref = rec[o[:, None] + np.arange(4, 8)].copy().view("<i4").ravel() if len(o) else np.zeros(0, "<i4")
    return ((ref >= contig_lo) & (ref < contig_hi)).astype(np.uint8)


def merge_sorted(rec_a, offs_a, rec_b, offs_b, keep_b, records_out=None):
    """Merge two coordinate-sorted record chains (B restricted to keep_b) -> (records, offsets)."""
    L = lib()
    rec_a, rec_b = np.ascontiguousarray(rec_a), np.ascontiguousarray(rec_b)
    offs_a, offs_b = np.ascontiguousarray(offs_a, dtype=np.uint64), np.ascontiguousarray(offs_b, dtype=np.uint64)
    keep_b = np.ascontiguousarray(keep_b, dtype=np.uint8)
    na, nb = len(offs_a) - 1, len(offs_b) - 1
    nbytes = C.c_uint64()
    n = L.oge_merge_sorted(rec_a.ctypes.data, offs_a.ctypes.data, na, rec_b.ctypes.data, offs_b.ctypes.data, nb,
                           keep_b.ctypes.data, None, None, C.byref(nbytes))
    out = records_out(nbytes.value) if records_out else np.empty(nbytes.value + 16, dtype=np.uint8)[: nbytes.value]
    offs = np.empty(n + 1, dtype=np.uint64)
    L.oge_merge_sorted(rec_a.ctypes.data, offs_a.ctypes.data, na, rec_b.ctypes.data, offs_b.ctypes.data, nb,
                       keep_b.ctypes.data, out.ctypes.data, offs.ctypes.data, C.byref(nbytes))
    return out, offs


def make(name: str, scale: float = 1.0, seed=None):
    """-> bamio.BamFile for config ``name`` at ``scale``."""
    from .bamio import BamFile
    cfg, contigs, rgs = config(name, scale)
    if seed is not None:
        cfg.seed = seed
    rec, offs = generate(cfg)
    return BamFile(text=header_text(contigs, rgs), refs=contigs, records=rec, offsets=offs)
