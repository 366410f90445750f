"""ctypes binding of libopenge_b200.so (include/oge_gpu_dedup.h) + a MarkDuplicates front-end.

This is host plumbing for tests, the CLI and bench.py; the algorithm runs in the CUDA
library.  There is no CPU fallback: loading fails loudly when the library is missing, and
``oge_gpu_dedup_create`` fails when no sm_100 device is present.

``MarkDuplicates`` mirrors the reference's algorithm module
(/root/reference/openge/src/algorithms/mark_duplicates.h:27-68): constructor takes the temp
directory (unused here: nothing is spilled), ``removeDuplicates`` is a public attribute, and
``run(bam)`` does what ``runInternal`` does to the record stream.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build, header as _header
from .bamio import BamFile

ABI_VERSION = 8

OGE_OK = 0
ERRORS = {-1: "OGE_ERR_INVALID_ARG", -2: "OGE_ERR_CUDA", -3: "OGE_ERR_NOMEM", -4: "OGE_ERR_KEY_RANGE",
          -5: "OGE_ERR_STATE", -6: "OGE_ERR_TOO_LARGE", -7: "OGE_ERR_BAD_RECORD"}


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n_ref", C.c_int32), ("max_ref_len", C.c_int32),
                ("clip_margin", C.c_int32), ("remove_duplicates", C.c_int32), ("verify_names", C.c_int32),
                ("compat_quiet_index_bug", C.c_int32), ("debug_keep_ends", C.c_int32), ("profile_events", C.c_int32),
                ("capacity_records", C.c_uint64), ("capacity_bytes", C.c_uint64),
                ("rank", C.c_int32), ("world", C.c_int32), ("index_base", C.c_uint64),
                ("debug_full_frag_sort", C.c_int32), ("debug_legacy_join", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_frag_entries", C.c_uint64), ("n_pair_entries", C.c_uint64),
                ("n_duplicates", C.c_uint64), ("n_complex_names", C.c_uint64), ("n_hash_mismatch", C.c_uint64),
                ("frag_key_bits", C.c_uint32), ("pair_key_bits", C.c_uint32),
                ("frag_sort_passes", C.c_uint32), ("pair_sort_passes", C.c_uint32),
                ("ms_total", C.c_float), ("ms_endbuild", C.c_float), ("ms_join", C.c_float),
                ("ms_sort_frag", C.c_float), ("ms_sort_pair", C.c_float), ("ms_select", C.c_float),
                ("ms_flags", C.c_float), ("launches", C.c_uint64),
                ("ms_sort_pass_kernels", C.c_float), ("sort_pass_launches", C.c_uint32), ("sort_pass_bytes", C.c_uint64),
                ("ms_inflate", C.c_float), ("ms_frame", C.c_float), ("inflate_blocks", C.c_uint64),
                ("inflate_bytes_in", C.c_uint64), ("inflate_bytes_out", C.c_uint64), ("frame_repairs", C.c_uint64),
                ("ms_inflate_h2d", C.c_float), ("ms_inflate_d2h", C.c_float),
                ("n_local_pairs", C.c_uint64), ("n_join_leftovers", C.c_uint64), ("n_local_retracted", C.c_uint64),
                ("ms_kernel", C.c_float * 8),
                ("ms_push_bgzf", C.c_float), ("inflate_mode", C.c_uint32), ("inflate_pieces", C.c_uint32),
                ("ms_inflate_start", C.c_float),
                ("ms_deflate", C.c_float), ("deflate_blocks", C.c_uint64), ("deflate_bytes_in", C.c_uint64), ("deflate_bytes_out", C.c_uint64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["ms_kernel"] = dict(zip(KERNEL_NAMES, (float(x) for x in self.ms_kernel)))
        return d


class StepInfo(C.Structure):
    _fields_ = [("published_in", C.c_uint64), ("routed_in", C.c_uint64), ("marks_in", C.c_uint64), ("exchanges", C.c_uint64),
                ("bytes_sent", C.c_uint64)]


KERNEL_NAMES = ("endbuild", "match", "emit", "global_join", "check", "select", "flags", "sort_hist")      # OGE_K_* order


FLAGSTAT_FIELDS = ("reads", "mapped", "forward", "reverse", "failed_qc", "duplicates", "paired", "proper_pair",
                   "both_mapped", "first_mate", "second_mate", "singletons", "sorted")      # oge_gpu_flagstats, in order


END_DTYPE = np.dtype([("eligible", "<i4"), ("pair_eligible", "<i4"), ("ref", "<i4"), ("coord", "<i4"),
                      ("orientation", "<i4"), ("read2Sequence", "<i4"), ("score", "<i2"), ("lib", "<i2")])

EXPORTS = ["oge_gpu_dedup_create", "oge_gpu_dedup_destroy", "oge_gpu_dedup_set_readgroups", "oge_gpu_dedup_push",
           "oge_gpu_dedup_push_bgzf", "oge_gpu_dedup_set_offsets", "oge_gpu_dedup_frame", "oge_gpu_dedup_offsets",
           "oge_gpu_dedup_sync", "oge_gpu_dedup_run", "oge_gpu_dedup_flags", "oge_gpu_dedup_pull", "oge_gpu_dedup_deflate", "oge_gpu_dedup_pull_bgzf", "oge_gpu_dedup_pull_bgzf_part", "oge_gpu_dedup_pull_bgzf_wait",
           "oge_gpu_dedup_reset", "oge_gpu_dedup_flagstats", "oge_gpu_dedup_get_stats", "oge_gpu_dedup_debug_ends",
           "oge_gpu_dedup_device_ptrs", "oge_gpu_host_alloc", "oge_gpu_host_free", "oge_gpu_device_count",
           "oge_gpu_last_error", "oge_gpu_abi_version", "oge_gpu_debug_sort128", "oge_gpu_debug_sort_bench",
           "oge_gpu_set_sort_variant", "oge_gpu_set_inflate_kernel", "oge_gpu_inflate_kernel", "oge_gpu_set_bgzf_chunk_bytes", "oge_gpu_set_bgzf_staging", "oge_gpu_shard_setup", "oge_gpu_shard_begin", "oge_gpu_shard_probe",
           "oge_gpu_shard_finish", "oge_gpu_shard_apply", "oge_gpu_copy_d2d", "oge_gpu_dedup_sort", "oge_gpu_dedup_sort_order",
           "oge_gpu_dedup_sort_stats", "oge_gpu_sizeof", "oge_gpu_shard_key_bytes", "oge_gpu_shard_set_entry_bytes", "oge_gpu_shard_replay",
           "oge_gpu_shard_comm_id", "oge_gpu_shard_comm_init", "oge_gpu_shard_comm_destroy", "oge_gpu_shard_step"]


class DedupError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "error"), code, msg))
        self.code = code


_lib = None
_testing_lib = None
_use_testing = False


class testing_library:
    """Tests only: inside the `with` block every call goes to libopenge_b200_testing.so, the same sources compiled
    with -DOGE_TESTING (forced-overflow capacities read from the environment, measurement knobs of the sort).  The
    product library carries none of these hooks.  Contexts must be created and closed inside the block."""

    def __enter__(self):
        global _use_testing
        self._prev, _use_testing = _use_testing, True
        lib()
        return self

    def __exit__(self, *a):
        global _use_testing
        _use_testing = self._prev


def lib():
    """Load (building if a compiler is present and sources are newer) the CUDA library."""
    global _lib, _testing_lib
    if _use_testing:
        if _testing_lib is None:
            _testing_lib = _load(_build.build_gpu(testing=True))
        return _testing_lib
    if _lib is None:
        _lib = _load(_build.build_gpu())
    return _lib


def _load(path):
    if True:
        if not os.path.exists(path):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % os.path.basename(path))
        L = C.CDLL(path)
        vp, u64 = C.c_void_p, C.c_uint64
        L.oge_gpu_dedup_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.oge_gpu_dedup_destroy.argtypes = [vp]
        L.oge_gpu_dedup_destroy.restype = None
        L.oge_gpu_dedup_set_readgroups.argtypes = [vp, C.POINTER(C.c_char_p), vp, C.c_int32, C.c_int16, C.c_int32]
        L.oge_gpu_dedup_push.argtypes = [vp, vp, u64, vp, u64]
        L.oge_gpu_dedup_sync.argtypes = [vp]
        L.oge_gpu_dedup_push_bgzf.argtypes = [vp, vp, u64, vp, vp, vp, u64, u64, vp]
        L.oge_gpu_dedup_set_offsets.argtypes = [vp, vp, u64]
        L.oge_gpu_dedup_frame.argtypes = [vp, C.POINTER(u64)]
        L.oge_gpu_dedup_offsets.argtypes = [vp, vp, u64]
        L.oge_gpu_dedup_run.argtypes = [vp]
        L.oge_gpu_dedup_sort.argtypes = [vp]
        L.oge_gpu_dedup_sort_order.argtypes = [vp, vp, u64]
        L.oge_gpu_dedup_sort_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_float)]
        L.oge_gpu_dedup_flags.argtypes = [vp, vp, u64]
        L.oge_gpu_dedup_pull.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(u64)]
        L.oge_gpu_dedup_deflate.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.oge_gpu_dedup_pull_bgzf.argtypes = [vp, vp, u64]
        L.oge_gpu_dedup_pull_bgzf_part.argtypes = [vp, u64, u64, vp]
        L.oge_gpu_dedup_pull_bgzf_wait.argtypes = [vp]
        L.oge_gpu_dedup_reset.argtypes = [vp]
        L.oge_gpu_dedup_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.oge_gpu_dedup_flagstats.argtypes = [vp, vp]
        L.oge_gpu_dedup_debug_ends.argtypes = [vp, vp, u64]
        L.oge_gpu_dedup_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.oge_gpu_host_alloc.argtypes = [C.c_size_t]
        L.oge_gpu_host_alloc.restype = vp
        L.oge_gpu_host_free.argtypes = [vp]
        L.oge_gpu_host_free.restype = None
        L.oge_gpu_last_error.restype = C.c_char_p
        L.oge_gpu_debug_sort128.argtypes = [C.c_int, vp, u64, C.c_int, C.c_int]
        L.oge_gpu_debug_sort_bench.argtypes = [C.c_int, u64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u64,
                                               C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oge_gpu_set_sort_variant.argtypes = [C.c_int]
        L.oge_gpu_set_inflate_kernel.argtypes = [C.c_int]
        L.oge_gpu_inflate_kernel.argtypes = [C.c_int]
        L.oge_gpu_set_bgzf_chunk_bytes.argtypes = [u64]
        L.oge_gpu_set_bgzf_staging.argtypes = [u64]
        L.oge_gpu_copy_d2d.argtypes = [vp, vp, vp, u64]
        L.oge_gpu_shard_setup.argtypes = [vp, u64, vp, vp, vp]
        L.oge_gpu_shard_key_bytes.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.oge_gpu_shard_set_entry_bytes.argtypes = [vp, C.c_uint32]
        L.oge_gpu_shard_begin.argtypes = [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
        L.oge_gpu_shard_probe.argtypes = [vp, vp, u64, vp, u64, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
        L.oge_gpu_shard_replay.argtypes = [vp, vp, u64, C.POINTER(vp), C.POINTER(u64)]
        L.oge_gpu_shard_finish.argtypes = [vp, vp, u64, C.POINTER(vp), C.POINTER(u64)]
        L.oge_gpu_shard_apply.argtypes = [vp, vp, u64]
        L.oge_gpu_shard_comm_id.argtypes = [vp]
        L.oge_gpu_shard_comm_init.argtypes = [vp, vp]
        L.oge_gpu_shard_comm_destroy.argtypes = [vp]
        L.oge_gpu_shard_comm_destroy.restype = None
        L.oge_gpu_shard_step.argtypes = [vp, C.POINTER(StepInfo)]
        for name in EXPORTS:
            getattr(L, name)
        L.oge_gpu_sizeof.argtypes = [C.c_int]
        if L.oge_gpu_sizeof(0) != C.sizeof(Config) or L.oge_gpu_sizeof(1) != C.sizeof(Stats):
            raise ImportError("%s was built from another oge_gpu_dedup.h: config %d / %d bytes, stats %d / %d bytes" % (
                os.path.basename(path), L.oge_gpu_sizeof(0), C.sizeof(Config), L.oge_gpu_sizeof(1), C.sizeof(Stats)))
    return L


def _check(rc):
    if rc != OGE_OK:
        raise DedupError(rc, lib().oge_gpu_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return int(lib().oge_gpu_device_count())


def debug_sort128(entries: np.ndarray, bit_lo: int, bit_hi: int, device: int = 0) -> np.ndarray:
    """Test hook: K3 alone.  entries: (n, 2) uint64 (lo, hi); returns the sorted copy."""
    e = np.ascontiguousarray(entries, dtype=np.uint64).copy()
    _check(lib().oge_gpu_debug_sort128(device, e.ctypes.data, len(e), bit_lo, bit_hi))
    return e


def set_sort_variant(variant: int):
    """Tuning hook: which onesweep pass kernel runs (see radix_sort.cuh)."""
    _check(lib().oge_gpu_set_sort_variant(variant))


INFLATE_KERNELS = {"threads": 0, "warp": 1, "engine": 2, "default": -1}


def set_inflate_kernel(kernel: str):
    """Which BGZF decoder push_bgzf uses: "engine" (the B200's hardware decompress engine, default where the device has
    one), "warp" (one warp per block), "threads" (one thread per block), "default"."""
    _check(lib().oge_gpu_set_inflate_kernel(INFLATE_KERNELS[kernel]))


def inflate_kernel(device: int = 0) -> str:
    """The decoder push_bgzf would use on `device` now."""
    k = lib().oge_gpu_inflate_kernel(device)
    if k < 0:
        _check(k)
    return {v: n for n, v in INFLATE_KERNELS.items()}[k]


def set_bgzf_staging(stage_bytes: int = 32 << 20):
    """Size of the two pinned staging buffers a compressed file in pageable memory goes up through (0: off)."""
    _check(lib().oge_gpu_set_bgzf_staging(stage_bytes))


def set_bgzf_chunk_bytes(nbytes: int):
    """Tuning hook: compressed bytes per piece of push_bgzf's overlapped upload (0 default, 2**64-1 one piece)."""
    _check(lib().oge_gpu_set_bgzf_chunk_bytes(nbytes))


def debug_sort_bench(n, bit_lo, bit_hi, variant=0, mode=0, reps=3, seed=1, device=0) -> dict:
    """K3 alone on device-generated entries: CUDA-event time per pass launch and per whole sort,
    verified on the device (sortedness on the bit range + order-independent checksums)."""
    ms_pass, ms_sort, n_pass, ok = C.c_float(), C.c_float(), C.c_int(), C.c_int()
    _check(lib().oge_gpu_debug_sort_bench(device, n, bit_lo, bit_hi, variant, mode, reps, seed, C.byref(ms_pass),
                                          C.byref(ms_sort), C.byref(n_pass), C.byref(ok)))
    return {"n": n, "bits": [bit_lo, bit_hi], "variant": variant, "mode": mode, "ms_per_pass": ms_pass.value,
            "ms_sort": ms_sort.value, "passes": n_pass.value, "verified": bool(ok.value),
            "pass_GBps": n * 32 / 1e9 / (ms_pass.value * 1e-3) if ms_pass.value > 0 else None}


class PinnedBuffer:
    """Page-locked host memory from oge_gpu_host_alloc, viewed as a uint8 numpy array."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = lib().oge_gpu_host_alloc(max(1, self.nbytes))
        if not self.ptr:
            raise MemoryError("oge_gpu_host_alloc(%d) failed" % nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, self.nbytes)).from_address(self.ptr))[: self.nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            lib().oge_gpu_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DedupContext:
    """One oge_gpu_dedup_ctx (one MarkDuplicates instance on one GPU)."""

    def __init__(self, n_ref=0, max_ref_len=0, device=0, remove_duplicates=False, verify_names=True,
                 compat_quiet_index_bug=False, debug_keep_ends=False, clip_margin=0, capacity_records=0,
                 capacity_bytes=0, index_base=0, rank=0, world=1, profile_events=False, full_frag_sort=False, legacy_join=False):
        self._h = C.c_void_p()
        cfg = Config(abi_version=ABI_VERSION, device=device, n_ref=n_ref, max_ref_len=max_ref_len, clip_margin=clip_margin,
                     remove_duplicates=int(remove_duplicates), verify_names=int(verify_names),
                     compat_quiet_index_bug=int(compat_quiet_index_bug), debug_keep_ends=int(debug_keep_ends),
                     profile_events=int(profile_events), capacity_records=capacity_records, capacity_bytes=capacity_bytes, rank=rank, world=world,
                     index_base=index_base, debug_full_frag_sort=int(full_frag_sort), debug_legacy_join=int(legacy_join))
        _check(lib().oge_gpu_dedup_create(C.byref(cfg), C.byref(self._h)))
        self.world = max(1, int(world))
        self.n = 0
        self.nbytes = 0

    def close(self):
        if self._h:
            lib().oge_gpu_dedup_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_header(self, text: str):
        """Resolve @RG ID -> LB -> library id on the host (openge_b200.header) and upload the table."""
        rg_ids, lib_ids, unknown, n_libs = _header.library_table(text)
        ids = (C.c_char_p * max(1, len(rg_ids)))(*rg_ids)
        libs = np.asarray(lib_ids if lib_ids else [0], dtype=np.int16)
        _check(lib().oge_gpu_dedup_set_readgroups(self._h, ids, libs.ctypes.data, len(rg_ids), unknown, n_libs))

    def push(self, records, offsets):
        """records: uint8 array (or PinnedBuffer.array); offsets: u64 n+1, relative to records[0]."""
        records = np.ascontiguousarray(records, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nrec = len(offsets) - 1
        _check(lib().oge_gpu_dedup_push(self._h, records.ctypes.data, records.nbytes, offsets.ctypes.data, nrec))
        _check(lib().oge_gpu_dedup_sync(self._h))      # numpy buffers may go away
        self.n += nrec
        self.nbytes += records.nbytes

    def sort(self):
        """ReadSorter in front of MarkDuplicates (`openge mergesort -M`): coordinate-sort the resident records on the device.
        Afterwards run / flags / pull refer to the sorted order; sort_order() gives the permutation."""
        _check(lib().oge_gpu_dedup_sort(self._h))

    def sort_order(self) -> np.ndarray:
        out = np.zeros(max(1, self.n), dtype=np.uint32)
        _check(lib().oge_gpu_dedup_sort_order(self._h, out.ctypes.data, self.n))
        return out[: self.n]

    def sort_stats(self) -> dict:
        a, b, c, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_float()
        _check(lib().oge_gpu_dedup_sort_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(ms)))
        return {"tied_records": a.value, "refinement_rounds": b.value, "launches": c.value, "ms": ms.value}

    def push_async(self, records_ptr, nbytes, offsets_ptr, nrec):
        _check(lib().oge_gpu_dedup_push(self._h, records_ptr, nbytes, offsets_ptr, nrec))
        self.n += nrec
        self.nbytes += nbytes

    def push_bgzf(self, comp_ptr, comp_bytes, in_off_ptr, csize_ptr, isize_ptr, n_blocks, header_bytes, host_copy_ptr=None):
        """Inflate a whole BGZF file on the device (decompress engine, or a kernel); records land in HBM, and in host_copy."""
        _check(lib().oge_gpu_dedup_push_bgzf(self._h, comp_ptr, comp_bytes, in_off_ptr, csize_ptr, isize_ptr, n_blocks, header_bytes, host_copy_ptr))

    def set_offsets(self, offsets_ptr, nrec, nbytes):
        _check(lib().oge_gpu_dedup_set_offsets(self._h, offsets_ptr, nrec))
        self.n = nrec
        self.nbytes = nbytes

    def frame(self, nbytes=0) -> int:
        """Frame the inflated records on the device (speculative parallel chain walk + proof) -> record count."""
        n = C.c_uint64()
        _check(lib().oge_gpu_dedup_frame(self._h, C.byref(n)))
        self.n = int(n.value)
        self.nbytes = nbytes
        return self.n

    def offsets(self) -> np.ndarray:
        out = np.empty(self.n + 1, dtype=np.uint64)
        _check(lib().oge_gpu_dedup_offsets(self._h, out.ctypes.data, self.n + 1))
        return out

    def sync(self):
        _check(lib().oge_gpu_dedup_sync(self._h))

    def reset(self):
        _check(lib().oge_gpu_dedup_reset(self._h))
        self.n = 0
        self.nbytes = 0

    def run(self):
        _check(lib().oge_gpu_dedup_run(self._h))

    def flags(self, out=None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n, dtype=np.uint16)
        _check(lib().oge_gpu_dedup_flags(self._h, out.ctypes.data, self.n))
        return out

    def pull(self):
        """-> (records uint8, offsets u64) after the flag rewrite (and -r compaction)."""
        rec = np.empty(self.nbytes, dtype=np.uint8)
        off = np.empty(self.n + 1, dtype=np.uint64)
        nb, nr = C.c_uint64(), C.c_uint64()
        _check(lib().oge_gpu_dedup_pull(self._h, rec.ctypes.data, rec.nbytes, off.ctypes.data, len(off), C.byref(nb), C.byref(nr)))
        if nr.value == 0:
            off[0] = 0
        return rec[: nb.value], off[: nr.value + 1]

    def deflate(self):
        """The output records (flag-patched, bins recomputed, -r applied) compressed into BGZF members on the device
        -> (member bytes as uint8, blocks, records).  Identical to the reference's output after decompression."""
        nb, nblk, nr = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().oge_gpu_dedup_deflate(self._h, C.byref(nb), C.byref(nblk), C.byref(nr)))
        out = np.empty(nb.value, dtype=np.uint8)
        _check(lib().oge_gpu_dedup_pull_bgzf(self._h, out.ctypes.data, out.nbytes))
        return out, int(nblk.value), int(nr.value)

    def stats(self) -> dict:
        st = Stats()
        _check(lib().oge_gpu_dedup_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def flagstats(self) -> dict:
        """The reference's Statistics counters (algorithms/statistics.cpp:77-162) over the records and their flag
        words after the run, reduced on the device."""
        out = np.zeros(len(FLAGSTAT_FIELDS), dtype=np.uint64)
        _check(lib().oge_gpu_dedup_flagstats(self._h, out.ctypes.data))
        return dict(zip(FLAGSTAT_FIELDS, (int(x) for x in out)))

    def ends(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=END_DTYPE)
        _check(lib().oge_gpu_dedup_debug_ends(self._h, out.ctypes.data, self.n))
        return out

    # ---- range sharding (include/oge_gpu_dedup.h, oge_gpu_shard_*): device pointers in and out
    def shard_setup(self, global_n, bases, split_ref, split_pos):
        b = np.ascontiguousarray(bases, dtype=np.uint64)
        r = np.ascontiguousarray(split_ref if len(split_ref) else [0], dtype=np.int32)
        p = np.ascontiguousarray(split_pos if len(split_pos) else [0], dtype=np.int32)
        _check(lib().oge_gpu_shard_setup(self._h, int(global_n), b.ctypes.data, r.ctypes.data, p.ctypes.data))

    # ---- range sharding: every output list is ordered by destination rank -> (device pointer, [count per rank])
    def _counts(self):
        return (C.c_uint64 * max(1, self.world))()

    def shard_key_bytes(self) -> int:
        v = C.c_uint32()
        _check(lib().oge_gpu_shard_key_bytes(self._h, C.byref(v)))
        return int(v.value)

    def shard_set_entry_bytes(self, entry_bytes: int):
        _check(lib().oge_gpu_shard_set_entry_bytes(self._h, int(entry_bytes)))
        self.entry_bytes = int(entry_bytes)

    @staticmethod
    def shard_comm_id() -> bytes:
        """128 bytes of a fresh NCCL unique id (one rank calls this and hands the bytes to the others)."""
        buf = (C.c_uint8 * 128)()
        _check(lib().oge_gpu_shard_comm_id(buf))
        return bytes(buf)

    def shard_comm_init(self, comm_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(comm_id)
        _check(lib().oge_gpu_shard_comm_init(self._h, buf))

    def shard_step(self) -> dict:
        """One whole sharded run driven from C++: the five phases with the four NCCL all-to-all exchanges between them."""
        info = StepInfo()
        _check(lib().oge_gpu_shard_step(self._h, C.byref(info)))
        return {"published": int(info.published_in), "routed": int(info.routed_in), "marks": int(info.marks_in),
                "exchanges": int(info.exchanges), "bytes_sent": int(info.bytes_sent)}

    def shard_begin(self):
        """-> (published entries by name owner, (hash pointer, count), routed fragment ends by key owner)"""
        p1, c1, ph, nh, p2, c2 = C.c_void_p(), self._counts(), C.c_void_p(), C.c_uint64(), C.c_void_p(), self._counts()
        _check(lib().oge_gpu_shard_begin(self._h, C.byref(p1), c1, C.byref(ph), C.byref(nh), C.byref(p2), c2))
        return (p1.value or 0, list(c1)), (ph.value or 0, int(nh.value)), (p2.value or 0, list(c2))

    def shard_probe(self, hash_ptr, n_hash, fr_ptr, n_fr):
        """-> (published round 2 by name owner, routed pair ends by key owner)"""
        p1, c1, p2, c2 = C.c_void_p(), self._counts(), C.c_void_p(), self._counts()
        _check(lib().oge_gpu_shard_probe(self._h, hash_ptr, n_hash, fr_ptr, n_fr, C.byref(p1), c1, C.byref(p2), c2))
        return (p1.value or 0, list(c1)), (p2.value or 0, list(c2))

    def shard_replay(self, pub_ptr, n_pub):
        """-> pair ends the replay formed for other ranks' key ranges, by key owner"""
        p1, c1 = C.c_void_p(), self._counts()
        _check(lib().oge_gpu_shard_replay(self._h, pub_ptr, n_pub, C.byref(p1), c1))
        return p1.value or 0, list(c1)

    def shard_finish(self, pr_ptr, n_pr):
        """-> marks by record owner"""
        p1, c1 = C.c_void_p(), self._counts()
        _check(lib().oge_gpu_shard_finish(self._h, pr_ptr, n_pr, C.byref(p1), c1))
        return p1.value or 0, list(c1)

    def copy_d2d(self, dst_ptr, src_ptr, nbytes):
        _check(lib().oge_gpu_copy_d2d(self._h, dst_ptr, src_ptr, nbytes))

    def shard_apply(self, ptr, n):
        _check(lib().oge_gpu_shard_apply(self._h, ptr, n))

    def device_ptrs(self):
        r, o, f = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().oge_gpu_dedup_device_ptrs(self._h, C.byref(r), C.byref(o), C.byref(f)))
        return r.value, o.value, f.value


def context_for(bam: BamFile, **kw) -> DedupContext:
    """A context whose key layout is sized from the BAM's reference dictionary."""
    max_len = max([l for _, l in bam.refs], default=0)
    ctx = DedupContext(n_ref=len(bam.refs), max_ref_len=max_len, **kw)
    ctx.set_header(bam.text)
    return ctx


class MarkDuplicates:
    """Drop-in for the reference's MarkDuplicates module (mark_duplicates.h:27-68) over whole BamFiles."""

    def __init__(self, temp_directory: str = "/tmp", device: int = 0):
        self.temp_directory = temp_directory      # kept for signature parity; nothing is spilled
        self.removeDuplicates = False             # mark_duplicates.h:41
        self.verbose = True                       # AlgorithmModule::verbose; False reproduces SURVEY F1
        self.device = device
        self.last_stats = None

    def run(self, bam: BamFile) -> BamFile:
        with context_for(bam, device=self.device, remove_duplicates=self.removeDuplicates,
                         compat_quiet_index_bug=not self.verbose) as ctx:
            ctx.push(bam.records, bam.offsets)
            ctx.run()
            rec, off = ctx.pull()
            self.last_stats = ctx.stats()
        return BamFile(text=bam.text, refs=list(bam.refs), records=rec, offsets=off)

    def flags(self, bam: BamFile) -> np.ndarray:
        with context_for(bam, device=self.device, compat_quiet_index_bug=not self.verbose) as ctx:
            ctx.push(bam.records, bam.offsets)
            ctx.run()
            out = ctx.flags()
            self.last_stats = ctx.stats()
        return out
