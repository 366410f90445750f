"""ctypes binding of liboge_bamhost.so (include/oge_bam_host.h): the host-side BAM streaming layer around the
GPU dedup path -- parallel BGZF inflate into one buffer, header model, record framing, flag/bin rewrite, parallel
BGZF deflate with the reference's exact block layout.  No CUDA in here; the duplicate flags come from
openge_b200.dedup (libopenge_b200.so).

``dedup_file`` is `openge dedup in.bam -o out.bam` (reference: commands/command_dedup.cpp:37-114) as one fused
path: load -> push/run/flags on the GPU -> apply_flags -> store.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

ERRORS = {-1: "OGE_BAM_ERR_IO", -2: "OGE_BAM_ERR_FORMAT", -3: "OGE_BAM_ERR_NOMEM", -4: "OGE_BAM_ERR_ARG"}

EXPORTS = ["oge_bam_load", "oge_bam_open_bgzf", "oge_bam_bgzf_index", "oge_bam_records_buffer", "oge_bam_frame_records", "oge_bam_adopt_offsets", "oge_bam_close", "oge_bam_header_text", "oge_bam_n_ref", "oge_bam_ref_name", "oge_bam_ref_len",
           "oge_bam_records", "oge_bam_records_bytes", "oge_bam_offsets", "oge_bam_n_records", "oge_bam_library_table",
           "oge_bam_apply_flags", "oge_bam_store", "oge_bam_store_members", "oge_bam_store_members_stream", "oge_bam_timings", "oge_bgzf_decompress", "oge_bgzf_compress",
           "oge_bam_header_render", "oge_bam_buffer_free", "oge_bam_last_error", "oge_bam_set_sort_order"]


class BamHostError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "error"), code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.ensure_bamhost()
        if not os.path.exists(path):
            raise ImportError("liboge_bamhost.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        vp, u64 = C.c_void_p, C.c_uint64
        L.oge_bam_load.argtypes = [C.c_char_p, C.c_int, vp, vp, C.POINTER(vp)]
        L.oge_bam_open_bgzf.argtypes = [C.c_char_p, C.c_int, vp, vp, C.POINTER(vp)]
        L.oge_bam_bgzf_index.argtypes = [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]
        L.oge_bam_records_buffer.argtypes = [vp]
        L.oge_bam_records_buffer.restype = vp
        L.oge_bam_frame_records.argtypes = [vp]
        L.oge_bam_adopt_offsets.argtypes = [vp, vp, u64]
        L.oge_bam_set_sort_order.argtypes = [vp, C.c_char_p]
        L.oge_bam_close.argtypes = [vp]
        L.oge_bam_close.restype = None
        L.oge_bam_header_text.argtypes = [vp]
        L.oge_bam_header_text.restype = C.c_char_p
        L.oge_bam_n_ref.argtypes = [vp]
        L.oge_bam_n_ref.restype = C.c_int32
        L.oge_bam_ref_name.argtypes = [vp, C.c_int32]
        L.oge_bam_ref_name.restype = C.c_char_p
        L.oge_bam_ref_len.argtypes = [vp, C.c_int32]
        L.oge_bam_ref_len.restype = C.c_int32
        L.oge_bam_records.argtypes = [vp]
        L.oge_bam_records.restype = vp
        L.oge_bam_records_bytes.argtypes = [vp]
        L.oge_bam_records_bytes.restype = u64
        L.oge_bam_offsets.argtypes = [vp]
        L.oge_bam_offsets.restype = vp
        L.oge_bam_n_records.argtypes = [vp]
        L.oge_bam_n_records.restype = u64
        L.oge_bam_library_table.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int32), C.POINTER(C.c_int16), C.POINTER(C.c_int32)]
        L.oge_bam_apply_flags.argtypes = [vp, vp, C.c_int, C.c_int]
        L.oge_bam_store.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
        L.oge_bam_store_members.argtypes = [vp, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, vp, C.c_uint64]
        L.oge_bam_timings.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
        L.oge_bgzf_decompress.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.oge_bgzf_compress.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.oge_bam_header_render.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.oge_bam_buffer_free.argtypes = [vp]
        L.oge_bam_buffer_free.restype = None
        L.oge_bam_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(L, name)
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise BamHostError(rc, lib().oge_bam_last_error().decode("utf-8", "replace"))


def _take(ptr, n) -> bytes:
    try:
        return bytes((C.c_char * n).from_address(ptr.value)) if n else b""      # (string_at takes a C int: 2 GB at most)
    finally:
        lib().oge_bam_buffer_free(ptr)


def bgzf_decompress(data: bytes, threads: int = 0) -> bytes:
    out, n = C.c_void_p(), C.c_size_t()
    buf = np.frombuffer(data, dtype=np.uint8)
    _check(lib().oge_bgzf_decompress(buf.ctypes.data if len(buf) else None, len(buf), threads, C.byref(out), C.byref(n)))
    return _take(out, n.value)


def bgzf_compress(raw: bytes, level: int = 6, threads: int = 0) -> bytes:
    """The reference's BgzfOutputStream (util/bgzf_output_stream.cpp) block for block, compressed in parallel."""
    out, n = C.c_void_p(), C.c_size_t()
    buf = np.frombuffer(raw, dtype=np.uint8)
    _check(lib().oge_bgzf_compress(buf.ctypes.data if len(buf) else None, len(buf), level, threads, C.byref(out), C.byref(n)))
    return _take(out, n.value)


def header_render(text: str) -> str:
    """BamHeader(text).toString() of the reference (util/bam_header.cpp:107-262)."""
    out = C.c_void_p()
    _check(lib().oge_bam_header_render(text.encode(), C.byref(out)))
    try:
        return C.string_at(out.value).decode()
    finally:
        lib().oge_bam_buffer_free(out)


class HostBam:
    """One loaded BAM file (oge_bam_file): inflated, framed, header parsed."""

    def __init__(self, path: str, threads: int = 0, pinned: bool = False, defer_inflate: bool = False):
        """defer_inflate: two-stage open (oge_bam_open_bgzf) -- only the header is inflated here; the caller inflates
        the blocks of bgzf_index() into records_buffer() (the GPU does: DedupContext.push_bgzf) and calls frame_records()."""
        self._h = C.c_void_p()
        alloc = free = None
        if pinned:      # the inflated stream lands in page-locked memory so that push() is one DMA
            from . import dedup
            g = dedup.lib()
            alloc = C.cast(g.oge_gpu_host_alloc, C.c_void_p)
            free = C.cast(g.oge_gpu_host_free, C.c_void_p)
        fn = lib().oge_bam_open_bgzf if defer_inflate else lib().oge_bam_load
        _check(fn(os.fsencode(path), threads, alloc, free, C.byref(self._h)))

    def bgzf_index(self) -> dict:
        comp, nb, io, cs, isz, n, hb = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(lib().oge_bam_bgzf_index(self._h, C.byref(comp), C.byref(nb), C.byref(io), C.byref(cs), C.byref(isz), C.byref(n), C.byref(hb)))
        return {"comp": comp.value, "comp_bytes": int(nb.value), "in_off": io.value, "csize": cs.value, "isize": isz.value,
                "n_blocks": int(n.value), "header_bytes": int(hb.value)}

    def records_buffer(self) -> int:
        p = lib().oge_bam_records_buffer(self._h)
        if not p:
            _check(-3)
        return p

    def frame_records(self):
        _check(lib().oge_bam_frame_records(self._h))

    def adopt_offsets(self, offsets: np.ndarray):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(lib().oge_bam_adopt_offsets(self._h, offsets.ctypes.data, len(offsets) - 1))

    def set_sort_order(self, so: str):
        """@HD SO of the stored file (what ReadSorter does to the header it hands on, read_sorter.cpp:256-258)."""
        _check(lib().oge_bam_set_sort_order(self._h, so.encode()))

    def close(self):
        if self._h:
            lib().oge_bam_close(self._h)
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

    @property
    def n(self) -> int:
        return int(lib().oge_bam_n_records(self._h))

    @property
    def text(self) -> str:
        return lib().oge_bam_header_text(self._h).decode()

    @property
    def refs(self):
        L = lib()
        return [(L.oge_bam_ref_name(self._h, i).decode(), int(L.oge_bam_ref_len(self._h, i))) for i in range(L.oge_bam_n_ref(self._h))]

    @property
    def records(self) -> np.ndarray:
        """A VIEW of the resident record bytes (valid until close / apply_flags with remove_duplicates)."""
        nb = int(lib().oge_bam_records_bytes(self._h))
        if nb == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array((C.c_uint8 * nb).from_address(lib().oge_bam_records(self._h)))

    @property
    def offsets(self) -> np.ndarray:
        return np.ctypeslib.as_array((C.c_uint64 * (self.n + 1)).from_address(lib().oge_bam_offsets(self._h)))

    def records_ptr(self):
        return lib().oge_bam_records(self._h), int(lib().oge_bam_records_bytes(self._h)), lib().oge_bam_offsets(self._h)

    def library_table(self):
        ids, libs, n, unk, nl = C.c_void_p(), C.c_void_p(), C.c_int32(), C.c_int16(), C.c_int32()
        _check(lib().oge_bam_library_table(self._h, C.byref(ids), C.byref(libs), C.byref(n), C.byref(unk), C.byref(nl)))
        names = [C.cast(ids.value, C.POINTER(C.c_char_p))[i] for i in range(n.value)]
        lib_ids = [C.cast(libs.value, C.POINTER(C.c_int16))[i] for i in range(n.value)]
        return names, lib_ids, int(unk.value), int(nl.value)

    def apply_flags(self, flags: np.ndarray, remove_duplicates: bool = False, threads: int = 0):
        flags = np.ascontiguousarray(flags, dtype=np.uint16)
        if len(flags) != self.n:
            raise ValueError("flags: %d values for %d records" % (len(flags), self.n))
        _check(lib().oge_bam_apply_flags(self._h, flags.ctypes.data, int(remove_duplicates), threads))

    def store(self, path: str, format: str | None = None, level: int = 6, pg_command_line: str | None = None,
              pg_version: str = "0.3-b200", threads: int = 0):
        _check(lib().oge_bam_store(self._h, os.fsencode(path), format.encode() if format else None, level,
                                   pg_command_line.encode() if pg_command_line else None, pg_version.encode(), threads))

    def store_members(self, path: str, members: np.ndarray, level: int = 6, pg_command_line: str | None = None,
                      pg_version: str = "0.3-b200"):
        """The output file around record members that are already compressed (DedupContext.deflate)."""
        m = np.ascontiguousarray(members, dtype=np.uint8)
        _check(lib().oge_bam_store_members(self._h, os.fsencode(path), level, pg_command_line.encode() if pg_command_line else None,
                                           pg_version.encode(), m.ctypes.data if m.nbytes else None, m.nbytes))

    def timings(self) -> dict:
        t = (C.c_double * 6)()
        _check(lib().oge_bam_timings(self._h, t, 6))
        return dict(zip(("read", "scan", "inflate", "frame", "apply_flags", "store"), (float(x) for x in t)))


def dedup_file(in_path: str, out_path: str, remove_duplicates: bool = False, level: int = 6, format: str | None = None,
               pg_command_line: str | None = None, threads: int = 0, device: int = 0, gpu_inflate: bool = True,
               pinned: bool = False, sort: bool = False, gpu_deflate: bool = False) -> dict:
    """`openge dedup in.bam -o out.bam` on the GPU, file to file.  -> stats (dedup counters, flag statistics, timings).
    gpu_inflate: the BGZF blocks are inflated on the device (DedupContext.push_bgzf), else by the host threads.
    sort: coordinate sort on the device in front of the dedup (`openge mergesort -M`).
    gpu_deflate: the output's BGZF blocks are made on the device as well (DedupContext.deflate): the records never come back
    uncompressed; the file equals the reference's after decompression (not byte for byte, which the default gives)."""
    from . import dedup
    with open(in_path, "rb") as fh:
        is_bgzf = fh.read(2) == b"\x1f\x8b"
    gpu_inflate = gpu_inflate and is_bgzf
    with HostBam(in_path, threads=threads, pinned=pinned, defer_inflate=gpu_inflate) as bam:
        refs = bam.refs
        gpu_deflate = gpu_deflate and (format in (None, "bam"))
        ctx = dedup.DedupContext(n_ref=len(refs), max_ref_len=max([l for _, l in refs], default=0), device=device,
                                 remove_duplicates=bool(remove_duplicates and gpu_deflate))
        with ctx:      # without gpu_deflate -r is applied by the host layer (apply_flags), so pull() returns every record
            ctx.set_header(bam.text)
            if gpu_inflate:
                # compressed bytes up, inflate + framing + dedup on the device, ONE copy of the (flag-patched) records back
                ix = bam.bgzf_index()
                ctx.push_bgzf(ix["comp"], ix["comp_bytes"], ix["in_off"], ix["csize"], ix["isize"], ix["n_blocks"], ix["header_bytes"], None)
                total = int(np.ctypeslib.as_array((C.c_uint32 * ix["n_blocks"]).from_address(ix["isize"])).sum()) if ix["n_blocks"] else 0
                ctx.frame(total - ix["header_bytes"])
                if sort:
                    ctx.sort()
                ctx.run()
                flags = ctx.flags()
                if not gpu_deflate:
                    nb, nr = C.c_uint64(), C.c_uint64()
                    offs = np.empty(ctx.n + 1, dtype=np.uint64)
                    from .dedup import _check as _gcheck, lib as _glib
                    _gcheck(_glib().oge_gpu_dedup_pull(ctx._h, bam.records_buffer(), ctx.nbytes, offs.ctypes.data, len(offs), C.byref(nb), C.byref(nr)))
                    bam.adopt_offsets(offs)
            else:
                ptr, nbytes, off_ptr = bam.records_ptr()
                ctx.push_async(ptr, nbytes, off_ptr, bam.n)
                if sort:
                    ctx.sort()
                ctx.run()
                flags = ctx.flags()
                if sort and bam.n and not gpu_deflate:      # the records come back in their new order
                    nb, nr = C.c_uint64(), C.c_uint64()
                    offs = np.empty(bam.n + 1, dtype=np.uint64)
                    from .dedup import _check as _gcheck, lib as _glib
                    _gcheck(_glib().oge_gpu_dedup_pull(ctx._h, ptr, nbytes, offs.ctypes.data, len(offs), C.byref(nb), C.byref(nr)))
                    bam.adopt_offsets(offs)
            members = None
            if gpu_deflate:
                members, _, n_out = ctx.deflate()
            out = {"dedup": ctx.stats(), "flagstats": ctx.flagstats(), "gpu_inflate": gpu_inflate, "gpu_deflate": gpu_deflate}
            if sort:
                bam.set_sort_order("coordinate")
                out["sort"] = ctx.sort_stats()
        if gpu_deflate:
            bam.store_members(out_path, members, level, pg_command_line)
            out["timings"] = bam.timings()
            out["n_out"] = n_out
        else:
            bam.apply_flags(flags, remove_duplicates, threads)
            bam.store(out_path, format, level, pg_command_line, threads=threads)
            out["timings"] = bam.timings()
            out["n_out"] = bam.n
    return out
