"""GPU suite: the output file's BGZF blocks made on the device (oge_gpu_dedup_deflate: bins, -r, one warp per block
deflate + CRC-32, packing).  The bar is "identical to the reference's output after decompression": every member must pass
gzip's checks (valid deflate stream by zlib's verdict, CRC-32, ISIZE), the inflated file must equal the inflated file of the
byte-identical host writer (whose compressed bytes are pinned to the compiled reference's, tests/test_gpu_fused.py), and the
compiled reference must read the file."""
import os
import subprocess
import tempfile
import zlib

import numpy as np
import pytest

import oracle
from openge_b200 import _build, bamhost, bamio, dedup, synth

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tmp():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        yield d


def inflate_members(z: bytes) -> bytes:
    """What gzip.decompress does for a multi-member file (which is quadratic in the number of members): every member's
    deflate stream inflated by zlib, its CRC-32 and ISIZE checked."""
    out, pos = [], 0
    while pos < len(z):
        bs = int.from_bytes(z[pos + 16:pos + 18], "little") + 1
        d = zlib.decompressobj(-15)
        raw = d.decompress(z[pos + 18:pos + bs - 8])
        assert d.eof and d.unused_data == b""
        assert zlib.crc32(raw) == int.from_bytes(z[pos + bs - 8:pos + bs - 4], "little")
        assert len(raw) == int.from_bytes(z[pos + bs - 4:pos + bs], "little")
        out.append(raw)
        pos += bs
    return b"".join(out)


def members_of(z: bytes):
    """-> [(csize, isize)] of a BGZF file, walking the member headers."""
    out, pos = [], 0
    while pos < len(z):
        assert z[pos:pos + 4] == b"\x1f\x8b\x08\x04" and z[pos + 12:pos + 16] == b"BC\x02\x00"
        bs = int.from_bytes(z[pos + 16:pos + 18], "little") + 1
        out.append((bs, int.from_bytes(z[pos + bs - 4:pos + bs], "little")))
        pos += bs
    assert pos == len(z)
    return out


@pytest.mark.parametrize("name,scale,seed,remove", [("C1", 0.02, 3, False), ("C3", 0.01, 99, False), ("C3", 0.01, 99, True),
                                                    ("C4", 0.01, 6, False), ("C4", 0.01, 6, True), ("C5", 0.001, 11, False)])
def test_device_made_file_equals_the_host_made_file_after_decompression(tmp, name, scale, seed, remove):
    bam = synth.make(name, scale, seed=seed)
    inp, out_host, out_dev = (os.path.join(tmp, f) for f in ("in.bam", "host.bam", "dev.bam"))
    bamio.write_bam(inp, bam)
    bamhost.dedup_file(inp, out_host, remove_duplicates=remove, level=1, pg_command_line="openge dedup x")
    st = bamhost.dedup_file(inp, out_dev, remove_duplicates=remove, level=1, pg_command_line="openge dedup x", gpu_deflate=True)
    assert st["gpu_deflate"] and st["dedup"]["deflate_blocks"] > 0 and st["dedup"]["ms_deflate"] > 0
    z_host, z_dev = open(out_host, "rb").read(), open(out_dev, "rb").read()
    want = inflate_members(z_host)
    got = inflate_members(z_dev)      # checks every member: valid deflate stream, CRC-32, ISIZE
    assert got == want
    ms = members_of(z_dev)
    assert ms[-1] == (28, 0)                                     # the end-of-file member
    assert all(i <= 65536 and c <= 65536 for c, i in ms)
    assert sum(i == 65280 for _, i in ms) >= len(ms) - 4         # record blocks are 65280 bytes, the last one and the header aside
    # compression in zlib level 1's class
    assert len(z_dev) < 1.05 * len(z_host)
    # flags inside equal the oracle's
    b2 = bamio.read_bam(out_dev)
    f = oracle.markdup(bam.records, bam.offsets, bam.text)
    keep = (f & 0x400) == 0 if remove else np.ones(bam.n, bool)
    assert b2.n == int(keep.sum()) == st["n_out"]
    o = b2.offsets[:-1].astype(np.int64)
    assert np.array_equal(b2.records[o + 18].astype(np.uint16) | (b2.records[o + 19].astype(np.uint16) << 8), f[keep])


def test_block_edges_and_tiny_files(tmp):
    # records ending exactly on a block boundary, one record, no record
    rec_len = None
    for n_target in ("exact", 1, 0):
        bam = synth.make("C1", 0.004, seed=5)
        if n_target == "exact":
            sizes = np.diff(bam.offsets.astype(np.int64))
            # the longest prefix whose byte length is a multiple of ... not controllable record by record: take the prefix that ends
            # closest below 3 * 65280 and pad nothing -- the edge that matters is "last block shorter than a word"
            k = int(np.searchsorted(bam.offsets, 3 * 65280)) - 1
        else:
            k = n_target
        sub = bamio.BamFile(text=bam.text, refs=bam.refs, records=bam.records[:int(bam.offsets[k])], offsets=bam.offsets[:k + 1])
        inp, out_host, out_dev = (os.path.join(tmp, "%s_%s" % (n_target, f)) for f in ("in.bam", "host.bam", "dev.bam"))
        bamio.write_bam(inp, sub)
        bamhost.dedup_file(inp, out_host, level=6)
        st = bamhost.dedup_file(inp, out_dev, level=6, gpu_deflate=True)
        assert inflate_members(open(out_dev, "rb").read()) == inflate_members(open(out_host, "rb").read())
        assert st["n_out"] == k


def test_incompressible_records_are_stored(tmp):
    # qualities and bases from all 256 byte values: dynamic Huffman cannot shrink such blocks, the kernel must fall back to
    # stored blocks that still fit a BGZF member
    bam = synth.make("C1", 0.01, seed=8)
    rec = bam.records.copy()
    rng = np.random.default_rng(0)
    off = bam.offsets.astype(np.int64)
    for i in range(bam.n):
        p = off[i]
        l_name, n_cig = int(rec[p + 12]), int(rec[p + 16]) | (int(rec[p + 17]) << 8)
        l_seq = int.from_bytes(rec[p + 20:p + 24].tobytes(), "little")
        a = p + 36 + l_name + 4 * n_cig
        b = a + (l_seq + 1) // 2 + l_seq
        rec[a:b] = rng.integers(0, 256, b - a, dtype=np.uint8)
    noisy = bamio.BamFile(text=bam.text, refs=bam.refs, records=rec, offsets=bam.offsets)
    inp, out_host, out_dev = (os.path.join(tmp, f) for f in ("in.bam", "host.bam", "dev.bam"))
    bamio.write_bam(inp, noisy)
    bamhost.dedup_file(inp, out_host, level=1)
    bamhost.dedup_file(inp, out_dev, level=1, gpu_deflate=True)
    assert inflate_members(open(out_dev, "rb").read()) == inflate_members(open(out_host, "rb").read())
    assert all(c <= 65536 for c, _ in members_of(open(out_dev, "rb").read()))


def test_the_compiled_reference_reads_the_device_made_file(tmp):
    if not oracle.ref_available():
        pytest.skip("compiled reference not present")
    bam = synth.make("C3", 0.01, seed=41)
    inp, out_dev = os.path.join(tmp, "in.bam"), os.path.join(tmp, "dev.bam")
    bamio.write_bam(inp, bam)
    bamhost.dedup_file(inp, out_dev, level=1, gpu_deflate=True)
    # the reference dedups the already-marked file: its reader (BgzfInputStream + BamDeserializer) must accept every block,
    # and marking is idempotent on marked input
    again = os.path.join(tmp, "again.bam")
    r = None
    for _ in range(4):      # the reference's pipeline occasionally does not terminate (tests/test_bamhost.py does the same)
        try:
            r = subprocess.run([_build.REF_BIN, "-T", tmp, "--nosplit", "-v", "-c", "1", out_dev, again], capture_output=True, timeout=120)
            break
        except subprocess.TimeoutExpired:
            r = None
    if r is None:
        pytest.skip("reference did not terminate")
    assert r.returncode == 0, r.stderr.decode()
    a, b = bamio.read_bam(out_dev), bamio.read_bam(again)
    assert a.n == b.n == bam.n and np.array_equal(a.records, b.records)


def test_fused_binary_with_gpu_deflate(tmp):
    bam = synth.make("C4", 0.005, seed=2)
    inp = os.path.join(tmp, "in.bam")
    bamio.write_bam(inp, bam)
    exe = _build.FUSED_BIN
    outs = []
    for extra in ([], ["--gpu-deflate"], ["--gpu-deflate", "-r"], ["-r"]):
        out = os.path.join(tmp, "o%d.bam" % len(outs))
        r = subprocess.run([exe, "dedup", inp, "-o", out, "-v", "-c", "1"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        assert (b"gpu deflate:" in r.stderr) == ("--gpu-deflate" in extra)
        outs.append(inflate_members(open(out, "rb").read()))
    # same command line apart from the flag: the @PG line differs by it, so compare from the first record on
    def records(raw):
        l_text = int.from_bytes(raw[4:8], "little")
        p = 8 + l_text
        n_ref = int.from_bytes(raw[p:p + 4], "little")
        p += 4
        for _ in range(n_ref):
            l = int.from_bytes(raw[p:p + 4], "little")
            p += 4 + l + 4
        return raw[p:]
    assert records(outs[0]) == records(outs[1]) and records(outs[2]) == records(outs[3]) and len(records(outs[2])) < len(records(outs[0]))
