"""GPU suite: BGZF inflate on the device (oge_gpu_dedup_push_bgzf, one warp per block) against zlib, and the fused
file path with it against the reference's output files."""
import ctypes as C
import hashlib
import os
import subprocess
import sys
import tempfile
import zlib

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


def bgzf_blocks(raw: bytes, level: int, strategy=zlib.Z_DEFAULT_STRATEGY, block=65536) -> bytes:
    """A BGZF file cut into `block`-byte payloads, every block compressed with the given zlib level and strategy."""
    out = []
    for k in list(range(0, len(raw), block)) + [len(raw)]:
        chunk = raw[k:k + block] if k < len(raw) else b""
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        z = co.compress(chunk) + co.flush()
        if len(z) + 26 > 65536:      # incompressible: store it
            co = zlib.compressobj(0, zlib.DEFLATED, -15)
            z = co.compress(chunk) + co.flush()
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + (len(z) + 25).to_bytes(2, "little") + z +
                   zlib.crc32(chunk).to_bytes(4, "little") + len(chunk).to_bytes(4, "little"))
    return b"".join(out)


def gpu_inflate_file(path):
    """-> (records bytes as inflated on the device and copied back, HostBam) via the two-stage open."""
    h = bamhost.HostBam(path, defer_inflate=True)
    refs = h.refs
    ctx = dedup.DedupContext(n_ref=len(refs), max_ref_len=max([l for _, l in refs], default=0))
    ix = h.bgzf_index()
    ctx.push_bgzf(ix["comp"], ix["comp_bytes"], ix["in_off"], ix["csize"], ix["isize"], ix["n_blocks"], ix["header_bytes"], h.records_buffer())
    h.frame_records()
    return ctx, h


@pytest.fixture(params=["engine", "warp", "threads"])
def inflate_kernel(request):
    """All three decoders: the hardware decompress engine (default on the B200), one warp per block, one thread per block."""
    dedup.set_inflate_kernel(request.param)
    if dedup.inflate_kernel() != request.param:
        dedup.set_inflate_kernel("default")
        pytest.skip("this device / driver has no hardware decompress engine")
    yield request.param
    dedup.set_inflate_kernel("default")
    dedup.set_bgzf_chunk_bytes(0)


@pytest.mark.parametrize("level,strategy,block", [(6, zlib.Z_DEFAULT_STRATEGY, 65536), (1, zlib.Z_DEFAULT_STRATEGY, 65280),
                                                  (9, zlib.Z_DEFAULT_STRATEGY, 30000), (0, zlib.Z_DEFAULT_STRATEGY, 65000),
                                                  (6, zlib.Z_FIXED, 40000), (6, zlib.Z_HUFFMAN_ONLY, 50000), (6, zlib.Z_RLE, 777)])
def test_device_inflate_matches_zlib(tmp, inflate_kernel, level, strategy, block):
    bam = synth.make("C3", 0.02, seed=31)
    raw = bamio.serialize_bam_stream(bam)
    p = os.path.join(tmp, "x.bam")
    open(p, "wb").write(bgzf_blocks(raw, level, strategy, block))
    ctx, h = gpu_inflate_file(p)
    with ctx, h:
        assert h.n == bam.n and h.text == bam.text
        assert np.array_equal(h.records, bam.records) and np.array_equal(h.offsets, bam.offsets)      # the D2H copy
        st = ctx.stats()
        assert st["inflate_bytes_out"] == len(raw) and st["ms_inflate"] > 0
        assert st["inflate_mode"] == dedup.INFLATE_KERNELS[inflate_kernel] and st["inflate_pieces"] == 1 and st["ms_push_bgzf"] > 0
        # and what stayed in HBM is the same: run the dedup on it
        ptr, nbytes, off_ptr = h.records_ptr()
        ctx.set_header(h.text)
        ctx.set_offsets(off_ptr, h.n, nbytes)
        ctx.run()
        assert np.array_equal(ctx.flags(), oracle.markdup(bam.records, bam.offsets, bam.text))
        rec, off = ctx.pull()
        want = bam.records.copy()
        f = ctx.flags()
        o = bam.offsets[:-1].astype(np.int64)
        want[o + 18] = (f & 0xFF).astype(np.uint8)
        want[o + 19] = (f >> 8).astype(np.uint8)
        assert np.array_equal(rec, want) and np.array_equal(off, bam.offsets)


def test_device_inflate_other_writers_and_large_headers(tmp, inflate_kernel):
    # python zlib at level 1 through bamio (its own block size), the host layer's writer, and a header spanning several blocks
    bam = synth.make("C1", 0.01, seed=8)
    big_text = "@HD\tVN:1.4\tSO:coordinate\n" + "".join("@SQ\tSN:contig%05d\tLN:%d\n" % (i, 1000 + i) for i in range(9000)) + "@RG\tID:rg1\tLB:lib1\n"
    refs = [("contig%05d" % i, 1000 + i) for i in range(9000)]
    b2 = bamio.BamFile(text=big_text, refs=refs, records=bam.records, offsets=bam.offsets)
    for k, (b, writer) in enumerate([(bam, "bamio"), (bam, "host"), (b2, "host")]):
        p = os.path.join(tmp, "w%d.bam" % k)
        if writer == "bamio":
            bamio.write_bam(p, b)
        else:
            open(p, "wb").write(bamhost.bgzf_compress(bamio.serialize_bam_stream(b), 6))
        ctx, h = gpu_inflate_file(p)
        with ctx, h:
            assert h.text == b.text and h.refs == b.refs
            assert np.array_equal(h.records, b.records) and np.array_equal(h.offsets, b.offsets)


@pytest.mark.parametrize("piece", [1, 150000, 1 << 20])
def test_device_inflate_in_pieces(tmp, inflate_kernel, piece):
    # the overlapped form of push_bgzf: upload, inflate and copy-back piece by piece on three streams (a piece of 1 byte =
    # one BGZF block per piece); the header spans several blocks, so the first pieces hold no record bytes
    bam = synth.make("C4", 0.01, seed=77)
    text = bam.text + "".join("@CO\tpadding line %06d to push the header over one BGZF block\n" % i for i in range(3000))
    b2 = bamio.BamFile(text=text, refs=bam.refs, records=bam.records, offsets=bam.offsets)
    raw = bamio.serialize_bam_stream(b2)
    p = os.path.join(tmp, "pieces.bam")
    open(p, "wb").write(bgzf_blocks(raw, 1, zlib.Z_DEFAULT_STRATEGY, 20000))
    dedup.set_bgzf_chunk_bytes(piece)
    # the file sits in pageable memory: with small staging buffers it goes up through them (several per piece, or several
    # pieces per buffer), else directly
    dedup.set_bgzf_staging(70000 if piece != 150000 else 32 << 20)
    try:
        ctx, h = gpu_inflate_file(p)
    finally:
        dedup.set_bgzf_staging()
    with ctx, h:
        st = ctx.stats()
        assert 3 < st["inflate_pieces"] <= -(-st["inflate_bytes_in"] // piece) and (piece > 1 or st["inflate_pieces"] == st["inflate_blocks"])
        assert h.n == bam.n and np.array_equal(h.records, bam.records) and np.array_equal(h.offsets, bam.offsets)      # the copy-back, piece by piece
        n = ctx.frame()      # and what stayed in HBM
        assert n == bam.n
        ctx.set_header(h.text)
        ctx.run()
        assert np.array_equal(ctx.flags(), oracle.markdup(bam.records, bam.offsets, bam.text))


def test_engine_reports_a_corrupt_stream_like_the_reference(tmp):
    # the hardware engine answers an invalid deflate stream with a launch failure that takes the CUDA context with it (the
    # reference exit(-1)s in the same place): checked in a process of its own; the library must still name the cause
    dedup.set_inflate_kernel("engine")
    have = dedup.inflate_kernel() == "engine"
    dedup.set_inflate_kernel("default")
    if not have:
        pytest.skip("this device / driver has no hardware decompress engine")
    bam = synth.make("C1", 0.01, seed=9)
    z = bytearray(bamhost.bgzf_compress(bamio.serialize_bam_stream(bam), 6))
    bs0 = int.from_bytes(z[16:18], "little") + 1
    for k in range(bs0 + 600, bs0 + 640):
        z[k] ^= 0xA5
    p = os.path.join(tmp, "bad.bam")
    open(p, "wb").write(bytes(z))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from openge_b200 import bamhost, dedup\n"
            "h = bamhost.HostBam(%r, defer_inflate=True)\n"
            "ctx = dedup.DedupContext(n_ref=len(h.refs), max_ref_len=100000000)\n"
            "ix = h.bgzf_index()\n"
            "try:\n"
            "    ctx.push_bgzf(ix['comp'], ix['comp_bytes'], ix['in_off'], ix['csize'], ix['isize'], ix['n_blocks'], ix['header_bytes'], None)\n"
            "except dedup.DedupError as e:\n"
            "    print('CODE', e.code, str(e)); sys.stdout.flush()\n"
            "    import os; os._exit(0)\n"
            "print('NO ERROR')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), p)
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300,
                       env=dict(os.environ, OGE_INFLATE_KERNEL="engine"))
    assert "CODE -7" in r.stdout and "Zlib inflate failed" in r.stdout, r.stdout


def test_device_inflate_rejects_corrupt_blocks(tmp, inflate_kernel):
    if inflate_kernel == "engine":
        pytest.skip("see test_engine_reports_a_corrupt_stream_like_the_reference: the engine's failure ends the CUDA context")
    bam = synth.make("C1", 0.01, seed=9)
    z = bytearray(bamhost.bgzf_compress(bamio.serialize_bam_stream(bam), 6))
    bs0 = int.from_bytes(z[16:18], "little") + 1
    for k in range(bs0 + 600, bs0 + 640):      # scramble the middle of the second block's deflate stream
        z[k] ^= 0xA5
    p = os.path.join(tmp, "bad.bam")
    open(p, "wb").write(bytes(z))
    with bamhost.HostBam(p, defer_inflate=True) as h, dedup.DedupContext(n_ref=len(h.refs), max_ref_len=100000000) as ctx:
        ix = h.bgzf_index()
        with pytest.raises(dedup.DedupError) as e:
            ctx.push_bgzf(ix["comp"], ix["comp_bytes"], ix["in_off"], ix["csize"], ix["isize"], ix["n_blocks"], ix["header_bytes"], h.records_buffer())
        assert e.value.code == -7 and "BGZF block 1" in str(e.value)
        # state errors: offsets before a successful push_bgzf; plain push after one
    with dedup.DedupContext(n_ref=1, max_ref_len=1000) as ctx:
        with pytest.raises(dedup.DedupError) as e:
            ctx.set_offsets(np.zeros(1, np.uint64).ctypes.data, 0, 0)
        assert e.value.code == -5


@pytest.mark.parametrize("name,scale,seed,level,remove", [("C3", 0.01, 99, 6, False), ("C3", 0.01, 99, 6, True), ("C4", 0.004, 6, 9, False)])
def test_fused_file_with_device_inflate_is_byte_identical_to_the_reference_output(tmp, name, scale, seed, level, remove):
    bam = synth.make(name, scale, seed=seed)
    inp, out = os.path.join(tmp, "in.bam"), os.path.join(tmp, "out.bam")
    bamio.write_bam(inp, bam)
    st = bamhost.dedup_file(inp, out, remove_duplicates=remove, level=level, gpu_inflate=True)
    assert st["gpu_inflate"] and st["dedup"]["inflate_blocks"] > 0
    key = "%s_%g_%d_c%d%s" % (name, scale, seed, level, "_r" if remove else "")
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == str(GOLD[key])
    out2 = os.path.join(tmp, "out2.bam")
    st2 = bamhost.dedup_file(inp, out2, remove_duplicates=remove, level=level, gpu_inflate=False)
    assert not st2["gpu_inflate"] and open(out2, "rb").read() == open(out, "rb").read()


def test_fused_binary_inflate_modes_agree(tmp):
    exe = _build.ensure_fused()
    bam, g = load_golden("synth_C3")
    inp = os.path.join(tmp, "in.bam")
    bamio.write_bam(inp, bam)
    outs = []
    for extra in ([], ["--cpu-inflate"], ["--pinned"]):
        out = os.path.join(tmp, "o%d.rawbam" % len(outs))
        r = subprocess.run([exe, inp, "-o", out, "-F", "rawbam", "--nopg", "-v"] + extra, capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        assert (b"gpu inflate" in r.stderr) == ("--cpu-inflate" not in extra)
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1] == outs[2]
    assert np.array_equal(bamio.parse_bam_stream(outs[0]).flags(), g["flags_nosplit_v"])
    # an uncompressed input falls back to the host loader
    raw_in = os.path.join(tmp, "in.rawbam")
    bamio.write_bam(raw_in, bam, raw=True)
    out = os.path.join(tmp, "o_raw.rawbam")
    r = subprocess.run([exe, raw_in, "-o", out, "-F", "rawbam", "--nopg"], capture_output=True, timeout=300)
    assert r.returncode == 0 and open(out, "rb").read() == outs[0]


# ---- record framing on the device (oge_gpu_dedup_frame): speculative parallel chain walk + proof
def device_frame(path):
    h = bamhost.HostBam(path, defer_inflate=True)
    ctx = dedup.DedupContext(n_ref=len(h.refs), max_ref_len=max([l for _, l in h.refs], default=0))
    ix = h.bgzf_index()
    ctx.push_bgzf(ix["comp"], ix["comp_bytes"], ix["in_off"], ix["csize"], ix["isize"], ix["n_blocks"], ix["header_bytes"], None)
    return ctx, h


@pytest.mark.parametrize("name,scale", [("C3", 0.05), ("C1", 0.02), ("C4", 0.004)])
def test_device_framing_equals_the_sequential_chain(tmp, name, scale):
    bam = synth.make(name, scale, seed=17)
    p = os.path.join(tmp, "x.bam")
    open(p, "wb").write(bamhost.bgzf_compress(bamio.serialize_bam_stream(bam), 1))
    ctx, h = device_frame(p)
    with ctx, h:
        assert ctx.frame(len(bam.records)) == bam.n
        assert np.array_equal(ctx.offsets(), bam.offsets)
        assert ctx.stats()["ms_frame"] > 0
        ctx.set_header(h.text)
        ctx.run()
        assert np.array_equal(ctx.flags(), oracle.markdup(bam.records, bam.offsets, bam.text))


def test_device_framing_is_proven_not_guessed(tmp):
    """A record whose tag bytes hold a perfectly plausible chain of fake records, placed so that it is the first thing a
    64 KB chunk sees: the guess takes the bait, the proof (entry == predecessor's exit) replaces it."""
    import fixtures
    text = "@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr1\tLN:100000\n@RG\tID:rg1\tLB:libA\n"
    small = [fixtures._rec("r%05d" % i, 0, 0, 100 + i, "100M", -1, 0, "I") for i in range(1200)]
    fake_chain = b"".join(fixtures._rec("fake%d" % i, 0, 0, 5 + i, "100M", -1, 0, "I") for i in range(4))
    recs, pos = [], 0
    i = 0
    while pos < 65536 - 2500:
        recs.append(small[i]); pos += len(small[i]); i += 1
    # the carrier: core + name + cigar + seq + qual come first; then RG and a B:C array: [0xFF padding][fake chain][0xFF padding]
    probe = bamio.build_record("carrier", 0, 0, 100 + i, 60, "100M", -1, -1, 0, fixtures.SEQ100, 40, bamio.tag_z("RG", "rg1") + b"XBBC" + (0).to_bytes(4, "little"))
    data_at = pos + len(probe)                      # stream offset of the first array byte
    pad_front = (65536 - data_at) + 3               # the fake chain starts 3 bytes into chunk 1
    assert 0 < pad_front < 4000
    arr = b"\xff" * pad_front + fake_chain + b"\xff" * 200
    carrier = bamio.build_record("carrier", 0, 0, 100 + i, 60, "100M", -1, -1, 0, fixtures.SEQ100, 40,
                                 bamio.tag_z("RG", "rg1") + b"XBBC" + len(arr).to_bytes(4, "little") + arr)
    recs.append(carrier)
    assert pos + len(carrier) > 65536 + 3 + len(fake_chain)
    recs += small[i:]
    records, offsets = bamio.concat_records(recs)
    bam = bamio.BamFile(text=text, refs=[("chr1", 100000)], records=records, offsets=offsets)
    assert bytes(records[65539:65539 + len(fake_chain)]) == fake_chain
    p = os.path.join(tmp, "trap.bam")
    open(p, "wb").write(bamhost.bgzf_compress(bamio.serialize_bam_stream(bam), 6))
    ctx, h = device_frame(p)
    with ctx, h:
        assert ctx.frame(len(records)) == bam.n
        assert np.array_equal(ctx.offsets(), bam.offsets)
        assert ctx.stats()["frame_repairs"] >= 1
        ctx.set_header(h.text)
        ctx.run()
        assert np.array_equal(ctx.flags(), oracle.markdup(bam.records, bam.offsets, bam.text))


def test_device_framing_reports_broken_chains_like_the_reference(tmp):
    bam = synth.make("C1", 0.01, seed=3)
    raw = bytearray(bamio.serialize_bam_stream(bam))
    first = len(raw) - len(bam.records)
    k = int(bam.offsets[len(bam.offsets) // 2])
    bad = bytearray(raw)
    bad[first + k: first + k + 4] = (20000).to_bytes(4, "little")
    for data, msg in ((bytes(bad), "Invalid BAM block size(20000)"), (bytes(raw[:-7]), "Expected more bytes reading BAM core")):
        p = os.path.join(tmp, "b.bam")
        open(p, "wb").write(bamhost.bgzf_compress(data, 1))
        ctx, h = device_frame(p)
        with ctx, h:
            with pytest.raises(dedup.DedupError) as e:
                ctx.frame(0)
            assert e.value.code == -7 and msg in str(e.value)


def test_device_framing_empty_file(tmp):
    import fixtures
    b1, _ = fixtures.fixture1()
    empty = bamio.BamFile(text=b1.text, refs=list(b1.refs), records=np.zeros(0, np.uint8), offsets=np.zeros(1, np.uint64))
    inp, out = os.path.join(tmp, "e.bam"), os.path.join(tmp, "o.bam")
    bamio.write_bam(inp, empty)
    st = bamhost.dedup_file(inp, out)
    assert st["n_out"] == 0 and bamio.read_bam(out).n == 0
