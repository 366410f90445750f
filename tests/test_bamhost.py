"""CPU suite: the host-side BAM streaming layer (include/oge_bam_host.h, SURVEY 8(f) f1 + f2) against the reference.

Golden values (tests/golden/bamfile.npz, made by tests/golden/make_bamfile_golden.py) are sha256 hashes of the output
FILES the compiled reference writes and the header texts it renders; where the compiled reference is present it is
also run live.  The duplicate flags in these tests come from the oracle (test infrastructure); the product path takes
them from the GPU (tests/test_gpu_fused.py)."""
import ctypes as C
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import fixtures
import oracle
from conftest import GOLDEN, ROOT, load_golden
from openge_b200 import _build, bamhost, bamio, synth

GOLD = dict(np.load(os.path.join(GOLDEN, "bamfile.npz")))


@pytest.fixture()
def tmp():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        yield d


def host_dedup_with_oracle_flags(bam, d, level, remove=False, fmt=None, threads=0, pg=None):
    inp, out = os.path.join(d, "in.bam"), os.path.join(d, "host_out.bam")
    bamio.write_bam(inp, bam)
    with bamhost.HostBam(inp, threads=threads) as h:
        flags = oracle.markdup(h.records.copy(), h.offsets.copy(), h.text)
        h.apply_flags(flags, remove, threads)
        h.store(out, fmt, level, pg, threads=threads)
    return open(out, "rb").read()


def test_header_symbols_exported():
    text = open(os.path.join(ROOT, "include", "oge_bam_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(oge_b(?:am|gzf)_\w+)\s*\(", text)))
    L = bamhost.lib()
    for n in names:
        assert hasattr(L, n), "liboge_bamhost.so does not export %s" % n
    assert sorted(bamhost.EXPORTS) == names


def test_load_frames_like_the_python_reader(tmp):
    bam = synth.make("C3", 0.01, seed=3)
    for raw in (False, True):
        p = os.path.join(tmp, "x.bam")
        bamio.write_bam(p, bam, raw=raw)
        with bamhost.HostBam(p, threads=3) as h:
            assert h.n == bam.n and h.text == bam.text and h.refs == bam.refs
            assert np.array_equal(h.records, bam.records) and np.array_equal(h.offsets, bam.offsets)
            rg, libs, unknown, n_libs = h.library_table()
            from openge_b200 import header
            ids2, libs2, unknown2, n2 = header.library_table(bam.text)
            assert rg == list(ids2) and unknown == unknown2 and n_libs == n2
            # library ids only have to induce the same partition of the read groups
            assert [libs.index(x) for x in libs] == [list(libs2).index(x) for x in libs2]


def test_two_stage_open_equals_load(tmp):
    """oge_bam_open_bgzf / records_buffer / frame_records (the GPU-inflate entry): here the buffer is filled by the host codec."""
    b1, _ = fixtures.fixture1()
    big_text = "@HD\tVN:1.4\tSO:coordinate\n" + "".join("@SQ\tSN:contig%05d\tLN:%d\n" % (i, 1000 + i) for i in range(9000)) + "@RG\tID:rg1\tLB:libA\n@RG\tID:rg2\tLB:libB\n"
    wide = bamio.BamFile(text=big_text, refs=[("contig%05d" % i, 1000 + i) for i in range(9000)], records=b1.records, offsets=b1.offsets)
    for bam in (synth.make("C3", 0.005, seed=4), wide):      # the second header spans several BGZF blocks
        p = os.path.join(tmp, "x.bam")
        bamio.write_bam(p, bam)
        with bamhost.HostBam(p, defer_inflate=True) as h:
            assert h.text == bam.text and h.refs == bam.refs and h.n == 0
            ix = h.bgzf_index()
            raw = bamhost.bgzf_decompress(open(p, "rb").read())
            assert ix["header_bytes"] == len(raw) - len(bam.records) and ix["comp_bytes"] == os.path.getsize(p)
            isize = np.ctypeslib.as_array((C.c_uint32 * ix["n_blocks"]).from_address(ix["isize"]))
            assert int(isize.sum()) == len(raw)
            C.memmove(h.records_buffer(), raw[ix["header_bytes"]:], len(raw) - ix["header_bytes"])
            h.frame_records()
            assert h.n == bam.n and np.array_equal(h.records, bam.records) and np.array_equal(h.offsets, bam.offsets)
            h.apply_flags(bam.flags())
            h.store(os.path.join(tmp, "y.rawbam"), "rawbam")
        assert bamio.read_bam(os.path.join(tmp, "y.rawbam")).n == bam.n
    p = os.path.join(tmp, "raw.bam")
    bamio.write_bam(p, bam, raw=True)
    with pytest.raises(bamhost.BamHostError) as e:
        bamhost.HostBam(p, defer_inflate=True)
    assert "uncompressed BAM stream" in str(e.value)


def test_bgzf_codec_round_trip_and_block_layout():
    rng = np.random.default_rng(1)
    for n in (0, 1, 65535, 65536, 65537, 3 * 65536, 1_000_003):
        raw = (rng.integers(0, 4, size=n, dtype=np.uint8) + 65).tobytes()
        for level in (0, 1, 6):
            z = bamhost.bgzf_compress(raw, level, threads=4)
            assert bamhost.bgzf_decompress(z, threads=4) == raw
            assert bamio.bgzf_decompress(z) == raw
            # block sequence of BgzfOutputStream::write + close: full blocks, the current (maybe empty) block, an empty block
            full = 65536 - 64 if level == 0 else 65536
            sizes, pos = [], 0
            while pos < len(z):
                bs = int.from_bytes(z[pos + 16: pos + 18], "little") + 1
                sizes.append(int.from_bytes(z[pos + bs - 4: pos + bs], "little"))
                pos += bs
            assert sizes == [full] * (n // full) + [n % full, 0]
    assert bamhost.bgzf_decompress(bamio.bgzf_compress(b"hello" * 100000)) == b"hello" * 100000


@pytest.mark.parametrize("name,scale,seed,level,remove", [("C3", 0.01, 99, 6, False), ("C3", 0.01, 99, 1, False), ("C3", 0.01, 99, 0, False),
                                                          ("C3", 0.01, 99, 6, True), ("C1", 0.02, 5, 6, False), ("C4", 0.004, 6, 9, False)])
def test_output_file_is_byte_identical_to_the_reference_golden(tmp, name, scale, seed, level, remove):
    bam = synth.make(name, scale, seed=seed)
    data = host_dedup_with_oracle_flags(bam, tmp, level, remove, threads=5)
    key = "%s_%g_%d_c%d%s" % (name, scale, seed, level, "_r" if remove else "")
    assert hashlib.sha256(data).hexdigest() == str(GOLD[key])


@pytest.mark.parametrize("case,level", [("yhet208", 6), ("yhet208", 1), ("edge_cases", 6), ("edge_cases", 1)])
def test_real_data_output_file_is_byte_identical_to_the_reference_golden(tmp, case, level):
    """test/data/208.yhet.bam of the reference (real reads, a header that is not in canonical form) and the edge-case records."""
    bam, _ = load_golden(case)
    data = host_dedup_with_oracle_flags(bam, tmp, level, threads=4)
    assert hashlib.sha256(data).hexdigest() == str(GOLD["%s_c%d" % (case, level)])


@pytest.mark.parametrize("case", ["reordered_fields", "regrouped_lines", "no_hd", "unterminated_last_line", "repeated_rg_and_no_lb", "crlf_and_numbers"])
def test_header_is_rendered_like_the_reference(tmp, case):
    bam = fixtures.header_cases()[case]
    assert bamhost.header_render(bam.text) == str(GOLD["header_" + case])
    data = host_dedup_with_oracle_flags(bam, tmp, 6)
    assert hashlib.sha256(data).hexdigest() == str(GOLD["headerfile_" + case])


def test_canonical_headers_pass_through_unchanged():
    for case in ("a3_fixture1", "edge_cases", "yhet208", "synth_C3"):
        bam, _ = load_golden(case)
        if case == "yhet208":
            continue      # its header is not in the reference's canonical form
        assert bamhost.header_render(bam.text) == bam.text


def test_pg_line_like_file_writer(tmp):
    bam, _ = fixtures.fixture1()
    bam.text += "@PG\tID:openge\tVN:0.1\n"
    data = host_dedup_with_oracle_flags(bam, tmp, 6, fmt="rawbam", pg="openge dedup in.bam -o out.bam ")
    out = bamio.parse_bam_stream(data)
    # file_writer.cpp:76-89: ID openge, or openge-2.. when taken; CL then VN (BamProgramRecord::toString)
    assert out.text.endswith("@PG\tID:openge\tVN:0.1\n@PG\tID:openge-2\tCL:openge dedup in.bam -o out.bam \tVN:0.3-b200\n")


def test_bins_are_recomputed_like_the_writer(tmp):
    # edge_cases holds zero-length / unmapped / clipped records; the golden flags come from the reference, and the
    # reference's rawbam output carries the bins its writer computed (util/bam_serializer.h:88-116)
    for case in ("edge_cases", "a3_fixture2", "synth_C3"):
        bam, g = load_golden(case)
        want = oracle.ref_dedup(bam) if oracle.ref_available() else None
        data = host_dedup_with_oracle_flags(bam, tmp, 6, fmt="rawbam")
        out = bamio.parse_bam_stream(data)
        assert np.array_equal(out.flags(), g["flags_nosplit_v"])
        if want is not None:
            assert np.array_equal(out.records, want.records) and out.text == want.text


@pytest.mark.skipif(not oracle.ref_available(), reason="compiled reference not present")
def test_output_file_equals_live_reference_file(tmp):
    bam = synth.make("C3", 0.02, seed=21)
    inp, ref_out = os.path.join(tmp, "in.bam"), os.path.join(tmp, "ref.bam")
    bamio.write_bam(inp, bam)
    for level in (6, 2):
        for _ in range(4):
            try:
                r = subprocess.run([_build.REF_BIN, "-T", tmp, "--nosplit", "-v", "-c", str(level), inp, ref_out], capture_output=True, timeout=120)
                break
            except subprocess.TimeoutExpired:
                r = None
        if r is None:
            pytest.skip("reference did not terminate")
        assert r.returncode == 0
        assert host_dedup_with_oracle_flags(bam, tmp, level, threads=3) == open(ref_out, "rb").read()
    # and the reference reads what this layer wrote
    with open(os.path.join(tmp, "mine.bam"), "wb") as f:
        f.write(host_dedup_with_oracle_flags(bam, tmp, 6))
    again = oracle.ref_dedup(bamio.read_bam(os.path.join(tmp, "mine.bam")))
    assert again.n == bam.n


def test_errors_are_reported(tmp):
    bam, _ = fixtures.fixture1()
    p = os.path.join(tmp, "x.bam")
    with pytest.raises(bamhost.BamHostError) as e:
        bamhost.HostBam(os.path.join(tmp, "missing.bam"))
    assert e.value.code == -1
    open(p, "wb").write(b"not a bam file at all, not even close....")
    with pytest.raises(bamhost.BamHostError) as e:
        bamhost.HostBam(p)
    assert e.value.code == -2
    good = bamio.bgzf_compress(bamio.serialize_bam_stream(bam))
    open(p, "wb").write(good[:-40])      # truncated inside the last blocks
    with pytest.raises(bamhost.BamHostError):
        bamhost.HostBam(p)
    raw = bytearray(bamio.serialize_bam_stream(bam))
    first = len(raw) - len(bam.records)
    raw[first: first + 4] = (20000).to_bytes(4, "little")      # block_size > 10000 (util/bam_deserializer.h:160)
    open(p, "wb").write(bytes(raw))
    with pytest.raises(bamhost.BamHostError) as e:
        bamhost.HostBam(p)
    assert "Invalid BAM block size(20000)" in str(e.value)
    # a gzip member whose extra field is not the 6-byte BC subfield: rejected like the reference does
    # (util/bgzf_input_stream.cpp:84-98: "BGZF GZ extra field is incorrect"), e.g. plain gzip output
    import gzip
    open(p, "wb").write(gzip.compress(bytes(raw[:1000])))
    with pytest.raises(bamhost.BamHostError) as e:
        bamhost.HostBam(p)
    assert e.value.code == -2 and ("unexpected flags" in str(e.value) or "extra field" in str(e.value))
    xl = bytearray(good)
    xl[10:12] = (8).to_bytes(2, "little")      # XLEN = 8
    open(p, "wb").write(bytes(xl))
    with pytest.raises(bamhost.BamHostError) as e:
        bamhost.HostBam(p)
    assert "extra field is incorrect" in str(e.value)
    bad = bam.text.replace("@RG\tID:rg1", "@XX\tID:rg1")
    with pytest.raises(bamhost.BamHostError) as e:
        bamhost.header_render(bad)
    assert "wasn't CO RG SQ PG or HD" in str(e.value)
    # a field shorter than "XX:" makes the reference die in substr (std::out_of_range, util/bam_header.cpp:44-46); here: an error
    with pytest.raises(bamhost.BamHostError) as e:
        bamhost.header_render(bam.text + "@PG\tID:p2\tCL:a b\tc\n")
    assert "too short" in str(e.value)
    with bamhost.HostBam(_write(tmp, bam)) as h:
        with pytest.raises(ValueError):
            h.apply_flags(np.zeros(3, np.uint16))
        with pytest.raises(bamhost.BamHostError):
            h.store(os.path.join(tmp, "o.sam"), "sam")


def _write(d, bam):
    p = os.path.join(d, "ok.bam")
    bamio.write_bam(p, bam)
    return p


def test_empty_file_round_trip(tmp):
    b1, _ = fixtures.fixture1()
    empty = bamio.BamFile(text=b1.text, refs=list(b1.refs), records=np.zeros(0, np.uint8), offsets=np.zeros(1, np.uint64))
    data = host_dedup_with_oracle_flags(empty, tmp, 6)
    out = bamio.parse_bam_stream(bamio.bgzf_decompress(data))
    assert out.n == 0 and out.text == b1.text and out.refs == b1.refs


def test_host_layer_has_no_cuda_and_no_oracle():
    src = open(os.path.join(ROOT, "openge_b200", "host", "bam_host.cpp")).read()
    assert "cuda" not in src.lower().replace("no cuda", "") and "oracle" not in src
    out = subprocess.run(["ldd", _build.ensure_bamhost()], capture_output=True, text=True).stdout
    assert "libcudart" not in out and "openge_b200" not in out


def test_store_members_writes_the_same_stream_as_store(tmp):
    """oge_bam_store_members: the output file around record members that were compressed elsewhere (on the device,
    oge_gpu_dedup_deflate; here: by python's zlib in blocks of another size).  After decompression the file must equal
    oge_bam_store's -- header re-rendered the same way, @PG line, the records, the empty end-of-file member last."""
    import zlib
    bam = synth.make("C3", 0.01, seed=12)
    inp = os.path.join(tmp, "in.bam")
    bamio.write_bam(inp, bam)
    for pg in (None, "openge dedup in.bam -o out.bam"):
        a, b = os.path.join(tmp, "a.bam"), os.path.join(tmp, "b.bam")
        with bamhost.HostBam(inp) as h:
            flags = oracle.markdup(h.records.copy(), h.offsets.copy(), h.text)
            h.apply_flags(flags, False, 0)
            h.store(a, None, 6, pg)
            rec = h.records.tobytes()      # flag-patched, bins recomputed
            members = []
            for k in range(0, len(rec), 50000):
                chunk = rec[k:k + 50000]
                co = zlib.compressobj(1, zlib.DEFLATED, -15)
                z = co.compress(chunk) + co.flush()
                members.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + (len(z) + 25).to_bytes(2, "little") + z +
                               zlib.crc32(chunk).to_bytes(4, "little") + len(chunk).to_bytes(4, "little"))
            h.store_members(b, np.frombuffer(b"".join(members), dtype=np.uint8), 6, pg)
        za, zb = open(a, "rb").read(), open(b, "rb").read()
        assert bamhost.bgzf_decompress(za) == bamhost.bgzf_decompress(zb)
        assert zb.endswith(za[-28:]) and za[-28:].endswith(b"\x00\x00\x00\x00\x00\x00\x00\x00")      # the empty member ends both files
        got = bamio.read_bam(b)
        assert got.n == bam.n and np.array_equal(got.flags(), flags) and (pg is None) == ("@PG" not in got.text.replace(bam.text, ""))
    # no records at all: header members and the end-of-file member
    with bamhost.HostBam(inp) as h:
        e = os.path.join(tmp, "e.bam")
        h.store_members(e, np.zeros(0, dtype=np.uint8), 6, None)
    assert bamio.read_bam(e).n == 0


def test_store_members_large_goes_through_the_mapping(tmp):
    # 80 MB of members: above the size from which the file is sized first and filled through a shared mapping by several threads
    import zlib
    bam = synth.make("C1", 0.002, seed=1)
    inp, out = os.path.join(tmp, "in.bam"), os.path.join(tmp, "big.bam")
    bamio.write_bam(inp, bam)
    rng = np.random.default_rng(5)
    payload = rng.integers(0, 256, 65000, dtype=np.uint8).tobytes()
    co = zlib.compressobj(0, zlib.DEFLATED, -15)
    z = co.compress(payload) + co.flush()
    member = (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + (len(z) + 25).to_bytes(2, "little") + z +
              zlib.crc32(payload).to_bytes(4, "little") + len(payload).to_bytes(4, "little"))
    n = (80 << 20) // len(member) + 1
    members = np.frombuffer(member * n, dtype=np.uint8)
    with bamhost.HostBam(inp) as h:
        h.store_members(out, members, 6, None)
        small = os.path.join(tmp, "small.bam")
        h.store_members(small, members[:len(member) * 3], 6, None)
    zb = open(out, "rb").read()
    zs = open(small, "rb").read()
    head_len = len(zs) - 3 * len(member) - 28
    assert zb[:head_len] == zs[:head_len] and zb[-28:] == zs[-28:]      # same header members, same end-of-file member
    assert zb[head_len:-28] == members.tobytes()
