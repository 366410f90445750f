"""Host-side BAM framing: BGZF/rawbam <-> (header, concatenated record bytes, offsets).

This is plumbing for tests, the CLI and bench.py; BGZF stays on the host
(BASELINE.json north_star).  The in-memory form it produces is exactly the device
record layout of DESIGN.md: raw BAM records, each starting with its 4-byte
``block_size``, back to back, plus an ``offsets`` array of ``n+1`` u64 byte offsets.

Reference behaviour mirrored (file:line under /root/reference/openge/src/util):
  * record framing             bam_deserializer.h:144-193
  * header + reference list    bam_serializer.h:46-81
  * "rawbam" = the same stream with no BGZF wrapper (magic ``BA``)
                               read_stream_reader.h:80-81, read_stream_reader.cpp:30
  * BGZF block geometry        bgzf_output_stream.h:26, bgzf_input_stream.cpp:65-142
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
BGZF_MAX_PAYLOAD = 65280


@dataclass
class BamFile:
    """A decoded BAM: header text, reference dictionary and the framed record buffer."""

    text: str
    refs: list = field(default_factory=list)          # [(name, length)]
    records: np.ndarray = None                          # uint8, concatenated raw records
    offsets: np.ndarray = None                          # uint64, n+1 entries

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    def flags(self) -> np.ndarray:
        """u16 flag word of every record (byte 18 of the record, see DESIGN.md layout)."""
        off = self.offsets[:-1].astype(np.int64) + 18
        return self.records[off].astype(np.uint16) | (self.records[off + 1].astype(np.uint16) << 8)


def bgzf_decompress(data: bytes) -> bytes:
    """Inflate every BGZF member of ``data`` (any XLEN, unlike bgzf_input_stream.cpp:95)."""
    out = []
    pos = 0
    n = len(data)
    while pos < n:
        if data[pos:pos + 2] != b"\x1f\x8b":
            raise ValueError("not a BGZF block at offset %d" % pos)
        xlen = struct.unpack_from("<H", data, pos + 10)[0]
        xpos = pos + 12
        bsize = None
        while xpos < pos + 12 + xlen:
            si1, si2, slen = data[xpos], data[xpos + 1], struct.unpack_from("<H", data, xpos + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", data, xpos + 4)[0] + 1
            xpos += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC field")
        payload = data[pos + 12 + xlen: pos + bsize - 8]
        isize = struct.unpack_from("<I", data, pos + bsize - 4)[0]
        if isize:
            out.append(zlib.decompress(payload, -15, isize))
        pos += bsize
    return b"".join(out)


def bgzf_compress(raw: bytes, level: int = 1) -> bytes:
    out = []
    for i in range(0, len(raw), BGZF_MAX_PAYLOAD):
        chunk = raw[i:i + BGZF_MAX_PAYLOAD]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = len(comp) + 25
        out.append(struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize))
        out.append(comp)
        out.append(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    out.append(BGZF_EOF)
    return b"".join(out)


def frame_records(buf, start: int = 0):
    """Walk the ``block_size`` chain from ``start``; returns offsets (u64, n+1) relative to ``start``."""
    mv = np.frombuffer(buf, dtype=np.uint8)
    n = len(mv)
    offs = [0]
    pos = start
    while pos < n:
        if pos + 4 > n:
            raise ValueError("truncated BAM record header")
        bs = int(mv[pos]) | int(mv[pos + 1]) << 8 | int(mv[pos + 2]) << 16 | int(mv[pos + 3]) << 24
        if bs < 32:
            raise ValueError("invalid BAM block size %d" % bs)
        pos += 4 + bs
        if pos > n:
            raise ValueError("truncated BAM record")
        offs.append(pos - start)
    return np.asarray(offs, dtype=np.uint64)


def parse_bam_stream(raw: bytes) -> BamFile:
    if raw[:4] != b"BAM\x01":
        raise ValueError("bad BAM magic")
    l_text = struct.unpack_from("<i", raw, 4)[0]
    text = raw[8:8 + l_text].decode("latin-1")
    pos = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, pos)[0]
    pos += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, pos)[0]
        name = raw[pos + 4: pos + 4 + l_name - 1].decode("latin-1")
        l_ref = struct.unpack_from("<i", raw, pos + 4 + l_name)[0]
        refs.append((name, l_ref))
        pos += 8 + l_name
    try:
        from . import synth
        offsets = synth.frame_records_fast(raw, pos)
    except Exception:
        offsets = frame_records(raw, pos)
    records = np.frombuffer(raw, dtype=np.uint8, count=int(offsets[-1]), offset=pos).copy()
    return BamFile(text=text.rstrip("\x00"), refs=refs, records=records, offsets=offsets)


def read_bam(path: str) -> BamFile:
    """Read a BGZF BAM or an OpenGE rawbam file."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] == b"\x1f\x8b":
        data = bgzf_decompress(data)
    return parse_bam_stream(data)


def serialize_bam_stream(bam: BamFile, records=None) -> bytes:
    text = bam.text.encode("latin-1")
    parts = [b"BAM\x01", struct.pack("<i", len(text)), text, struct.pack("<i", len(bam.refs))]
    for name, length in bam.refs:
        nb = name.encode("latin-1") + b"\x00"
        parts.append(struct.pack("<i", len(nb)) + nb + struct.pack("<i", length))
    rec = bam.records if records is None else records
    parts.append(rec.tobytes() if isinstance(rec, np.ndarray) else bytes(rec))
    return b"".join(parts)


def write_bam(path: str, bam: BamFile, records=None, level: int = 1, raw: bool = False) -> None:
    """Write ``bam`` as BGZF BAM, or as OpenGE rawbam when ``raw`` (uncompressed stream)."""
    stream = serialize_bam_stream(bam, records)
    with open(path, "wb") as f:
        f.write(stream if raw else bgzf_compress(stream, level))


# --------------------------------------------------------------------------------------
# Small record builder used by fixtures/tests (SAM-like fields -> raw BAM record bytes).

_CIGAR_OPS = "MIDNSHP=X"
_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def reg2bin(beg: int, end: int) -> int:
    """SAM-spec bin for the 0-based half-open interval [beg, end) (bam_serializer.h:92-103)."""
    end -= 1
    if beg >> 14 == end >> 14:
        return 4681 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return 585 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return 73 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return 9 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return 1 + (beg >> 26)
    return 0


def parse_cigar(cigar: str):
    ops = []
    if cigar in ("*", ""):
        return ops
    num = ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            ops.append((int(num), _CIGAR_OPS.index(ch)))
            num = ""
    return ops


def build_record(name: str, flag: int, ref_id: int, pos0: int, mapq: int, cigar: str,
                 mate_ref: int, mate_pos0: int, tlen: int, seq: str, qual, tags: bytes = b"") -> bytes:
    """One raw BAM record (with its leading block_size).  ``pos0``/``mate_pos0`` are 0-based;
    ``qual`` is raw phred bytes (or an int applied to every base)."""
    ops = parse_cigar(cigar)
    l_seq = len(seq)
    if isinstance(qual, int):
        qual = bytes([qual]) * l_seq
    assert len(qual) == l_seq
    nb = name.encode("latin-1") + b"\x00"
    ref_len = sum(l for l, o in ops if o in (0, 2, 3, 7, 8))
    end = pos0 + ref_len if ref_len else pos0 + 1
    b = reg2bin(max(pos0, 0), max(end, 1)) if pos0 >= 0 else 4680
    cig = b"".join(struct.pack("<I", (l << 4) | o) for l, o in ops)
    packed = bytearray((l_seq + 1) // 2)
    for i, ch in enumerate(seq):
        code = _SEQ_CODE.get(ch.upper(), 15)
        packed[i >> 1] |= code << (4 if (i & 1) == 0 else 0)
    body = struct.pack("<iiBBHHHiiii", ref_id, pos0, len(nb), mapq, b, len(ops), flag, l_seq,
                       mate_ref, mate_pos0, tlen) + nb + cig + bytes(packed) + bytes(qual) + tags
    return struct.pack("<i", len(body)) + body


def tag_z(tag: str, value: str) -> bytes:
    return tag.encode() + b"Z" + value.encode("latin-1") + b"\x00"


def concat_records(recs) -> tuple:
    """list of raw records -> (records uint8 array, offsets u64 n+1)."""
    offs = np.zeros(len(recs) + 1, dtype=np.uint64)
    if recs:
        offs[1:] = np.cumsum([len(r) for r in recs], dtype=np.uint64)
    return np.frombuffer(b"".join(recs), dtype=np.uint8).copy(), offs
