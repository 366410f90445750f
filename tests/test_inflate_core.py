"""CPU suite: the product's DEFLATE decoder (openge_b200/csrc/inflate_core.cuh, the body of the GPU BGZF inflate kernel)
compiled for the host with one lane (tests/native/inflate_host.cpp) against zlib.  The warp-parallel execution of the same
source is covered by tests/test_gpu_inflate.py."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    out_dir = os.path.join(ROOT, "tests", "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboge_inflate_host.so")
    src = os.path.join(ROOT, "tests", "native", "inflate_host.cpp")
    core = os.path.join(ROOT, "openge_b200", "csrc", "inflate_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I", os.path.join(ROOT, "openge_b200", "csrc"), src, "-o", so], check=True)
    L = C.CDLL(so)
    L.oge_test_inflate_block2.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_uint, C.c_int]
    return L


SMALL = [0]      # 1: the table widths of the thread-per-block kernel (9 / 7 bits); 2: its state-machine form


def inflate(L, z, n):
    out = np.zeros(max(1, n), dtype=np.uint8)
    rc = L.oge_test_inflate_block2(z, len(z), out.ctypes.data, n, SMALL[0])
    return rc, out[:n].tobytes()


def payloads():
    rng = np.random.default_rng(0)
    yield b""
    yield b"a"
    yield b"hello hello hello hello"
    yield bytes(65536)                                                      # one long run: overlapping matches, distance 1
    yield bytes(rng.integers(0, 256, 65536, dtype=np.uint8))                # incompressible: stored blocks at every level
    yield bytes(rng.integers(65, 69, 65536, dtype=np.uint8))                # 2 bits of entropy per byte
    yield (b"ACGT" * 100 + bytes(rng.integers(0, 256, 50, dtype=np.uint8))) * 100
    for n in (1, 2, 3, 100, 1000, 30000, 65536):
        yield bytes(rng.integers(0, 40, n, dtype=np.uint8) + 33)            # quality-like
    p = np.array([2.0 ** -(i / 6) for i in range(256)])
    yield bytes(rng.choice(256, size=65536, p=p / p.sum()).astype(np.uint8))  # skewed: codes longer than the 10-bit primary table
    from openge_b200 import bamio, synth
    raw = bamio.serialize_bam_stream(synth.make("C3", 0.002, seed=5))
    for k in range(0, min(len(raw), 4 * 65536), 65536):
        yield raw[k:k + 65536]                                              # real BAM bytes


def test_tables_fit_the_shared_memory_budget(lib):
    assert lib.oge_test_inflate_tables_bytes() * 8 <= 48 * 1024      # 8 warps per CTA, static shared memory


@pytest.mark.parametrize("small", [0, 1, 2])
def test_decoder_matches_zlib_on_every_block_type(lib, small):
    SMALL[0] = small
    n = 0
    for data in payloads():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
                z = co.compress(data) + co.flush()
                rc, out = inflate(lib, z, len(data))
                assert rc == 0 and out == data, (len(data), level, strategy, rc)
                n += 1
    assert n > 300
    SMALL[0] = 0


@pytest.mark.parametrize("small", [0, 2])
def test_multi_block_streams_with_sync_flushes(lib, small):
    SMALL[0] = small
    _multi(lib)
    SMALL[0] = 0


def _multi(lib):
    rng = np.random.default_rng(3)
    data = bytes(rng.integers(0, 8, 40000, dtype=np.uint8))
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    z = b""
    for k in range(0, len(data), 7000):      # Z_SYNC_FLUSH puts an empty stored block between Huffman blocks
        z += co.compress(data[k:k + 7000]) + co.flush(zlib.Z_SYNC_FLUSH)
    z += co.flush()
    rc, out = inflate(lib, z, len(data))
    assert rc == 0 and out == data


@pytest.mark.parametrize("small", [0, 2])
def test_corrupt_streams_are_rejected_not_overrun(lib, small):
    SMALL[0] = small
    _corrupt(lib)
    SMALL[0] = 0


def _corrupt(lib):
    rng = np.random.default_rng(4)
    data = bytes(rng.integers(0, 40, 20000, dtype=np.uint8))
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    z = co.compress(data) + co.flush()
    assert inflate(lib, z, len(data) - 1)[0] != 0          # ISIZE too small: output overrun detected
    assert inflate(lib, z, len(data) + 1)[0] != 0          # ISIZE too large: short stream
    assert inflate(lib, z[: len(z) // 2], len(data))[0] != 0
    assert inflate(lib, b"\x07" + z[1:], len(data))[0] != 0      # BTYPE 3
    bad = 0
    for k in range(200):      # random corruption never crashes; it is either detected or decodes to other bytes
        zz = bytearray(z)
        zz[int(rng.integers(0, len(zz)))] ^= 1 << int(rng.integers(0, 8))
        rc, out = inflate(lib, bytes(zz), len(data))
        bad += rc != 0 or out != data
    assert bad >= 190


def test_decoder_fuzz_against_zlib(lib):
    """Randomised structure: runs, repeats at every distance class, literals of varying entropy, all levels/strategies,
    window sizes and mem levels, every decoder form.  A few thousand streams; any mismatch prints its seed."""
    rng = np.random.default_rng(2024)
    forms = (0, 1, 2)
    n_cases = 0
    for case in range(400):
        parts = []
        size = int(rng.integers(0, 65536))
        while sum(len(p) for p in parts) < size:
            kind = int(rng.integers(0, 5))
            if kind == 0:      # literals over an alphabet of random width
                parts.append(rng.integers(0, int(rng.integers(1, 257)), int(rng.integers(1, 4000)), dtype=np.uint8).tobytes())
            elif kind == 1:    # a run
                parts.append(bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 3000)))
            elif kind == 2 and parts:      # copy from earlier output at a random distance (up to the 32 KB window and beyond)
                prev = b"".join(parts)
                d = int(rng.integers(1, min(len(prev), 40000) + 1))
                l = int(rng.integers(3, 600))
                parts.append((prev[-d:] * (l // d + 1))[:l])
            elif kind == 3:    # geometric symbol distribution: long codes
                p = np.array([2.0 ** -(i / float(rng.integers(2, 12))) for i in range(256)])
                parts.append(rng.choice(256, size=int(rng.integers(1, 5000)), p=p / p.sum()).astype(np.uint8).tobytes())
            else:              # text-like
                words = [b"chr1", b"ACGT", b"read", b"\x00\x01\x02", b"IIIIIIII", b"RG:Z:rg1"]
                parts.append(b"".join(words[int(x)] for x in rng.integers(0, len(words), int(rng.integers(1, 300)))))
        data = b"".join(parts)[:65536]
        level = int(rng.integers(0, 10))
        strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED][int(rng.integers(0, 5))]
        wbits = -int(rng.integers(9, 16))
        mem = int(rng.integers(1, 10))
        co = zlib.compressobj(level, zlib.DEFLATED, wbits, mem, strategy)
        z = co.compress(data) + co.flush()
        for form in forms:
            SMALL[0] = form
            rc, out = inflate(lib, z, len(data))
            assert rc == 0 and out == data, ("case", case, "form", form, len(data), level, strategy, wbits, mem, rc)
            n_cases += 1
    SMALL[0] = 0
    assert n_cases == 1200


class _Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, nbits):            # LSB first (header fields, extra bits)
        self.v |= value << self.n
        self.n += nbits

    def code(self, code, nbits):            # a Huffman code: most significant bit first
        for k in range(nbits - 1, -1, -1):
            self.put((code >> k) & 1, 1)

    def bytes(self):
        return self.v.to_bytes((self.n + 7) // 8, "little")


def _canonical(lens):
    """{symbol: length} -> {symbol: (code, length)} (RFC 1951 3.2.2)."""
    out, code = {}, 0
    for l in range(1, 16):
        for s in sorted(k for k, v in lens.items() if v == l):
            out[s] = (code, l)
            code += 1
        code <<= 1
    return out


def _dynamic_block(lit, dist, symbols, cl=None):
    """One final dynamic-Huffman block whose code lengths are given verbatim (no repeat codes), the code-length code being
    `cl` (default: a complete code over the length values 0, 1, 2)."""
    cl = cl or {0: 1, 1: 2, 2: 2}
    cl_codes = _canonical(cl)
    n_lit, n_dist = 257, max(dist, default=0) + 1
    b = _Bits()
    b.put(1, 1); b.put(2, 2); b.put(n_lit - 257, 5); b.put(n_dist - 1, 5); b.put(19 - 4, 4)
    for s in (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15):
        b.put(cl.get(s, 0), 3)
    for s in range(n_lit):
        b.code(*cl_codes[lit.get(s, 0)])
    for s in range(n_dist):
        b.code(*cl_codes[dist.get(s, 0)])
    lit_codes = _canonical(lit)
    for s in symbols:
        b.code(*lit_codes[s])
    b.put(0, 16)      # padding the decoders may look at
    return b.bytes()


@pytest.mark.parametrize("small", [0, 1, 2])
def test_incomplete_code_sets_get_zlibs_verdict(lib, small):
    # zlib (and with it the reference, util/bgzf_input_stream.cpp:110-128) rejects over-subscribed AND incomplete sets of code
    # lengths, except a literal/length or distance code made of one code of length 1
    SMALL[0] = small
    cases = [
        ("complete literal code, single distance code of length 1", _dynamic_block({65: 1, 256: 1}, {0: 1}, [65, 65, 256]), b"AA"),
        ("incomplete literal code (two codes of length 2)", _dynamic_block({65: 2, 256: 2}, {0: 1}, [65, 256]), None),
        ("one literal code of length 1: the end-of-block symbol alone", _dynamic_block({256: 1}, {}, [256]), b""),
        ("incomplete distance code (two codes of length 2)", _dynamic_block({65: 1, 256: 1}, {0: 2, 1: 2}, [65, 256]), None),
        ("incomplete code-length code", _dynamic_block({65: 1, 256: 1}, {0: 1}, [65, 256], cl={0: 2, 1: 2, 2: 2}), None),
        ("over-subscribed literal code", _dynamic_block({65: 1, 66: 1, 256: 1}, {0: 1}, [65, 256]), None),
    ]
    for what, z, want in cases:
        d = zlib.decompressobj(-15)
        try:
            theirs = d.decompress(z)
            ok = d.eof
        except zlib.error:
            theirs, ok = None, False
        assert (theirs if ok else None) == want, what      # the test's own expectation equals zlib's verdict
        rc, out = inflate(lib, z, len(want) if want is not None else 2)
        assert (rc == 0) == (want is not None), (what, rc)
        if want is not None:
            assert out == want, what
    SMALL[0] = 0
