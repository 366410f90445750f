"""CPU suite: the product's DEFLATE compressor (openge_b200/csrc/deflate_core.cuh, the body of the device kernel that
writes the output file's BGZF blocks) compiled for the host with one lane (tests/native/deflate_host.cpp): zlib must
inflate every stream it writes back to the input (zlib rejects incomplete or over-subscribed codes, bad distances and
stored-block length mismatches), the CRC must equal zlib's.  The warp-parallel execution of the same source is covered
by tests/test_gpu_deflate.py."""
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
    so = os.path.join(out_dir, "liboge_deflate_host.so")
    src = os.path.join(ROOT, "tests", "native", "deflate_host.cpp")
    core = os.path.join(ROOT, "openge_b200", "csrc", "deflate_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I", os.path.join(ROOT, "openge_b200", "csrc"), src, "-o", so], check=True)
    L = C.CDLL(so)
    L.oge_test_deflate_block.argtypes = [C.c_char_p, C.c_uint, C.c_void_p]
    L.oge_test_crc32.argtypes = [C.c_char_p, C.c_uint, C.c_int]
    L.oge_test_crc32.restype = C.c_uint
    L.oge_test_build_code.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.oge_test_build_code.restype = C.c_longlong
    L.oge_test_emulate32.argtypes = [C.c_char_p, C.c_uint]
    L.oge_test_emulate32.restype = C.c_longlong
    return L


def deflate(L, data: bytes) -> bytes:
    out = np.zeros(len(data) + 64, dtype=np.uint8)
    n = L.oge_test_deflate_block(data, len(data), out.ctypes.data)
    assert n > 0
    return out[:n].tobytes()


def payloads():
    rng = np.random.default_rng(0)
    yield b"a"
    yield b"ab"
    yield b"abc"
    yield b"abcd"
    yield b"aaaa"
    yield b"aaaaa"
    yield b"hello hello hello hello"
    yield bytes(65535)                                                      # one long run: distance 1, 258-byte matches
    yield bytes(65280)
    yield bytes(rng.integers(0, 256, 65280, dtype=np.uint8))                # incompressible: stored
    yield bytes(rng.integers(0, 256, 65535, dtype=np.uint8))
    yield bytes(rng.integers(65, 69, 65280, dtype=np.uint8))                # 2 bits of entropy per byte
    yield (b"ACGT" * 100 + bytes(rng.integers(0, 256, 50, dtype=np.uint8))) * 100
    yield bytes(rng.integers(0, 256, 40000, dtype=np.uint8)) + bytes(rng.integers(0, 2, 25000, dtype=np.uint8))
    for n in (1, 2, 3, 5, 100, 1000, 30000, 65280):
        yield bytes(rng.integers(0, 40, n, dtype=np.uint8) + 33)            # quality-like
    p = np.array([2.0 ** -(i / 6) for i in range(256)])
    yield bytes(rng.choice(256, size=65280, p=p / p.sum()).astype(np.uint8))  # skewed: long codes
    p = np.array([2.0 ** -i for i in range(40)])
    yield bytes(rng.choice(40, size=65535, p=p / p.sum()).astype(np.uint8))   # probabilities down to 2^-39: the 15-bit limit
    blk = bytes(rng.integers(0, 256, 3000, dtype=np.uint8))
    yield blk + bytes(29000) + blk + bytes(100) + blk                        # far matches (distance 32100 and 3100)
    yield blk * 21                                                           # distance 3000 throughout
    yield bytes(rng.integers(0, 256, 32768 + 100, dtype=np.uint8)) * 1 + bytes(10)
    far = bytes(rng.integers(0, 256, 600, dtype=np.uint8))
    yield far + bytes(rng.integers(0, 4, 32768 - 600, dtype=np.uint8)) + far + far   # a candidate exactly 32768 back, one beyond
    from openge_b200 import bamio, synth
    raw = bamio.serialize_bam_stream(synth.make("C3", 0.002, seed=5))
    for k in range(0, min(len(raw), 6 * 65280), 65280):
        yield raw[k:k + 65280]                                              # real BAM bytes


def test_streams_inflate_back_with_zlib(lib):
    sizes = []
    for data in payloads():
        z = deflate(lib, data)
        d = zlib.decompressobj(-15)
        back = d.decompress(z)
        assert d.eof and d.unused_data == b"" and back == data, (len(data), len(z))
        assert len(z) <= len(data) + 5
        sizes.append((len(data), len(z), len(zlib.compress(data, 1)) - 6))
    # compression is in zlib level 1's class on the BAM blocks (the last six payloads)
    ours = sum(s[1] for s in sizes[-6:])
    theirs = sum(s[2] for s in sizes[-6:])
    assert ours < 1.15 * theirs, (ours, theirs)


def test_the_warp_form_compresses_like_zlib_level_1(lib):
    # what the 32 lanes do in lockstep, emulated lane by lane (tests/native/deflate_host.cpp): candidates come from earlier
    # windows only, and only consumed positions enter the hash table.  (Entering all 32 positions of a window cost 8 % on
    # synthetic and 30 % on real BAM data: a position that comes up again in the next window then finds itself.)
    import gzip
    from openge_b200 import bamio, synth
    sets = [bamio.serialize_bam_stream(synth.make(c, sc, seed=3)) for c, sc in (("C1", 0.02), ("C2", 0.0005), ("C3", 0.002), ("C4", 0.002))]
    real = os.path.join(ROOT, "tests", "golden", "208.yhet.bam")
    if os.path.exists(real):
        sets.append(gzip.open(real, "rb").read())
    for raw in sets:
        ours = theirs = 0
        for k in range(0, min(len(raw), 30 * 65280), 65280):
            d = raw[k:k + 65280]
            ours += lib.oge_test_emulate32(d, len(d))
            c = zlib.compressobj(1, zlib.DEFLATED, -15)
            theirs += len(c.compress(d) + c.flush())
        assert ours < 1.03 * theirs, (ours, theirs)


def test_crc_equals_zlib(lib):
    rng = np.random.default_rng(1)
    for n in (0, 1, 3, 4, 5, 127, 128, 129, 1000, 65279, 65280, 65535):
        data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        want = zlib.crc32(data)
        assert lib.oge_test_crc32(data, n, 1) == want
        assert lib.oge_test_crc32(data, n, 32) == want      # the 32-slice form with the x^(8 n) combination


def test_codes_are_complete_and_short(lib):
    rng = np.random.default_rng(2)
    cases = [np.array([1, 1] + [0] * 28), np.array([5] + [0] * 29), np.zeros(30, int), np.array([65535, 1] + [0] * 284),
             rng.integers(0, 3, 286), rng.integers(0, 1000, 286), (2.0 ** -np.arange(286 // 8 + 1).repeat(8)[:286] * 65000).astype(int),
             np.ones(286, int), np.array([1 << k for k in range(16)] + [0] * 14)]
    for f in cases:
        f = np.ascontiguousarray(f, dtype=np.uint32)
        lens = np.zeros(len(f), dtype=np.uint32)
        codes = np.zeros(len(f), dtype=np.uint32)
        bits = lib.oge_test_build_code(f.ctypes.data, len(f), lens.ctypes.data, codes.ctypes.data)
        assert bits >= 0
        used = lens > 0
        assert used.sum() >= 2 and lens.max() <= 15 and np.all(used[f > 0])
        assert sum(2.0 ** -int(l) for l in lens[used]) == 1.0      # complete
        # prefix-free: canonical codes in (length, symbol) order, read back from their bit-reversed form
        vals = sorted((int(l), int(format(int(c), "0%db" % l)[::-1], 2)) for l, c in zip(lens[used], codes[used]))
        strs = [format(v, "0%db" % l) for l, v in vals]
        assert all(not b.startswith(a) for i, a in enumerate(strs) for b in strs[i + 1:i + 2])
        # within a bit per symbol of the entropy bound (Shannon lengths, then the slack handed out)
        if f.sum() > 0 and (f > 0).sum() >= 2:
            p = f[f > 0] / f.sum()
            entropy_bits = float(-(f[f > 0] * np.log2(p)).sum())
            assert bits <= entropy_bits + f.sum() + 16


def test_compressor_fuzz_round_trips_through_zlib(lib):
    """Randomised structure (runs, repeats at every distance class, literals of varying entropy, text-like records): every
    stream the compressor writes must inflate back to its input with zlib; a few hundred blocks of every size class."""
    rng = np.random.default_rng(77)
    n_cases = 0
    for case in range(250):
        parts = []
        size = int(rng.integers(1, 65536))
        while sum(len(p) for p in parts) < size:
            kind = int(rng.integers(0, 5))
            if kind == 0:
                parts.append(rng.integers(0, int(rng.integers(1, 257)), int(rng.integers(1, 4000)), dtype=np.uint8).tobytes())
            elif kind == 1:
                parts.append(bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 3000)))
            elif kind == 2 and parts:
                prev = b"".join(parts)
                d = int(rng.integers(1, min(len(prev), 40000) + 1))
                l = int(rng.integers(3, 600))
                parts.append((prev[-d:] * (l // d + 1))[:l])
            elif kind == 3:
                p = np.array([2.0 ** -(i / float(rng.integers(2, 12))) for i in range(256)])
                parts.append(rng.choice(256, size=int(rng.integers(1, 5000)), p=p / p.sum()).astype(np.uint8).tobytes())
            else:
                words = [b"chr1", b"ACGT", b"read", b"\x00\x01\x02", b"IIIIIIII", b"RG:Z:rg1"]
                parts.append(b"".join(words[int(x)] for x in rng.integers(0, len(words), int(rng.integers(1, 300)))))
        data = b"".join(parts)[:65535]
        z = deflate(lib, data)
        d = zlib.decompressobj(-15)
        back = d.decompress(z)
        assert d.eof and d.unused_data == b"" and back == data, ("case", case, len(data), len(z))
        assert len(z) <= len(data) + 5
        assert lib.oge_test_crc32(data, len(data), 32) == zlib.crc32(data)
        n_cases += 1
    assert n_cases == 250
