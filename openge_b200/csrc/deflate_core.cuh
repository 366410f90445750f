// DEFLATE compressor for BGZF blocks: the body of the device kernel that writes the OUTPUT file's blocks (bgzf_deflate.cu),
// standing in for BgzfOutputStream::BgzfBlock::runJob's zlib call (reference util/bgzf_output_stream.cpp:59-144) where the
// caller accepts an output file that is identical after decompression rather than byte for byte (zlib's match finder is
// sequential by construction; the byte-identical writer stays on the host, bam_host.cpp).
//
// One WARP compresses one block of at most 65535 bytes:
//   parse    32 positions per step.  Every lane hashes the 4 bytes at its position, looks up the most recent earlier
//            position with the same hash (a per-warp table of 4096 16-bit positions in shared memory) and tries that
//            candidate and the position one byte back (runs); the first lane with a match wins, the lanes in front of it
//            are literals, all lanes extend the match together 32 bytes per step (up to 258), the parse continues behind it.
//            What the parse leaves is a list of sequences (literal run, match length, distance) and the symbol histograms.
//   codes    length-limited prefix codes from the histograms without a sort: Shannon lengths ceil(log2(total / f)) on
//            frequencies scaled to a total of at most 2^15 (so no length exceeds 15 and the Kraft sum is at most 1), then the
//            slack is handed out by shortening codes class by class until the code is complete (zlib's inflate rejects an
//            incomplete literal/length code).  Canonical codes, stored bit-reversed.  Code lengths go into the block header
//            verbatim under a fixed 4-bit code-length code (158 bytes per block, 0.24 %).
//   encode   32 literals per step: code lookup, warp prefix sum of the bit lengths, shared-memory atomicOr into a ring of
//            64 words, complete words flushed with coalesced stores.  One dynamic-Huffman block per BGZF block; a block
//            that would not shrink is stored.
//   crc      CRC-32 of the payload: every lane takes a contiguous slice byte by byte through a 256-entry table, the 32
//            partial CRCs are combined with x^(8 * bytes behind the slice) mod P (the identity zlib's crc32_combine uses).
//
// The same source compiles for the host with LANES = 1 (tests/native/deflate_host.cpp), which is how the logic is tested
// against zlib without a GPU (tests/test_deflate_core.py).  RFC 1951 is the specification restated here.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define OGE_DHD __host__ __device__ __forceinline__
#else
#define OGE_DHD static inline
#endif

namespace oge_deflate {

constexpr int HBITS = 12;                     // hash table: 4096 positions per warp
constexpr uint32_t MAX_BLOCK = 65535;         // a stored block's LEN field, and positions fit 16 bits with 0xFFFF = empty
constexpr uint32_t MAX_SEQ = MAX_BLOCK / 4 + 2;      // a match is at least 4 bytes long: at most this many sequences
constexpr uint32_t MIN_MATCH = 4, MAX_MATCH = 258, MAX_DIST = 32768;

struct Seq {            // one step of the parse: `run` literals, then a match (len == 0: none, the end of the block)
    uint32_t run;
    uint32_t len_dist;  // len | dist << 16 (dist 32768 fits: 0x8000)
};

struct Work {           // per-warp working set (9.8 KB of shared memory on the device)
    uint16_t htab[1 << HBITS];
    uint32_t lit[288];  // frequency -> (code length << 16 | bit-reversed code)
    uint32_t dst[32];
    uint32_t ring[64];  // bit output staging
    uint32_t cnt[32];   // per-length counts / next codes / quotas (lane 0)
};

// ---------------------------------------------------------------------------------------------- lane collectives
template <int LANES>
OGE_DHD void sync_lanes() {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) __syncwarp();
#endif
}
template <int LANES>
OGE_DHD uint32_t ballot(bool p) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) return __ballot_sync(0xFFFFFFFFu, p);
#endif
    return p ? 1u : 0u;
}
template <int LANES>
OGE_DHD uint32_t bcast(uint32_t v, int src) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) return __shfl_sync(0xFFFFFFFFu, v, src);
#endif
    (void) src;
    return v;
}
template <int LANES>
OGE_DHD uint32_t excl_scan(uint32_t v, int lane, uint32_t *total) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) {
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        *total = __shfl_sync(0xFFFFFFFFu, x, 31);
        return x - v;
    }
#endif
    (void) lane;
    *total = v;
    return 0;
}
template <int LANES>
OGE_DHD uint32_t xor_all(uint32_t v) {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) {
#pragma unroll
        for (int d = 16; d; d >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, d);
    }
#endif
    return v;
}
OGE_DHD void add_shared(uint32_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
OGE_DHD void or_shared(uint32_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
OGE_DHD int first_set(uint32_t v) {      // index of the lowest set bit (v != 0)
#if defined(__CUDA_ARCH__)
    return __ffs((int) v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
OGE_DHD int bit_len(uint32_t v) {        // 0 for 0
#if defined(__CUDA_ARCH__)
    return 32 - __clz((int) v);
#else
    return v ? 32 - __builtin_clz(v) : 0;
#endif
}
OGE_DHD uint32_t bit_reverse(uint32_t v, int n) {      // the low n bits of v, reversed
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// ---------------------------------------------------------------------------------------------- symbols (RFC 1951 3.2.5)
OGE_DHD uint32_t len_symbol(uint32_t len, uint32_t *ebits, uint32_t *eval) {      // 3..258 -> 257..285
    if (len == 258) { *ebits = 0; *eval = 0; return 285; }
    const uint32_t l = len - 3;
    if (l < 8) { *ebits = 0; *eval = 0; return 257 + l; }
    const uint32_t e = (uint32_t) bit_len(l) - 3;
    *ebits = e;
    *eval = l & ((1u << e) - 1);
    return 261 + 4 * e + ((l >> e) & 3);
}
OGE_DHD uint32_t dist_symbol(uint32_t dist, uint32_t *ebits, uint32_t *eval) {    // 1..32768 -> 0..29
    const uint32_t d = dist - 1;
    if (d < 4) { *ebits = 0; *eval = 0; return d; }
    const uint32_t e = (uint32_t) bit_len(d) - 2;
    *ebits = e;
    *eval = d & ((1u << e) - 1);
    return 2 * e + 2 + ((d >> e) & 1);
}

// 4 bytes at src + p, little endian; src is 4-byte aligned and readable 7 bytes beyond the last position asked for
OGE_DHD uint32_t load4(const uint8_t *src, uint32_t p) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(src + (p & ~3u));
    const uint32_t lo = w[0], hi = w[1], sh = (p & 3u) * 8;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}

// ---------------------------------------------------------------------------------------------- parse
// -> number of sequences written to seqs; histograms in W.lit / W.dst (end-of-block symbol included); *extra = extra bits
// of all matches.  Every lane returns the same values.
template <int LANES>
OGE_DHD uint32_t parse(const uint8_t *src, uint32_t n, Work &W, Seq *seqs, int lane, uint32_t *extra) {
    for (int i = lane; i < (1 << HBITS) / 2; i += LANES) reinterpret_cast<uint32_t *>(W.htab)[i] = 0xFFFFFFFFu;
    for (int i = lane; i < 288; i += LANES) W.lit[i] = 0;
    for (int i = lane; i < 32; i += LANES) W.dst[i] = 0;
    sync_lanes<LANES>();
    const uint32_t all = LANES == 32 ? 0xFFFFFFFFu : ((1u << (LANES & 31)) - 1);
    uint32_t pos = 0, lit_start = 0, nseq = 0, xbits = 0;
    while (pos < n) {
        const uint32_t p = pos + lane;
        const bool valid = p + MIN_MATCH <= n;
        const uint32_t w = valid ? load4(src, p) : 0;
        const uint32_t h = (w * 2654435761u) >> (32 - HBITS);
        const uint32_t cand = valid ? W.htab[h] : 0xFFFFu;
        uint32_t dist = 0;
        if (valid) {
            if (cand != 0xFFFFu && p - cand <= MAX_DIST && load4(src, cand) == w) dist = p - cand;
            else if (p >= 1 && load4(src, p - 1) == w) dist = 1;
        }
        const uint32_t found = ballot<LANES>(dist != 0);
        sync_lanes<LANES>();      // every lane has read its candidate: the table may change
        if (!found) {      // a window of literals
            if (valid) W.htab[h] = (uint16_t) p;      // lanes with equal hashes: any of them stays, all are recent
            if (p < n) add_shared(&W.lit[src[p]], 1);
            sync_lanes<LANES>();
            pos += LANES;
            continue;
        }
        const int f = first_set(found);
        const uint32_t mp = pos + f, d = bcast<LANES>(dist, f);
        if (lane < f) add_shared(&W.lit[src[p]], 1);
        uint32_t len = MIN_MATCH;
        for (;;) {      // all lanes extend the winner's match
            const uint32_t q = mp + len + lane;
            const bool ok = q < n && len + lane < MAX_MATCH && src[q] == src[q - d];
            const uint32_t b = ballot<LANES>(ok);
            if (b == all) { len += LANES; continue; }
            len += first_set(~b);
            break;
        }
        // only the positions this step consumes enter the table: the ones behind the match come up again in the next
        // window, and a position that finds ITSELF there has lost its real candidate
        if (valid && p < mp + len) W.htab[h] = (uint16_t) p;
        sync_lanes<LANES>();
        if (lane == 0) {
            seqs[nseq].run = mp - lit_start;
            seqs[nseq].len_dist = len | (d << 16);
            uint32_t eb, ev;
            add_shared(&W.lit[len_symbol(len, &eb, &ev)], 1);
            xbits += eb;
            add_shared(&W.dst[dist_symbol(d, &eb, &ev)], 1);
            xbits += eb;
        }
        nseq++;
        lit_start = pos = mp + len;
    }
    if (lane == 0) {
        seqs[nseq].run = n - lit_start;
        seqs[nseq].len_dist = 0;
        add_shared(&W.lit[256], 1);
    }
    nseq++;
    sync_lanes<LANES>();
    *extra = bcast<LANES>(xbits, 0);
    return nseq;
}

// ---------------------------------------------------------------------------------------------- codes (one lane)
// tab[0..nsym): frequencies in, (length << 16 | bit-reversed canonical code) out, 0 for unused symbols.  A complete
// prefix code with lengths <= 15 over at least two symbols.  -> bits the coded symbols take (sum of f * length), or ~0 if
// the slack could not be handed out (never seen; the caller stores the block then).
OGE_DHD uint64_t build_code(uint32_t *tab, int nsym, uint32_t *cnt, int *n_used_out) {
    int used = 0;
    for (int s = 0; s < nsym; s++) used += tab[s] != 0;
    for (int s = 0; s < nsym && used < 2; s++)      // zlib's decoder wants a complete code: at least two symbols
        if (tab[s] == 0) { tab[s] = 1; used++; }
    // frequencies scaled so that their sum is at most 2^15: every Shannon length then is at most 15
    uint32_t shift = 0, total;
    for (;; shift++) {
        total = 0;
        for (int s = 0; s < nsym; s++)
            if (tab[s]) total += (tab[s] + (1u << shift) - 1) >> shift;
        if (total <= 32768) break;
    }
    // length = ceil(log2(total / f)): the Kraft sum is at most 1; kept in the top byte next to the (24-bit) frequency
    uint32_t kraft = 0;      // in units of 2^-15
    for (int s = 0; s < nsym; s++) {
        const uint32_t f = tab[s];
        if (!f) continue;
        const uint32_t fs = (f + (1u << shift) - 1) >> shift;
        uint32_t l = 1;
        while ((fs << l) < total) l++;
        tab[s] = f | (l << 24);
        kraft += 1u << (15 - l);
    }
    // the slack goes to shorter codes, class by class from the frequent end, until the code is complete
    uint32_t slack = 32768 - kraft;
    for (int guard = 0; slack && guard < 64; guard++) {
        for (int c = 0; c < 16; c++) cnt[c] = 0;
        for (int s = 0; s < nsym; s++)
            if (tab[s]) cnt[tab[s] >> 24]++;
        for (int c = 2; c < 16; c++) {      // cnt[c] becomes the number of codes of length c that get one bit shorter
            const uint32_t k = 1u << (15 - c), q = slack / k < cnt[c] ? slack / k : cnt[c];
            cnt[c] = q;
            slack -= q * k;
        }
        cnt[1] = 0;
        for (int s = 0; s < nsym; s++) {
            const uint32_t c = tab[s] >> 24;
            if (c && cnt[c]) { cnt[c]--; tab[s] -= 1u << 24; }
        }
    }
    if (slack) return ~0ull;
    // canonical codes (RFC 1951 3.2.2), bit-reversed for the LSB-first bit stream
    for (int c = 0; c < 16; c++) cnt[c] = 0;
    for (int s = 0; s < nsym; s++)
        if (tab[s]) cnt[tab[s] >> 24]++;
    uint32_t code = 0, prev = 0;
    for (int c = 1; c < 16; c++) {
        code = (code + prev) << 1;
        prev = cnt[c];
        cnt[c] = code;      // next code of this length
    }
    uint64_t bits = 0;
    for (int s = 0; s < nsym; s++) {
        const uint32_t l = tab[s] >> 24;
        if (!l) continue;
        bits += (uint64_t) (tab[s] & 0xFFFFFFu) * l;
        tab[s] = (l << 16) | bit_reverse(cnt[l]++, (int) l);
    }
    *n_used_out = used;
    return bits;
}

// ---------------------------------------------------------------------------------------------- bit output
struct BitOut {
    uint32_t *out32;      // 4-byte aligned destination
    uint32_t bitpos;      // bits written so far
    uint32_t flushed;     // words already stored
};

// every lane contributes nbits (0..28) bits of code, in lane order
template <int LANES>
OGE_DHD void put(Work &W, BitOut &B, uint32_t code, uint32_t nbits, int lane) {
    uint32_t total;
    const uint32_t bp = B.bitpos + excl_scan<LANES>(nbits, lane, &total);
    if (nbits) {
        const uint32_t w = bp >> 5, sh = bp & 31;
        or_shared(&W.ring[w & 63], code << sh);
        if (sh + nbits > 32) or_shared(&W.ring[(w + 1) & 63], code >> (32 - sh));
    }
    B.bitpos += total;
    sync_lanes<LANES>();
    const uint32_t full = B.bitpos >> 5;
    for (uint32_t i = B.flushed + lane; i < full; i += LANES) {
        B.out32[i] = W.ring[i & 63];
        W.ring[i & 63] = 0;
    }
    B.flushed = full;
    sync_lanes<LANES>();
}

// -> bytes of the stream
template <int LANES>
OGE_DHD uint32_t finish(Work &W, BitOut &B, int lane) {
    if ((B.bitpos & 31) && lane == 0) {
        B.out32[B.flushed] = W.ring[B.flushed & 63];
        W.ring[B.flushed & 63] = 0;
    }
    sync_lanes<LANES>();
    return (B.bitpos + 7) >> 3;
}

// ---------------------------------------------------------------------------------------------- one block
// src: n bytes (1..65535), 4-byte aligned, readable to src + n + 7.  out: 4-byte aligned, room for n + 16 bytes.
// -> bytes of the raw deflate stream written to out (one final block: dynamic Huffman, or stored when that is not smaller).
template <int LANES>
OGE_DHD uint32_t deflate_block(const uint8_t *src, uint32_t n, uint8_t *out, Work &W, Seq *seqs, int lane) {
    uint32_t extra;
    const uint32_t nseq = parse<LANES>(src, n, W, seqs, lane, &extra);
    // code tables (lane 0), sizes
    uint32_t nlit = 257, ndist = 1;
    uint64_t bits = 0;
    if (lane == 0) {
        int used;
        for (int s = 257; s < 286; s++)
            if (W.lit[s]) nlit = s + 1;
        for (int s = 0; s < 30; s++)
            if (W.dst[s]) ndist = s + 1;
        const uint64_t b1 = build_code(W.lit, 286, W.cnt, &used);
        const uint64_t b2 = build_code(W.dst, 30, W.cnt, &used);
        if (ndist < 2) ndist = 2;      // build_code made symbols 0 and 1 the code when no or one distance was used
        for (int s = 0; s < 30; s++)
            if (W.dst[s] && (uint32_t) s + 1 > ndist) ndist = s + 1;
        bits = (b1 == ~0ull || b2 == ~0ull) ? ~0ull : 17 + 57 + 4ull * (nlit + ndist) + b1 + b2 + extra;
        W.cnt[16] = nlit;
        W.cnt[17] = ndist;
        W.cnt[18] = bits == ~0ull || (bits + 7) / 8 >= (uint64_t) n + 5 ? 1u : 0u;      // 1: store the block
    }
    sync_lanes<LANES>();
    nlit = W.cnt[16];
    ndist = W.cnt[17];
    const bool stored = W.cnt[18] != 0;
    sync_lanes<LANES>();
    if (stored) {
        if (lane == 0) {
            out[0] = 1;      // BFINAL = 1, BTYPE = 00, padding
            out[1] = (uint8_t) n; out[2] = (uint8_t) (n >> 8);
            out[3] = (uint8_t) ~n; out[4] = (uint8_t) (~n >> 8);
        }
        for (uint32_t i = lane; i < n; i += LANES) out[5 + i] = src[i];
        sync_lanes<LANES>();
        return n + 5;
    }
    for (int i = lane; i < 64; i += LANES) W.ring[i] = 0;
    sync_lanes<LANES>();
    BitOut B;
    B.out32 = reinterpret_cast<uint32_t *>(out);
    B.bitpos = 0;
    B.flushed = 0;
    // header: BFINAL 1, BTYPE 10, HLIT, HDIST, HCLEN = 15 (all 19 code-length code lengths follow)
    put<LANES>(W, B, 1u | (2u << 1) | ((nlit - 257) << 3) | ((ndist - 1) << 8) | (15u << 13), lane == 0 ? 17 : 0, lane);
    // code-length code: lengths 0..15 get 4 bits each, the repeat symbols 16, 17, 18 (first in the transmission order) none
    for (int i = 0; i < 19; i += LANES) {
        const int k = i + lane;
        put<LANES>(W, B, k < 3 ? 0u : 4u, k < 19 ? 3 : 0, lane);
    }
    // the code lengths themselves: canonical 4-bit code of v is v, sent most significant bit first
    for (uint32_t i = 0; i < nlit + ndist; i += LANES) {
        const uint32_t k = i + lane;
        const uint32_t l = k < nlit ? W.lit[k] >> 16 : (k < nlit + ndist ? W.dst[k - nlit] >> 16 : 0);
        put<LANES>(W, B, bit_reverse(l, 4), k < nlit + ndist ? 4 : 0, lane);
    }
    // the sequences
    uint32_t at = 0;
    for (uint32_t s = 0; s < nseq; s++) {
        const uint32_t run = seqs[s].run, ld = seqs[s].len_dist;
        for (uint32_t i = 0; i < run; i += LANES) {
            const uint32_t k = i + lane;
            const uint32_t e = k < run ? W.lit[src[at + k]] : 0;
            put<LANES>(W, B, e & 0xFFFFu, e >> 16, lane);
        }
        at += run;
        if (ld) {
            const uint32_t len = ld & 0xFFFFu, d = ld >> 16;
            uint32_t eb, ev;
            const uint32_t ls = len_symbol(len, &eb, &ev), le = W.lit[ls];
            const uint32_t lcode = (le & 0xFFFFu) | (ev << (le >> 16)), lbits = (le >> 16) + eb;
            const uint32_t ds = dist_symbol(d, &eb, &ev), de = W.dst[ds];
            const uint32_t dcode = (de & 0xFFFFu) | (ev << (de >> 16)), dbits = (de >> 16) + eb;
            if (LANES > 1) {
                put<LANES>(W, B, lane == 0 ? lcode : dcode, lane == 0 ? lbits : (lane == 1 ? dbits : 0), lane);
            } else {
                put<LANES>(W, B, lcode, lbits, lane);
                put<LANES>(W, B, dcode, dbits, lane);
            }
            at += len;
        }
    }
    put<LANES>(W, B, W.lit[256] & 0xFFFFu, lane == 0 ? W.lit[256] >> 16 : 0, lane);
    return finish<LANES>(W, B, lane);
}

// ---------------------------------------------------------------------------------------------- CRC-32 (IEEE, reflected)
constexpr uint32_t CRC_POLY = 0xEDB88320u;

OGE_DHD uint32_t crc_table_entry(uint32_t i) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (CRC_POLY & (0u - (c & 1u)));
    return c;
}
// a * b mod P, polynomials over GF(2) in the reflected representation (bit 31 is the coefficient of x^0)
OGE_DHD uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 31; i >= 0; i--) {
        if ((a >> i) & 1u) p ^= b;
        b = (b >> 1) ^ (CRC_POLY & (0u - (b & 1u)));
    }
    return p;
}
// x^(8 * nbytes) mod P
OGE_DHD uint32_t x_pow_8n(uint32_t nbytes) {
    uint32_t r = 0x80000000u, p = 0x00800000u;      // x^0, x^8
    while (nbytes) {
        if (nbytes & 1u) r = gf2_mulmod(r, p);
        p = gf2_mulmod(p, p);
        nbytes >>= 1;
    }
    return r;
}
// crc32 of src[0..n): lane L takes the L-th slice; crc(A || B) = crc(A) * x^(8 |B|) + crc(B) for finished CRCs
template <int LANES>
OGE_DHD uint32_t crc32_block(const uint8_t *src, uint32_t n, const uint32_t *table256, int lane) {
    const uint32_t slice = ((n + LANES * 4 - 1) / (LANES * 4)) * 4;
    const uint32_t a = slice * lane < n ? slice * lane : n, b = a + slice < n ? a + slice : n;
    uint32_t c = 0xFFFFFFFFu;
    for (uint32_t i = a; i < b; i++) c = table256[(c ^ src[i]) & 0xFFu] ^ (c >> 8);
    c = b > a ? ~c : 0u;      // an empty slice has CRC 0
    return xor_all<LANES>(gf2_mulmod(x_pow_8n(n - b), c));
}

}  // namespace oge_deflate
