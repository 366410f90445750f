// DEFLATE (RFC 1951) decoder for one BGZF block, written for a warp: the decode state (bit buffer, positions) is
// kept redundantly in every lane -- all lanes execute the same instruction stream, loads of the compressed words
// and of the table entries are warp-wide broadcasts -- so there is no divergence, and when a match has to be
// copied the 32 lanes copy it together.  Huffman tables live in a per-warp block of shared memory.
//
// Replaces the zlib call of BgzfInputStream::BgzfBlock::decompress (reference util/bgzf_input_stream.cpp:100-138:
// raw inflate, window 15, whole block in one call).  Like the reference, the CRC of the block is not checked; the
// inflated length is (:137).
//
// LANES is a template parameter.  LANES == 32: the warp-cooperative form described above.  LANES == 1: one THREAD per
// block -- 32 independent streams per warp in SIMT lockstep, every lane doing useful decoding work; this is the form the
// device uses by default (bgzf_inflate.cu: the warp-cooperative kernel turned out to be issue-bound, 31 of 32 lanes
// repeating the same decode), and also the form that compiles for the host: tests/native/inflate_host.cpp runs it
// against zlib on the CPU, so the decoding logic is checked without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OGE_HD __host__ __device__ __forceinline__
#define OGE_HD_NOINLINE __host__ __device__ __noinline__
#else
#define OGE_HD static inline
#define OGE_HD_NOINLINE static
#endif

namespace oge_inflate {

constexpr int LIT_BITS = 10;      // primary table widths (warp form): codes up to this length decode with one lookup,
constexpr int DIST_BITS = 8;      // longer ones (rare) walk the canonical code bit by bit
constexpr int CL_BITS = 7;

template <int LANES>
OGE_HD void sync_lanes() {
#if defined(__CUDA_ARCH__)
    if (LANES > 1) __syncwarp();
#endif
}

enum {
    INF_OK = 0,
    INF_ERR_BTYPE = 1,        // reserved block type
    INF_ERR_STORED = 2,       // LEN != ~NLEN
    INF_ERR_LENGTHS = 3,      // bad code-length sequence / over-subscribed code
    INF_ERR_SYMBOL = 4,       // no code matches / invalid literal-length or distance symbol
    INF_ERR_DISTANCE = 5,     // match reaches before the start of the block
    INF_ERR_OVERRUN = 6,      // more output than ISIZE, or input exhausted
    INF_ERR_SHORT = 7         // stream ended before ISIZE bytes
};

// Per-warp working set (about 3.9 KB).
struct Tables {
    uint16_t lit_tab[1 << LIT_BITS];
    uint16_t dist_tab[1 << DIST_BITS];
    uint16_t cl_tab[1 << CL_BITS];
    uint16_t lit_sym[288], dist_sym[32], cl_sym[20];
    uint16_t lit_cnt[16], dist_cnt[16], cl_cnt[16];
    uint8_t lens[320];
    int32_t status;      // scratch for lane 0's verdicts
};

// Where a decoder's tables live (one contiguous Tables block per warp, or hot tables in shared memory and the rest
// in thread-local memory for the thread-per-block form).
struct TablesRef {
    uint16_t *lit_tab, *dist_tab, *cl_tab;
    uint16_t *lit_sym, *dist_sym, *cl_sym;
    uint16_t *lit_cnt, *dist_cnt, *cl_cnt;
    uint8_t *lens;        // 320
    int32_t *status;
};

OGE_HD TablesRef tables_ref(Tables *T) {
    TablesRef R;
    R.lit_tab = T->lit_tab; R.dist_tab = T->dist_tab; R.cl_tab = T->cl_tab;
    R.lit_sym = T->lit_sym; R.dist_sym = T->dist_sym; R.cl_sym = T->cl_sym;
    R.lit_cnt = T->lit_cnt; R.dist_cnt = T->dist_cnt; R.cl_cnt = T->cl_cnt;
    R.lens = T->lens;
    R.status = &T->status;
    return R;
}

struct BitReader {
    const uint32_t *wp;      // next aligned word
    const uint32_t *end;     // first word that must not be read
    uint64_t buf;
    int cnt;
};

OGE_HD void br_init(BitReader &r, const uint8_t *in, uint32_t len) {
    const uintptr_t a = (uintptr_t) in;
    r.wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t) 3);
    r.end = reinterpret_cast<const uint32_t *>((a + len + 3 + 8) & ~(uintptr_t) 3);      // the 8-byte gzip footer follows the data
    const int skip = (int) (a & 3) * 8;
    r.buf = (uint64_t) (*r.wp++) >> skip;
    r.cnt = 32 - skip;
}

OGE_HD void br_refill(BitReader &r) {      // afterwards at least 33 bits are valid (zero-padded past the end)
    if (r.cnt <= 32) {
        const uint32_t w = r.wp < r.end ? *r.wp : 0u;
        r.wp++;
        r.buf |= (uint64_t) w << r.cnt;
        r.cnt += 32;
    }
}

OGE_HD uint32_t br_peek(const BitReader &r, int n) { return (uint32_t) r.buf & ((1u << n) - 1u); }
OGE_HD void br_drop(BitReader &r, int n) { r.buf >>= n; r.cnt -= n; }
OGE_HD uint32_t br_take(BitReader &r, int n) { const uint32_t v = br_peek(r, n); br_drop(r, n); return v; }

OGE_HD uint32_t bit_reverse(uint32_t v, int n) {
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// Canonical Huffman tables from code lengths (RFC 1951 3.2.2).  cnt[l] = codes of length l, sym[] = symbols in
// canonical order; tab[] = primary lookup on the next `bits` stream bits: symbol | length << 9, or 0 for "longer
// than `bits`, or no such code".  Lane 0 counts and orders, all lanes fill the lookup table.
// Returns (in every lane) 0, or INF_ERR_LENGTHS for a set zlib's inflate_table rejects: over-subscribed, or incomplete --
// except that a literal/length or distance code (not the code-length code: is_cl) may consist of ONE code of length 1, and
// that a set without any code passes here and fails when a symbol is asked of it (inftrees.c: "no symbols, but wait for
// decoding to report error").  The reference inflates with zlib (util/bgzf_input_stream.cpp:110-128), so this is its verdict.
template <int LANES>
OGE_HD_NOINLINE int build_tables(const uint8_t *lens, int n, int bits, uint16_t *tab, uint16_t *sym, uint16_t *cnt, int32_t *status, int lane,
                                 bool is_cl = false) {
    sync_lanes<LANES>();
    if (lane == 0) {
        for (int l = 0; l < 16; l++) cnt[l] = 0;
        for (int i = 0; i < n; i++) cnt[lens[i]]++;
        int left = 1, bad = 0, max_len = 0;
        for (int l = 1; l < 16; l++) {
            left <<= 1;
            left -= cnt[l];
            if (left < 0) bad = 1;
            if (cnt[l]) max_len = l;
        }
        if (left > 0 && max_len > 0 && (is_cl || max_len != 1)) bad = 1;      // incomplete
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t) (offs[l] + cnt[l]);
        for (int i = 0; i < n; i++)
            if (lens[i]) sym[offs[lens[i]]++] = (uint16_t) i;
        *status = bad ? INF_ERR_LENGTHS : 0;
    }
    for (int k = lane; k < (1 << bits); k += LANES) tab[k] = 0;
    sync_lanes<LANES>();
    if (*status) return *status;
    // first code and first canonical index of every length
    uint32_t first_code[16], first_idx[16];
    {
        uint32_t code = 0, idx = 0;
        first_code[0] = first_idx[0] = 0;
        for (int l = 1; l < 16; l++) {
            code = (code + (l > 1 ? cnt[l - 1] : 0)) << 1;
            first_code[l] = code;
            first_idx[l] = idx;
            idx += cnt[l];
        }
    }
    int total = 0;
    for (int l = 1; l < 16; l++) total += cnt[l];
    for (int i = lane; i < total; i += LANES) {
        const uint32_t s = sym[i];
        const int l = lens[s];
        if (l > bits) continue;
        const uint32_t code = first_code[l] + ((uint32_t) i - first_idx[l]);
        const uint16_t e = (uint16_t) (s | ((uint32_t) l << 9));
        for (uint32_t k = bit_reverse(code, l); k < (1u << bits); k += 1u << l) tab[k] = e;
    }
    sync_lanes<LANES>();
    return 0;
}

// One symbol: primary lookup, else the canonical walk (puff-style) over the peeked bits.  -1 = no code matches.
OGE_HD int decode_symbol(BitReader &r, const uint16_t *tab, int bits, const uint16_t *sym, const uint16_t *cnt) {
    const uint32_t e = tab[br_peek(r, bits)];
    if (e) {
        br_drop(r, (int) (e >> 9));
        return (int) (e & 0x1FFu);
    }
    int code = 0, first = 0, index = 0;
    const uint32_t window = br_peek(r, 15);
    for (int l = 1; l <= 15; l++) {
        code |= (int) ((window >> (l - 1)) & 1u);
        const int c = cnt[l];
        if (code - c < first) {
            br_drop(r, l);
            return sym[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Order in which the code-length code lengths are stored (RFC 1951 3.2.7), packed 5 bits each.
OGE_HD int cl_order(int i) {
    // 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
    const uint64_t lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) | (9ull << 30) | (6ull << 35) |
                        (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
    const uint64_t hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
    return (int) ((i < 12 ? lo >> (5 * i) : hi >> (5 * (i - 12))) & 31u);
}

// One deflate block header (RFC 1951 3.2.3): BFINAL, BTYPE; a stored block is copied here and then; for a Huffman
// block the code lengths are read (3.2.6 / 3.2.7) and the lookup tables built.  *huff = a symbol stream follows.
// (Inlined on purpose: taken by reference into a real call, the bit reader would live in local memory and every
// symbol of the hot loop would pay for it -- measured: 750 M local loads, 22 GB of DRAM reads for a 1 GB file.)
template <int LANES, int LB, int DB>
OGE_HD int block_header(BitReader &r, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, uint32_t &pos,
                                 const TablesRef TR, int lane, uint32_t *last_out, bool *huff) {
    br_refill(r);
    const uint32_t last = br_take(r, 1), type = br_take(r, 2);
    *last_out = last;
    *huff = false;
    if (type == 0) {      // stored
        br_drop(r, r.cnt & 7);
        br_refill(r);
        const uint32_t len = br_take(r, 16);
        br_refill(r);
        const uint32_t nlen = br_take(r, 16);
        if ((len ^ nlen) != 0xFFFFu) return INF_ERR_STORED;
        if (pos + len > out_len) return INF_ERR_OVERRUN;
        // byte position of the reader: words consumed minus the bytes still buffered
        const uint8_t *src = reinterpret_cast<const uint8_t *>(r.wp) - (r.cnt >> 3);
        if (src + len > in + in_len) return INF_ERR_OVERRUN;
        for (uint32_t j = lane; j < len; j += LANES) out[pos + j] = src[j];
        pos += len;
        br_init(r, src + len, (uint32_t) (in + in_len - (src + len)));
        return INF_OK;
    }
    if (type == 3) return INF_ERR_BTYPE;
    int n_lit, n_dist;
    if (type == 1) {      // fixed code (RFC 1951 3.2.6)
        for (int i = lane; i < 288; i += LANES) TR.lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
        for (int i = lane; i < 32; i += LANES) TR.lens[288 + i] = 5;      // 32 codes of 5 bits: a complete code; 30 and 31 are invalid when met
        n_lit = 288;
        n_dist = 32;
    } else {              // dynamic code (3.2.7)
        br_refill(r);
        n_lit = (int) br_take(r, 5) + 257;
        n_dist = (int) br_take(r, 5) + 1;
        const int n_cl = (int) br_take(r, 4) + 4;
        if (n_lit > 286 || n_dist > 30) return INF_ERR_LENGTHS;
        uint8_t *cl = TR.lens + 300;      // 19 code-length code lengths, parked at the end of lens[]
        sync_lanes<LANES>();
        for (int i = lane; i < 19; i += LANES) cl[i] = 0;
        sync_lanes<LANES>();
        for (int i = 0; i < n_cl; i++) {
            br_refill(r);
            const uint32_t v = br_take(r, 3);
            if (lane == 0) cl[cl_order(i)] = (uint8_t) v;
        }
        int rc = build_tables<LANES>(cl, 19, CL_BITS, TR.cl_tab, TR.cl_sym, TR.cl_cnt, TR.status, lane, true);
        if (rc) return rc;
        int i = 0;
        uint32_t prev = 0;
        while (i < n_lit + n_dist) {
            br_refill(r);
            const int s = decode_symbol(r, TR.cl_tab, CL_BITS, TR.cl_sym, TR.cl_cnt);
            if (s < 0) return INF_ERR_SYMBOL;
            if (s < 16) {
                if (lane == 0) TR.lens[i] = (uint8_t) s;
                prev = (uint32_t) s;
                i++;
            } else {
                uint32_t v = 0;
                int rep;
                br_refill(r);
                if (s == 16) {
                    if (i == 0) return INF_ERR_LENGTHS;
                    v = prev;
                    rep = 3 + (int) br_take(r, 2);
                } else if (s == 17) {
                    rep = 3 + (int) br_take(r, 3);
                } else {
                    rep = 11 + (int) br_take(r, 7);
                }
                if (i + rep > n_lit + n_dist) return INF_ERR_LENGTHS;
                for (int k = lane; k < rep; k += LANES) TR.lens[i + k] = (uint8_t) v;
                i += rep;
                if (s != 16) prev = 0;
            }
        }
        sync_lanes<LANES>();
        if (TR.lens[256] == 0) return INF_ERR_LENGTHS;      // no end-of-block code
        // distance lengths follow the literal/length ones in the stream; give them their own start
        if (n_lit < 288) {
            sync_lanes<LANES>();
            uint8_t d[32];
            for (int k = 0; k < n_dist; k++) d[k] = TR.lens[n_lit + k];
            sync_lanes<LANES>();
            for (int k = lane; k < n_dist; k += LANES) TR.lens[288 + k] = d[k];
            sync_lanes<LANES>();
        }
    }
    int rc = build_tables<LANES>(TR.lens, n_lit, LB, TR.lit_tab, TR.lit_sym, TR.lit_cnt, TR.status, lane);
    if (rc) return rc;
    rc = build_tables<LANES>(TR.lens + 288, n_dist, DB, TR.dist_tab, TR.dist_sym, TR.dist_cnt, TR.status, lane);
    if (rc) return rc;
    *huff = true;
    return INF_OK;
}

// Length and distance of a match whose literal/length symbol s (257..285) has been decoded (RFC 1951 3.2.5: bases and
// extra bits by formula).  Returns 0 or an error.
template <int DB>
OGE_HD int match_params(BitReader &r, int s, const TablesRef T, uint32_t *len_out, uint32_t *dist_out) {
    if (s > 285) return INF_ERR_SYMBOL;
    uint32_t len;
    br_refill(r);
    if (s < 265) len = (uint32_t) s - 254;
    else if (s == 285) len = 258;
    else {
        const int e = (s - 261) >> 2;
        len = ((4u + (uint32_t) ((s - 261) & 3)) << e) + 3 + br_take(r, e);
    }
    br_refill(r);
    const int ds = decode_symbol(r, T.dist_tab, DB, T.dist_sym, T.dist_cnt);
    if (ds < 0 || ds > 29) return INF_ERR_SYMBOL;
    uint32_t dist;
    br_refill(r);
    if (ds < 4) dist = (uint32_t) ds + 1;
    else {
        const int e = (ds >> 1) - 1;
        dist = ((2u + (uint32_t) (ds & 1)) << e) + 1 + br_take(r, e);
    }
    *len_out = len;
    *dist_out = dist;
    return INF_OK;
}

// Inflates one raw deflate stream of `in_len` bytes (a BGZF block's payload) into exactly `out_len` bytes.
// Called by all lanes of a warp with identical arguments (LANES == 32), or by one thread (LANES == 1).  Readable slack:
// up to 12 bytes past in + in_len (the gzip footer and the next block header are there).
template <int LANES, int LB, int DB>
OGE_HD_NOINLINE int inflate_block(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, const TablesRef TR, int lane) {
    // the table pointers by value, in registers: reached through a reference they would be reloaded from local memory
    // for every symbol (the byte stores to `out` could alias them as far as the compiler knows)
    BitReader r;
    br_init(r, in, in_len);
    uint32_t pos = 0;
    const uint32_t *in_stop = r.end + 3;      // reading zeros beyond the data means the stream is broken
    while (true) {
        uint32_t last;
        bool huff;
        const int hrc = block_header<LANES, LB, DB>(r, in, in_len, out, out_len, pos, TR, lane, &last, &huff);
        if (hrc) return hrc;
        if (huff) {
            while (true) {      // ---- the symbol loop
                br_refill(r);
                int s = decode_symbol(r, TR.lit_tab, LB, TR.lit_sym, TR.lit_cnt);
                if (s < 256) {
                    if (s < 0) return INF_ERR_SYMBOL;
                    if (pos >= out_len) return INF_ERR_OVERRUN;
                    if (lane == 0) out[pos] = (uint8_t) s;
                    pos++;
                    continue;
                }
                if (s == 256) break;
                uint32_t len, dist;
                const int mrc = match_params<DB>(r, s, TR, &len, &dist);
                if (mrc) return mrc;
                if (dist > pos) return INF_ERR_DISTANCE;
                if (pos + len > out_len) return INF_ERR_OVERRUN;
                sync_lanes<LANES>();      // the bytes this match copies may have been written by other lanes
                if (LANES == 1 || dist >= len) {      // one lane copies front to back, which is the definition of an overlapping match
                    for (uint32_t j = lane; j < len; j += LANES) out[pos + j] = out[pos - dist + j];
                } else {                 // overlapping: the source repeats with period dist
                    for (uint32_t j = lane; j < len; j += LANES) out[pos + j] = out[pos - dist + (j % dist)];
                }
                pos += len;
                if (r.wp > in_stop) return INF_ERR_OVERRUN;
            }
        }
        if (last) break;
        if (r.wp > in_stop) return INF_ERR_OVERRUN;
    }
    sync_lanes<LANES>();
    return pos == out_len ? INF_OK : INF_ERR_SHORT;
}

// The same decoder for 32 INDEPENDENT streams held by the 32 lanes of a warp, written as a state machine so that the
// lanes stay converged: every trip of the loop, every lane that still has work does ONE step of its own stream (a
// literal, a match, or a block header) and all lanes meet again at the end of the trip.  Written the obvious way (each
// lane running inflate_block<1> on its own) the lanes drift apart after the first data-dependent branch and the warp
// executes one lane at a time: measured, 64 G warp-instructions for 1.5 GB -- as many as the warp-cooperative form.
// `active` = this lane has a stream; lanes without one only keep the others company.
template <int LB, int DB>
OGE_HD_NOINLINE int inflate_lockstep(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, const TablesRef TR, bool active) {
    enum { ST_HEADER = 0, ST_DECODE = 1, ST_DONE = 2 };
    BitReader r;
    r.wp = r.end = nullptr;
    r.buf = 0;
    r.cnt = 0;
    const uint32_t *in_stop = nullptr;
    uint32_t pos = 0, last = 0;
    int rc = INF_OK, state = ST_DONE;
    if (active) {
        br_init(r, in, in_len);
        in_stop = r.end + 3;
        state = ST_HEADER;
    }
    while (true) {
#if defined(__CUDA_ARCH__)
        if (!__any_sync(0xFFFFFFFFu, state != ST_DONE)) break;
#else
        if (state == ST_DONE) break;
#endif
        if (state == ST_DECODE) {
            br_refill(r);
            const int s = decode_symbol(r, TR.lit_tab, LB, TR.lit_sym, TR.lit_cnt);
            if (s < 256) {
                if (s < 0) { rc = INF_ERR_SYMBOL; state = ST_DONE; }
                else if (pos >= out_len) { rc = INF_ERR_OVERRUN; state = ST_DONE; }
                else out[pos++] = (uint8_t) s;
            } else if (s == 256) {
                state = last ? ST_DONE : ST_HEADER;
                if (r.wp > in_stop) { rc = INF_ERR_OVERRUN; state = ST_DONE; }
            } else {
                uint32_t len = 0, dist = 0;
                rc = match_params<DB>(r, s, TR, &len, &dist);
                if (!rc && dist > pos) rc = INF_ERR_DISTANCE;
                if (!rc && pos + len > out_len) rc = INF_ERR_OVERRUN;
                if (!rc && r.wp > in_stop) rc = INF_ERR_OVERRUN;
                if (rc) state = ST_DONE;
                else {
                    for (uint32_t j = 0; j < len; j++) out[pos + j] = out[pos - dist + j];      // front to back: overlap-safe
                    pos += len;
                }
            }
        } else if (state == ST_HEADER) {
            bool huff = false;
            rc = block_header<1, LB, DB>(r, in, in_len, out, out_len, pos, TR, 0, &last, &huff);
            if (rc) state = ST_DONE;
            else if (huff) state = ST_DECODE;
            else state = last ? ST_DONE : ST_HEADER;
            if (!rc && r.wp > in_stop) { rc = INF_ERR_OVERRUN; state = ST_DONE; }
        }
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
    }
    if (active && rc == INF_OK && pos != out_len) rc = INF_ERR_SHORT;
    return rc;
}

}  // namespace oge_inflate
