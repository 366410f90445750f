// TEST HARNESS: compiles the product's DEFLATE compressor (openge_b200/csrc/deflate_core.cuh, the body of the device
// kernel that writes BGZF blocks) for the host with one lane, so that tests/test_deflate_core.py can check its output
// with zlib without a GPU.
#include <stdlib.h>
#include <string.h>

#include "deflate_core.cuh"

// -> bytes of the raw deflate stream in out (capacity n + 16), or -1
extern "C" int oge_test_deflate_block(const unsigned char *in, unsigned n, unsigned char *out) {
    using namespace oge_deflate;
    if (n == 0 || n > MAX_BLOCK) return -1;
    Work *W = (Work *) calloc(1, sizeof(Work));
    Seq *seqs = (Seq *) calloc(MAX_SEQ, sizeof(Seq));
    unsigned char *src = (unsigned char *) aligned_alloc(16, ((size_t) n + 8 + 15) & ~(size_t) 15);
    unsigned char *dst = (unsigned char *) aligned_alloc(16, ((size_t) n + 16 + 15) & ~(size_t) 15);
    memcpy(src, in, n);
    memset(src + n, 0xAB, 8);      // what lies behind the block must not matter
    const uint32_t bytes = deflate_block<1>(src, n, dst, *W, seqs, 0);
    memcpy(out, dst, bytes);
    free(W); free(seqs); free(src); free(dst);
    return (int) bytes;
}

extern "C" unsigned oge_test_crc32(const unsigned char *in, unsigned n, int lanes) {
    using namespace oge_deflate;
    uint32_t table[256];
    for (uint32_t i = 0; i < 256; i++) table[i] = crc_table_entry(i);
    if (lanes <= 1) return crc32_block<1>(in, n, table, 0);
    // the 32-lane split done lane by lane (the device XORs the lanes' terms with shuffles)
    uint32_t acc = 0;
    const uint32_t slice = ((n + 32 * 4 - 1) / (32 * 4)) * 4;
    for (int lane = 0; lane < 32; lane++) {
        const uint32_t a = slice * lane < n ? slice * lane : n, b = a + slice < n ? a + slice : n;
        uint32_t c = 0xFFFFFFFFu;
        for (uint32_t i = a; i < b; i++) c = table[(c ^ in[i]) & 0xFFu] ^ (c >> 8);
        c = b > a ? ~c : 0u;
        acc ^= gf2_mulmod(x_pow_8n(n - b), c);
    }
    return acc;
}

// code lengths of build_code for a histogram: lens[nsym] out; -> coded bits, or -1
extern "C" long long oge_test_build_code(const unsigned *freq, int nsym, unsigned *lens, unsigned *codes) {
    using namespace oge_deflate;
    uint32_t tab[288], cnt[32];
    for (int s = 0; s < nsym; s++) tab[s] = freq[s];
    int used = 0;
    const uint64_t bits = build_code(tab, nsym, cnt, &used);
    if (bits == ~0ull) return -1;
    for (int s = 0; s < nsym; s++) { lens[s] = tab[s] >> 16; codes[s] = tab[s] & 0xFFFFu; }
    return (long long) bits;
}

extern "C" unsigned oge_test_deflate_work_bytes(void) { return (unsigned) sizeof(oge_deflate::Work); }

// ---- a lane-by-lane emulation of parse<32> (what the warp does in lockstep), for tuning the match finder on the CPU:
// -> the bytes deflate_block<32> would produce for this block (dynamic block; stored when not smaller).
// (Two refinements were measured with it and dropped, neither gained on BAM data: candidates inside the window, and letting
// the longest of the first few matches win instead of the first.)
extern "C" long long oge_test_emulate32(const unsigned char *in, unsigned n) {
    using namespace oge_deflate;
    Work *W = (Work *) calloc(1, sizeof(Work));
    unsigned char *src = (unsigned char *) aligned_alloc(16, ((size_t) n + 8 + 15) & ~(size_t) 15);
    memcpy(src, in, n);
    memset(src + n, 0xAB, 8);
    memset(W->htab, 0xFF, sizeof(W->htab));
    uint32_t pos = 0, xbits = 0;
    auto match_len = [&](uint32_t mp, uint32_t d) {
        uint32_t len = MIN_MATCH;
        while (mp + len < n && len < MAX_MATCH && src[mp + len] == src[mp + len - d]) len++;
        return len;
    };
    while (pos < n) {
        uint32_t w[32], cand[32], dist[32];
        bool valid[32];
        for (int l = 0; l < 32; l++) {
            const uint32_t p = pos + l;
            valid[l] = p + MIN_MATCH <= n;
            w[l] = valid[l] ? load4(src, p) : 0;
            cand[l] = valid[l] ? W->htab[(w[l] * 2654435761u) >> (32 - HBITS)] : 0xFFFFu;
        }
        auto insert_upto = [&](uint32_t end) {
            for (int l = 0; l < 32; l++)
                if (valid[l] && pos + l < end) W->htab[(w[l] * 2654435761u) >> (32 - HBITS)] = (uint16_t) (pos + l);
        };
        uint32_t found = 0;
        for (int l = 0; l < 32; l++) {
            const uint32_t p = pos + l;
            dist[l] = 0;
            if (!valid[l]) continue;
            if (cand[l] != 0xFFFFu && p - cand[l] <= MAX_DIST && load4(src, cand[l]) == w[l]) dist[l] = p - cand[l];
            else if (p >= 1 && load4(src, p - 1) == w[l]) dist[l] = 1;
            if (dist[l]) found |= 1u << l;
        }
        if (!found) {
            insert_upto(pos + 32);
            for (int l = 0; l < 32; l++)
                if (pos + l < n) W->lit[src[pos + l]]++;
            pos += 32;
            continue;
        }
        const int f = __builtin_ctz(found);
        const uint32_t len = match_len(pos + f, dist[f]);
        insert_upto(pos + f + len);
        for (int l = 0; l < f; l++) W->lit[src[pos + l]]++;
        uint32_t eb, ev;
        W->lit[len_symbol(len, &eb, &ev)]++;
        xbits += eb;
        W->dst[dist_symbol(dist[f], &eb, &ev)]++;
        xbits += eb;
        pos = pos + f + len;
    }
    W->lit[256]++;
    uint32_t nlit = 257, ndist = 2;
    for (int s = 257; s < 286; s++) if (W->lit[s]) nlit = s + 1;
    int used;
    const uint64_t b1 = build_code(W->lit, 286, W->cnt, &used), b2 = build_code(W->dst, 30, W->cnt, &used);
    for (int s = 0; s < 30; s++) if (W->dst[s] && (uint32_t) s + 1 > ndist) ndist = s + 1;
    uint64_t bytes = (17 + 57 + 4ull * (nlit + ndist) + b1 + b2 + xbits + 7) / 8;
    if (bytes >= (uint64_t) n + 5) bytes = n + 5;
    free(W); free(src);
    return (long long) bytes;
}
