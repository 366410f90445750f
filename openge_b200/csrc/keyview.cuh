// Exact view of a record's pairing key, RG value + ":" + read name (reference algorithms/mark_duplicates.cpp:210-214),
// read from the record in global memory.  Shared by the exact path of the mate join (join.cu) and the range-sharded
// phases (shard.cu).
#pragma once
#include "kernels.cuh"

namespace oge {

// RG value location by the same tag walk as endbuild.cu (FindTag + SkipToNextTag, BamAlignment.cpp:270-294, 699-786)
static __device__ int find_rg_global(const uint8_t *tags, uint32_t n, uint32_t *len);   // defined below

struct KeyView {
    const uint8_t *rg, *name;
    uint32_t rg_len, name_len;
};

static __device__ KeyView key_view(const uint8_t *rec, const uint64_t *off, uint64_t i) {
    const uint8_t *p = rec + off[i];
    uint32_t rec_len = (uint32_t) (off[i + 1] - off[i]);
    uint32_t l_name = p[12];
    uint32_t n_cig = (uint32_t) p[16] | ((uint32_t) p[17] << 8);
    uint32_t l_seq = (uint32_t) p[20] | ((uint32_t) p[21] << 8) | ((uint32_t) p[22] << 16) | ((uint32_t) p[23] << 24);
    uint32_t o_tags = 36 + l_name + 4 * n_cig + ((l_seq + 1) >> 1) + l_seq;
    KeyView v;
    v.name = p + 36;
    v.name_len = l_name ? l_name - 1 : 0;
    uint32_t rl;
    int at = find_rg_global(p + o_tags, rec_len - o_tags, &rl);
    v.rg = at >= 0 ? p + o_tags + at : p;
    v.rg_len = at >= 0 ? rl : 0;
    return v;
}

__device__ __forceinline__ uint8_t key_byte(const KeyView &v, uint32_t i) {
    return i < v.rg_len ? v.rg[i] : (i == v.rg_len ? (uint8_t) ':' : v.name[i - v.rg_len - 1]);
}

static __device__ bool key_equal(const KeyView &a, const KeyView &b) {
    uint32_t la = a.rg_len + 1 + a.name_len, lb = b.rg_len + 1 + b.name_len;
    if (la != lb) return false;
    for (uint32_t i = 0; i < la; i++)
        if (key_byte(a, i) != key_byte(b, i)) return false;
    return true;
}

static __device__ int find_rg_global(const uint8_t *tags, uint32_t n, uint32_t *len) {
    uint32_t parsed = 0;
    *len = 0;
    while (parsed < n) {
        if (n - parsed < 3) return -1;
        uint8_t t0 = tags[parsed], t1 = tags[parsed + 1], type = tags[parsed + 2];
        parsed += 3;
        if (t0 == 'R' && t1 == 'G') {
            uint32_t l = 0;
            while (parsed + l < n && tags[parsed + l]) l++;
            *len = l;
            return (int) parsed;
        }
        switch (type) {
            case 'A': case 'c': case 'C': parsed += 1; break;
            case 's': case 'S': parsed += 2; break;
            case 'f': case 'i': case 'I': parsed += 4; break;
            case 'Z': case 'H':
                while (parsed < n && tags[parsed]) parsed++;
                parsed++;
                break;
            case 'B': {
                if (parsed + 5 > n) return -1;
                uint8_t at = tags[parsed];
                int32_t cnt = (int32_t) ((uint32_t) tags[parsed + 1] | ((uint32_t) tags[parsed + 2] << 8) |
                                         ((uint32_t) tags[parsed + 3] << 16) | ((uint32_t) tags[parsed + 4] << 24));
                parsed += 5;
                long long skip;
                if (at == 'c' || at == 'C') skip = cnt;
                else if (at == 's' || at == 'S') skip = 2ll * cnt;
                else if (at == 'f' || at == 'i' || at == 'I') skip = 4ll * cnt;
                else return -1;
                if (skip < 0 || (long long) parsed + skip > (long long) n) return -1;
                parsed += (uint32_t) skip;
                break;
            }
            default: return -1;
        }
        if (parsed >= n) return -1;
        if (tags[parsed] == 0) return -1;
    }
    return -1;
}


}  // namespace oge
