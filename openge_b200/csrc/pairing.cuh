// Device helpers shared by the mate join (join.cu) and the range-sharded phases (shard.cu).
#pragma once
#include "kernels.cuh"

namespace oge {

constexpr uint32_t SLOT_NO_PAIR = 0xFFFFFFFFu;
constexpr uint32_t SLOT_PAIR_FAR = 0x80000000u;      // pair_pos flag: the entry sits in the far-pair list

__device__ __forceinline__ uint64_t slot_of(uint64_t h, uint64_t n_slots) { return __umul64hi(h, n_slots); }

// ---- pair entry ---------------------------------------------------------------------------------
// first = the record seen first in the file (the map's stored ReadEnds), second = the current one.
// Generic form: field by field, any widths.
static __device__ __noinline__ E128 make_pair_entry_generic(const KeyLayout &L, const E128 &first, const E128 &second,
                                                     uint32_t *idx1_local, uint32_t *idx2_local, uint64_t idx_base, bool *far) {
    uint64_t lib = bits_get(first, L.f_lib, L.lib_bits);      // library of the first-seen end (:218)
    uint64_t ref_f = bits_get(first, L.f_ref, L.ref_bits), ref_s = bits_get(second, L.f_ref, L.ref_bits);
    uint64_t co_f = bits_get(first, L.f_coord, L.coord_bits), co_s = bits_get(second, L.f_coord, L.coord_bits);
    uint64_t rev_f = bits_get(first, L.f_orient, 1), rev_s = bits_get(second, L.f_orient, 1);
    uint64_t idx_f = bits_get(first, L.f_idx, L.idx_bits), idx_s = bits_get(second, L.f_idx, L.idx_bits);
    uint32_t score = ((uint32_t) first.lo + (uint32_t) second.lo) & 0xFFFFu;      // short + short (:245)
    E128 e;
    e.lo = score;
    e.hi = 0;
    // second >= first in (sequence, coordinate): keep order, else flip (:229-243)
    bool keep = ref_s > ref_f || (ref_s == ref_f && co_s >= co_f);
    uint64_t r1 = keep ? ref_f : ref_s, c1 = keep ? co_f : co_s, v1 = keep ? rev_f : rev_s, i1 = keep ? idx_f : idx_s;
    uint64_t r2 = keep ? ref_s : ref_f, c2 = keep ? co_s : co_f, v2 = keep ? rev_s : rev_f, i2 = keep ? idx_s : idx_f;
    bits_or(e, L.p_idx, i1);
    // same reference and a short distance (c2 >= c1 then, by the flip rule): the near form
    *far = !(r1 == r2 && c2 - c1 < (1ull << L.delta_bits));
    if (*far) {
        bits_or(e, L.p_coord2, c2);
        bits_or(e, L.p_ref2, r2);
    } else {
        bits_or(e, L.n_delta, c2 - c1);
    }
    bits_or(e, L.p_orient, (v1 << 1) | v2);      // getOrientationByte(read1Negative, read2Negative) (:169-178)
    bits_or(e, L.p_coord1, c1);
    bits_or(e, L.p_ref1, r1);
    bits_or(e, L.p_lib, lib);
    *idx1_local = (uint32_t) (i1 - idx_base);
    *idx2_local = (uint32_t) (i2 - idx_base);
    return e;
}

// Word form, for layouts with L.fast (every human-sized file): in a fragment entry [coord][ref][lib] is one
// contiguous block above bit f_coord (0 < f_coord < 64), and the pair key holds the same block [coord1][ref1][lib]
// in the same order, so it moves as a whole; (ref, coord) compares as the integer [coord][ref]; and the near key
// (block + orientation + distance) is at most 64 bits wide and, anchored at bit 127, lies in the high word.
__device__ __forceinline__ E128 make_pair_entry(const KeyLayout &L, const E128 &first, const E128 &second,
                                                uint32_t *idx1_local, uint32_t *idx2_local, uint64_t idx_base, bool *far) {
    if (!L.fast) return make_pair_entry_generic(L, first, second, idx1_local, idx2_local, idx_base, far);
    const int fc = L.f_coord, cr = L.coord_bits + L.ref_bits;
    const uint64_t blk_f = (first.lo >> fc) | (first.hi << (64 - fc)), blk_s = (second.lo >> fc) | (second.hi << (64 - fc));
    const uint64_t cr_mask = (1ull << cr) - 1;
    const uint64_t pos_f = blk_f & cr_mask, pos_s = blk_s & cr_mask;      // [coord][ref] as one number
    const bool keep = pos_s >= pos_f;                                      // (:229-243)
    const uint64_t p1 = keep ? pos_f : pos_s, p2 = keep ? pos_s : pos_f;
    const uint64_t lo1 = keep ? first.lo : second.lo, lo2 = keep ? second.lo : first.lo;
    const uint64_t idx_mask = (1ull << L.idx_bits) - 1;
    const uint64_t i1 = (lo1 >> 16) & idx_mask, i2 = (lo2 >> 16) & idx_mask;
    const uint64_t orient = (((lo1 >> L.f_orient) & 1) << 1) | ((lo2 >> L.f_orient) & 1);      // (:169-178)
    const uint64_t lib = blk_f >> cr;                                      // library of the first-seen end (:218); bits above are zero
    const uint64_t block1 = (lib << cr) | p1;
    const uint64_t delta = p2 - p1;
    *far = ((p1 ^ p2) >> L.coord_bits) != 0 || delta >= (1ull << L.delta_bits);
    E128 e;
    e.lo = (((uint32_t) first.lo + (uint32_t) second.lo) & 0xFFFFu) | (i1 << 16);      // short + short (:245)
    e.hi = 0;
    if (!*far) {
        const uint64_t key = (((block1 << 2) | orient) << L.delta_bits) | delta;
        e.hi = key << (L.n_delta - 64);
    } else {
        bits_or(e, L.p_coord2, p2);      // [coord2][ref2], contiguous like the block
        bits_or(e, L.p_orient, orient);
        bits_or(e, L.p_coord1, block1);
    }
    *idx1_local = (uint32_t) (i1 - idx_base);
    *idx2_local = (uint32_t) (i2 - idx_base);
    return e;
}

__device__ __forceinline__ E128 ld_frag(const E128 *p) {
    ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    E128 e;
    e.lo = v.x;
    e.hi = v.y;
    return e;
}

// unaligned little-endian 32-bit read from global memory: two aligned words + funnel shift
__device__ __forceinline__ uint32_t ldg_u32_unaligned(const uint8_t *p) {
    const uintptr_t a = (uintptr_t) p, wa = a & ~(uintptr_t) 3;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(wa);
    return __funnelshift_r(w[0], w[1], (uint32_t) (a & 3) * 8);
}

// n bytes at a and at b equal?  (global memory; reads past the n bytes stay inside the record buffer and are masked)
__device__ __forceinline__ bool bytes_equal_global(const uint8_t *a, const uint8_t *b, uint32_t n) {
    uint32_t j = 0;
    for (; j + 4 <= n; j += 4)
        if (ldg_u32_unaligned(a + j) != ldg_u32_unaligned(b + j)) return false;
    if (j < n) {
        const uint32_t m = (1u << (8 * (n - j))) - 1;
        if ((ldg_u32_unaligned(a + j) ^ ldg_u32_unaligned(b + j)) & m) return false;
    }
    return true;
}

// the pairing-key tag of a record (kernels.cuh: NameTag), rebuilt from the record in global memory
__device__ __forceinline__ void make_tag_global(const uint8_t *rec, uint32_t rgc, uint32_t l_name, NameTag *out) {
    const uint32_t nlen = l_name ? l_name - 1 : 0;
    uint32_t w[8];
    w[0] = rgc | (l_name << 16) | (nlen ? ((uint32_t) rec[36] << 24) : 0u);
#pragma unroll
    for (int k = 1; k < 8; k++) {
        const uint32_t first = 4 * k - 3;      // name bytes [first, first + 4)
        uint32_t v = 0;
        if (first < nlen) {
            v = ldg_u32_unaligned(rec + 36 + first);
            const uint32_t have = nlen - first;
            if (have < 4) v &= (1u << (8 * have)) - 1;
        }
        w[k] = v;
    }
    uint4 *t = reinterpret_cast<uint4 *>(out);
    t[0] = make_uint4(w[0], w[1], w[2], w[3]);
    t[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

__device__ __forceinline__ E128 complex_entry(uint64_t h, uint32_t ordinal) {
    E128 e;
    e.lo = (h << 32) | ordinal;
    e.hi = h >> 32;
    return e;
}


}  // namespace oge
