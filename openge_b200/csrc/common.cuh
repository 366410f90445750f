// Shared definitions of the B200 duplicate-marking path (libopenge_b200.so).
//
// Data layout in HBM:
//   records   u8[]   raw BAM records back to back (reference layout: util/bam_deserializer.h:144-193)
//   offsets   u64[n+1]
//   frag      E128[n]  one 16-byte end entry per record (the ReadEnds of picard_structures.h:29-54,
//                      packed), written by the end-build kernel and sorted in place of fragSort
//   pair      E128[]   one 16-byte entry per matched pair (pairSort)
//
// A 16-byte entry is a little-endian 128-bit integer: payload in the low bits, the duplicate
// key in the high bits, so that the LSD radix sort only has to walk the key's bit range.
//
//   frag:  [score:16][idx:idx_bits][paired:1] | key: [orient:1][coord][ref][lib]
//   pair:  [score:16][idx1:idx_bits] ...0...  | key: [coord2][ref2][orient:2][coord1][ref1][lib]        "far"
//          [score:16][idx1:idx_bits] ...0...  | key:        [delta][orient:2][coord1][ref1][lib]        "near"
// Pair keys are anchored at bit 127, so lib/ref1/coord1/orient sit at the same place in both forms.
// A pair whose ends lie on the same reference less than 2^delta_bits apart (every ordinary insert)
// is a NEAR pair: its second end is stored as the distance, which makes the key 3 radix passes
// shorter.  Near and far pairs live in separate lists; equal duplicate keys imply equal class.
//
// The key order is irrelevant (only equality defines a duplicate group, and the survivor is
// chosen by an order-independent (score desc, index asc) reduction; picard_structures.h:56-68
// + mark_duplicates.cpp:494,528), which is why the fields can be packed in any order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oge {

struct __align__(16) E128 {
    uint64_t lo, hi;
};

struct KeyLayout {
    int idx_bits, coord_bits, ref_bits, lib_bits;
    long long coord_bias;    // stored coordinate = coord + coord_bias, must land in [0, 2^coord_bits)
    uint32_t lib_invalid;    // lib field value of "no entry" (all ones)
    // frag entry bit positions
    int f_idx, f_paired, f_orient, f_coord, f_ref, f_lib, f_end;   // f_orient = first key bit
    // pair entry bit positions (keys anchored at the top: p_end == 128)
    int p_idx, p_coord2, p_ref2, p_orient, p_coord1, p_ref1, p_lib, p_end;   // far pairs: p_coord2 = first key bit
    int n_delta, delta_bits;                                                 // near pairs: n_delta = first key bit
    int fast;      // 1: field widths allow the 64-bit word form of the pair-entry builder (pairing.cuh); set by compute_layout
};

// ---- 128-bit field helpers (positions and widths are warp-uniform) ---------------------------
__host__ __device__ __forceinline__ uint64_t bits_get(const E128 &e, int pos, int width) {
    uint64_t v;
    if (pos >= 64) v = e.hi >> (pos - 64);
    else if (pos == 0) v = e.lo;
    else v = (e.lo >> pos) | (e.hi << (64 - pos));
    return width >= 64 ? v : (v & ((1ull << width) - 1));
}

__host__ __device__ __forceinline__ void bits_or(E128 &e, int pos, uint64_t v) {   // v already masked
    if (pos >= 64) e.hi |= v << (pos - 64);
    else {
        e.lo |= v << pos;
        if (pos) e.hi |= v >> (64 - pos);
    }
}

// e >> pos, as a 128-bit value: the duplicate key (everything at and above `pos`).
__host__ __device__ __forceinline__ E128 bits_from(const E128 &e, int pos) {
    E128 r;
    if (pos >= 64) { r.lo = e.hi >> (pos - 64); r.hi = 0; }
    else if (pos == 0) r = e;
    else { r.lo = (e.lo >> pos) | (e.hi << (64 - pos)); r.hi = e.hi >> pos; }
    return r;
}

__host__ __device__ __forceinline__ uint32_t digit_of(const E128 &e, int shift, uint32_t mask) {
    uint64_t v;
    if (shift >= 64) v = e.hi >> (shift - 64);
    else if (shift == 0) v = e.lo;
    else v = (e.lo >> shift) | (e.hi << (64 - shift));
    return (uint32_t) v & mask;
}

// ---- device error bits (OR-ed into ctx counters) ---------------------------------------------
enum : uint32_t {
    DEV_ERR_KEY_RANGE = 1u,      // refID / coordinate / library outside the configured key layout
    DEV_ERR_BAD_RECORD = 2u,     // record sections overrun block_size, or offsets disagree with block_size
    DEV_ERR_CAPACITY = 4u,       // an output list overran its capacity (internal sizing error)
};

// ---- counters block (one u32 array per context) ----------------------------------------------
enum {
    CNT_ERR = 0,
    CNT_FRAG,           // eligible records (fragSort.size())
    CNT_UNPAIRED,       // eligible records that are not an end of a pair (the only ones the fragment pass can mark)
    CNT_UFRAG,          // fragment entries collected for the reduced fragment sort
    CNT_PAIR_ELIGIBLE,  // records that enter the mate map
    CNT_PAIRS,          // pair entries emitted (pairSort.size())
    CNT_COMPLEX,        // half-pair entries sent to the exact slow path
    CNT_HASH_MISMATCH,  // hash-equal couples rejected by the byte comparison
    CNT_MARKS,          // addIndexAsDuplicate calls (mark_duplicates.cpp:477-480)
    CNT_DUPS,           // records with 0x400 after the flag pass
    CNT_KEPT,           // records kept by pull (remove_duplicates)
    CNT_COMPLEX_SEGS,
    CNT_COMPLEX_SLOTS,  // slots that saw a third arrival
    CNT_PAIRS_RETRACTED,
    CNT_PAIRS_FAR,      // far-pair entries emitted
    CNT_FAR_RETRACTED,
    CNT_FOREIGN_MARKS,  // sharded: marks on records of other ranks
    CNT_PUB,            // sharded: published entries
    CNT_ROUTE,          // sharded: entries handed to their key owner
    CNT_FRAG_EXTRA,     // sharded: fragment entries received
    CNT_FM,             // sharded: foreign-mate couples
    CNT_FRAG_VALID,     // sharded: fragment entries to select over (device-side count)
    CNT_FOREIGN_MARKS_FRAG,
    CNT_SCRATCH0,
    CNT_SCRATCH1,
    CNT_SCRATCH2,
    CNT_LEFT,           // fused end-build: records handed to the global join
    CNT_N = 32
};

// ---- launch bookkeeping ----------------------------------------------------------------------
struct LaunchCount {
    uint64_t n = 0;
};

#define OGE_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) return oge::fail_cuda(_e, #expr, __FILE__, __LINE__);   \
    } while (0)

int fail_cuda(cudaError_t e, const char *what, const char *file, int line);
int fail_msg(int code, const char *fmt, ...);

}  // namespace oge
