// LSD onesweep radix sort of 16-byte entries on a bit range (replaces ogeSortMt,
// reference util/thread_pool.h:359-397, called at algorithms/mark_duplicates.cpp:262-271).
#pragma once
#include "common.cuh"

namespace oge {

constexpr int RS_MAX_PASSES = 16;       // 128 key bits / 8
constexpr int RS_RADIX_BITS = 8;
constexpr int RS_RADIX = 1 << RS_RADIX_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;      // 2048 entries = 32 KB of shared memory
constexpr int RS_WARPS = RS_THREADS / 32;

struct SortPlan {
    int n_pass;
    int shift[RS_MAX_PASSES];
    int bits[RS_MAX_PASSES];
};

SortPlan make_sort_plan(int bit_lo, int bit_hi);

// Scratch a sort of up to `n` entries needs (bytes), excluding the ping-pong buffer.
size_t sort_scratch_bytes(uint64_t n);

// Sorts `n` entries by bits [bit_lo, bit_hi).  `a` holds the input; `b` is a same-sized
// ping-pong buffer.  Returns the buffer holding the result through *result (a or b).
// `n_dev`, when not null, is a device word holding the real entry count (<= n); the kernels
// then ignore the tail, so no host round trip is needed to size the sort.
// Optional per-launch timing of the pass kernel: events are taken from `pool` (pairs), `used` counts pairs.
struct PassTimer {
    cudaEvent_t *pool;
    int cap, used;
    uint64_t bytes;      // algorithmic bytes of the timed launches
};

int radix_sort_128(E128 *a, E128 *b, uint64_t n, const uint32_t *n_dev, int bit_lo, int bit_hi, void *scratch,
                   cudaStream_t stream, E128 **result, uint64_t *launches, PassTimer *timer = nullptr);

int radix_sort_init();   // one-time function attributes

// Pass-kernel variant (tuning / A-B measurement): 0 = 256-thread CTAs (2048-entry tiles), 2 = 512-thread CTAs
// (4096-entry tiles, the default); bits 4-6 are measurement knobs (radix_sort.cu), bits 8.. the L2 prefetch distance.
void radix_sort_set_variant(int v);
int radix_sort_get_variant();
void radix_sort_set_prefetch(int tiles);

}  // namespace oge
