// Reduced fragment pass.  The fragment half of generateDuplicateIndexes (reference
// algorithms/mark_duplicates.cpp:371-390) only ever acts on a run of equal keys that holds at
// least one end that is NOT part of a pair (:379 containsFrags), and only marks such ends
// (markDuplicateFragments, :515-540): ends of pairs merely testify "a pair starts here" (:517).
// So a fragment entry matters iff it is unpaired, or it is paired and shares its key with an
// unpaired one.  On paired-end data that is a tiny subset -- often empty -- and sorting only that
// subset instead of every record's entry gives the same marks:
//   ff_collect    unpaired entries -> list U
//   ff_set_build  their keys -> open-addressing set
//   ff_filter     paired entries whose key is in the set -> appended to U
// then K3 + K4 run on U.  The caller falls back to sorting everything when unpaired ends are common.
#include "kernels.cuh"

namespace oge {

constexpr int FF_THREADS = 256;

__device__ __forceinline__ uint32_t ff_warp_append(bool want, uint32_t *counter) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, want);
    if (!m) return 0;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t) __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

__device__ __forceinline__ E128 ff_ld(const E128 *p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    E128 e;
    e.lo = v.x;
    e.hi = v.y;
    return e;
}
__device__ __forceinline__ bool ff_dead(const E128 &e) { return (e.lo & e.hi) == ~0ull; }
// the duplicate key of a fragment entry as one word (the caller guarantees it is at most 63 bits wide)
__device__ __forceinline__ uint64_t ff_key(const KeyLayout &L, const E128 &e) { return bits_get(e, L.f_orient, L.f_end - L.f_orient); }
__device__ __forceinline__ uint64_t ff_slot(uint64_t key, uint64_t mask) {
    uint64_t x = key * 0x9E3779B97F4A7C15ull;
    return (x ^ (x >> 29)) & mask;
}

__global__ void __launch_bounds__(FF_THREADS) ff_collect_kernel(const E128 *__restrict__ frag, uint64_t n, KeyLayout L, E128 *__restrict__ out,
                                                                uint32_t cap, uint32_t *__restrict__ counters) {
    const uint64_t i = (uint64_t) blockIdx.x * FF_THREADS + threadIdx.x;
    bool want = false;
    E128 e;
    e.lo = e.hi = 0;
    if (i < n) {
        e = ff_ld(frag + i);
        want = !ff_dead(e) && bits_get(e, L.f_paired, 1) == 0;
    }
    const uint32_t at = ff_warp_append(want, &counters[CNT_UFRAG]);
    if (want && at < cap) reinterpret_cast<ulonglong2 *>(out)[at] = make_ulonglong2(e.lo, e.hi);
}

// set slots hold key + 1 (0 = empty); n_u entries of `list` are inserted
__global__ void __launch_bounds__(FF_THREADS) ff_set_build_kernel(const E128 *__restrict__ list, const uint32_t *__restrict__ n_dev, uint32_t n_max,
                                                                  KeyLayout L, unsigned long long *__restrict__ set, uint64_t mask) {
    const uint32_t j = blockIdx.x * FF_THREADS + threadIdx.x;
    const uint32_t n_u = min(*n_dev, n_max);
    if (j >= n_u) return;
    const E128 e = ff_ld(list + j);
    const unsigned long long k = ff_key(L, e) + 1;
    uint64_t s = ff_slot(k, mask);
    while (true) {
        unsigned long long old = set[s];
        if (old == k) return;
        if (old == 0) {
            old = atomicCAS(&set[s], 0ull, k);
            if (old == 0 || old == k) return;
        }
        s = (s + 1) & mask;
    }
}

__global__ void __launch_bounds__(FF_THREADS) ff_filter_kernel(const E128 *__restrict__ frag, uint64_t n, KeyLayout L,
                                                               const unsigned long long *__restrict__ set, uint64_t mask, E128 *__restrict__ out,
                                                               uint32_t cap, uint32_t *__restrict__ counters, const uint32_t *__restrict__ n_set) {
    if (n_set && *n_set == 0) return;      // no unpaired end anywhere: nothing can share a key with one
    const uint64_t i = (uint64_t) blockIdx.x * FF_THREADS + threadIdx.x;
    bool want = false;
    E128 e;
    e.lo = e.hi = 0;
    if (i < n) {
        e = ff_ld(frag + i);
        if (!ff_dead(e) && bits_get(e, L.f_paired, 1) != 0) {
            const unsigned long long k = ff_key(L, e) + 1;
            uint64_t s = ff_slot(k, mask);
            while (true) {
                const unsigned long long v = __ldg(&set[s]);
                if (v == k) { want = true; break; }
                if (v == 0) break;
                s = (s + 1) & mask;
            }
        }
    }
    const uint32_t at = ff_warp_append(want, &counters[CNT_UFRAG]);
    if (want && at < cap) reinterpret_cast<ulonglong2 *>(out)[at] = make_ulonglong2(e.lo, e.hi);
}

static inline uint32_t ff_grid(uint64_t n) { return (uint32_t) ((n + FF_THREADS - 1) / FF_THREADS); }

int launch_ff_collect(const E128 *frag, uint64_t n, const KeyLayout &L, E128 *out, uint32_t cap, uint32_t *counters, cudaStream_t s,
                      uint64_t *launches) {
    if (!n) return 0;
    ff_collect_kernel<<<ff_grid(n), FF_THREADS, 0, s>>>(frag, n, L, out, cap, counters);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_ff_set_build(const E128 *list, const uint32_t *n_dev, uint32_t n_max, const KeyLayout &L, unsigned long long *set, uint64_t n_slots,
                        cudaStream_t s, uint64_t *launches) {
    if (!n_max) return 0;
    ff_set_build_kernel<<<ff_grid(n_max), FF_THREADS, 0, s>>>(list, n_dev, n_max, L, set, n_slots - 1);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_ff_filter(const E128 *frag, uint64_t n, const KeyLayout &L, const unsigned long long *set, uint64_t n_slots, E128 *out, uint32_t cap,
                     uint32_t *counters, const uint32_t *n_set, cudaStream_t s, uint64_t *launches) {
    if (!n) return 0;
    ff_filter_kernel<<<ff_grid(n), FF_THREADS, 0, s>>>(frag, n, L, set, n_slots - 1, out, cap, counters, n_set);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
