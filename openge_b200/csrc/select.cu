// K4: group-and-select.  Replaces generateDuplicateIndexes, markDuplicatePairs and
// markDuplicateFragments (reference algorithms/mark_duplicates.cpp:326-400, 488-507, 515-540)
// over the radix-sorted entries.
//
// A duplicate group is a run of equal keys (areComparableForDuplicates, :402-414).  The
// reference keeps the first entry, in ReadEnds::compare order (util/picard_structures.h:56-68),
// holding the strictly greatest score (:494, :528); since the order inside a group continues
// with read1IndexInFile and a record is read 1 of at most one pair, the survivor is
//     argmax over the run of (score as int16, then smallest index)
// which is order independent -- so the sort need not reproduce the comparator's order.
//   pairs: every non-survivor of a run of 2+ marks both of its records (:499-505)
//   frags: runs of 2+ holding at least one unpaired end (:379); if the run also holds an end
//          of a pair, all unpaired ends are marked (:517-522), else all but the survivor (:524-538)
//
// One CTA per 1024-entry tile, four consecutive entries per thread (eight were measured: slower, 80 registers).  The survivor of a run is the
// maximum of one packed 64-bit candidate per entry, (score + 2^15) << 32 | ~index, so "greatest score,
// then smallest index" is a single max.  Runs that lie inside one thread's eight entries (almost all of
// them: most runs are singletons) are reduced in registers with a forward and a backward sweep; only
// the partial runs at a thread's edges go through shared-memory atomics, into the slot of the thread
// the run starts in (found with a block-wide max-scan of "last thread holding a run head").
// A run that starts in a tile is owned by that tile even when it spills into the next ones: the
// owner keeps reading until the key changes; a tile's leading entries that continue an earlier
// run are skipped.  HBM traffic: the sorted entries once (+ the spill) and one byte per mark.
#include "kernels.cuh"

namespace oge {

constexpr int SEL_THREADS = 256;
#ifndef OGE_SEL_ITEMS
#define OGE_SEL_ITEMS 4
#endif
constexpr int SEL_ITEMS = OGE_SEL_ITEMS;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;      // 2048
constexpr int SEL_PAD = SEL_TILE + SEL_TILE / 8 + 2;   // one skew slot per 8 entries: conflict-free blocked reads
constexpr int SEL_WARPS = SEL_THREADS / 32;

constexpr uint32_t RUN_HAS_PAIRED = 1u, RUN_HAS_UNPAIRED = 2u;

constexpr size_t SEL_SMEM = (size_t) SEL_PAD * sizeof(E128);

__device__ __forceinline__ int pad_index(int j) { return j + (j >> 3); }

__device__ __forceinline__ E128 ldg_entry(const E128 *p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    E128 e;
    e.lo = v.x;
    e.hi = v.y;
    return e;
}

// equal on every bit at and above key_lo?
__device__ __forceinline__ bool key_eq(const E128 &a, const E128 &b, int key_lo) {
    const uint64_t xh = a.hi ^ b.hi, xl = a.lo ^ b.lo;
    if (key_lo >= 64) return (xh >> (key_lo - 64)) == 0;
    return xh == 0 && (xl >> key_lo) == 0;
}

// the rare path (an entry that is a duplicate): kept out of line so the common path stays small
__device__ __noinline__ void mark_record(uint8_t *dup, uint64_t idx_base, uint64_t n_records, uint32_t *foreign_counter,
                                         uint32_t *foreign_marks, uint32_t foreign_cap, uint32_t g) {
    const uint64_t l = (uint64_t) g - idx_base;      // wraps for ordinals below the base
    if (l < n_records) dup[l] = 1;
    else {
        uint32_t at = atomicAdd(foreign_counter, 1u);
        if (at < foreign_cap) foreign_marks[at] = g;
    }
}
__device__ __noinline__ uint32_t foreign_mate(const uint64_t *fm, uint32_t n_fm, uint32_t g1) {
    uint32_t lo = 0, hi = n_fm;      // first couple with idx1 >= g1
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t) (fm[mid] >> 32) < g1) lo = mid + 1;
        else hi = mid;
    }
    return (uint32_t) fm[lo];
}

// what a run needs to know about its entries
struct RunAgg {
    unsigned long long cand;      // max of (score + 2^15) << 32 | ~index
    uint32_t cnt, fl;
};
__device__ __forceinline__ void agg_add(RunAgg &a, unsigned long long cand, uint32_t fl) {
    a.cand = cand > a.cand ? cand : a.cand;
    a.cnt += 1;
    a.fl |= fl;
}

// KIND: 0 fragment ends, 1 near pairs, 2 far pairs (they differ in where the key starts)
template <int KIND>
__global__ void __launch_bounds__(SEL_THREADS, SEL_ITEMS == 8 ? 3 : 5) select_kernel(SelectParams P) {
    constexpr bool PAIRS = KIND != 0;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    E128 *s_e = reinterpret_cast<E128 *>(smem_raw);
    __shared__ unsigned long long s_cand[SEL_THREADS];
    __shared__ uint32_t s_cnt[SEL_THREADS], s_fl[SEL_THREADS];
    __shared__ int s_wmax[SEL_WARPS];
    __shared__ uint32_t s_ext_end;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = P.n_dev ? min(*P.n_dev, P.n_max) : P.n_max;
    const uint32_t base = blockIdx.x * SEL_TILE;
    if (base >= n) return;
    const uint32_t count = min((uint32_t) SEL_TILE, n - base);
    const KeyLayout &L = P.kl;
    const int key_lo = KIND == 0 ? L.f_orient : (KIND == 1 ? L.n_delta : L.p_coord2);
    const int idx_pos = PAIRS ? L.p_idx : L.f_idx;

    // ---- load the tile (striped, coalesced) + its predecessor into shared memory
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; k++) {
        int j = k * SEL_THREADS + tid;
        if ((uint32_t) j < count) s_e[pad_index(j) + 1] = ldg_entry(P.sorted + base + j);
    }
    s_cand[tid] = 0;
    s_cnt[tid] = 0;
    s_fl[tid] = 0;
    if (tid == 0) {
        if (base > 0) s_e[0] = ldg_entry(P.sorted + base - 1);
        s_ext_end = 0xFFFFFFFFu;
    }
    __syncthreads();

    // global ordinals (< 2^32); [idx_base, idx_base + n_records) are this rank's records
    auto idx_of = [&](const E128 &v) { return (uint32_t) bits_get(v, idx_pos, L.idx_bits); };
    auto cand_of = [&](const E128 &v) {
        const uint32_t sc = (uint32_t) ((int) (int16_t) (uint16_t) (v.lo & 0xFFFFu) + 32768);
        return ((unsigned long long) sc << 32) | (uint32_t) ~idx_of(v);
    };
    auto flags_of = [&](const E128 &v) -> uint32_t {
        if (PAIRS) return 0;
        return bits_get(v, L.f_paired, 1) ? RUN_HAS_PAIRED : RUN_HAS_UNPAIRED;
    };
    uint32_t marks = 0;
    // the verdict on one entry given its run's totals (mark_duplicates.cpp:488-507, 515-540)
    auto decide = [&](const E128 &v, unsigned long long cand, const RunAgg &r) {
        if (r.cnt < 2) return;
        const bool is_best = cand == r.cand;
        if (PAIRS) {
            if (!is_best) {
                const uint32_t g1 = idx_of(v);
                const uint64_t l1 = (uint64_t) g1 - P.idx_base;
                const uint32_t g2 = l1 < P.n_records ? P.mate_of[l1] : foreign_mate(P.fm, P.n_fm, g1);
                mark_record(P.dup, P.idx_base, P.n_records, P.foreign_counter, P.foreign_marks, P.foreign_cap, g1);
                mark_record(P.dup, P.idx_base, P.n_records, P.foreign_counter, P.foreign_marks, P.foreign_cap, g2);
                marks += 2;
            }
        } else {
            if (!(r.fl & RUN_HAS_UNPAIRED)) return;
            if (P.world > 1) {      // range shards: a run whose key another rank owns is judged there (its entries were copied over)
                const uint64_t packed = bits_get(v, L.f_coord, L.coord_bits + L.ref_bits);
                int owner = 0;
                for (int q = 0; q + 1 < P.world; q++) owner += P.split[q] <= packed ? 1 : 0;
                if (owner != P.rank) return;
            }
            const bool mark_it = (r.fl & RUN_HAS_PAIRED) ? !(bits_get(v, L.f_paired, 1) != 0) : !is_best;
            if (mark_it) {
                mark_record(P.dup, P.idx_base, P.n_records, P.foreign_counter, P.foreign_marks, P.foreign_cap, idx_of(v));
                marks += 1;
            }
        }
    };

    // ---- blocked view: thread t owns entries [8t, 8t+8); head flags
    E128 e[SEL_ITEMS];
    uint32_t heads = 0;      // bit k: entry k starts a run
    int n_mine = 0;
    {
        E128 prev;
        const int j0 = tid * SEL_ITEMS;
        prev = j0 == 0 ? s_e[0] : s_e[pad_index(j0 - 1) + 1];
#pragma unroll
        for (int k = 0; k < SEL_ITEMS; k++) {
            const int j = j0 + k;
            if ((uint32_t) j < count) {
                e[k] = s_e[pad_index(j) + 1];
                const bool head = (base + j == 0) || !key_eq(e[k], prev, key_lo);
                heads |= (head ? 1u : 0u) << k;
                prev = e[k];
                n_mine = k + 1;
            }
        }
    }
    const int first_head = heads ? __ffs(heads) - 1 : n_mine;      // entries before it continue a run from the left
    const int last_head = heads ? 31 - __clz(heads) : -1;         // entries from it on are this thread's trailing run

    // ---- slot of the run my leading entries continue: the nearest thread to the left holding a head
    int left;
    {
        int v = heads ? tid : -1, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x = max(x, y);
        }
        if (lane == 31) s_wmax[warp] = x;
        int prev_lane = __shfl_up_sync(0xFFFFFFFFu, x, 1);
        __syncthreads();
        int wprev = -1;
        for (int w = 0; w < warp; w++) wprev = max(wprev, s_wmax[w]);
        left = lane == 0 ? wprev : max(wprev, prev_lane);
    }
    int tile_last_slot = -1;      // slot of the tile's last run (-1: the whole tile continues an earlier tile's run)
    for (int w = 0; w < SEL_WARPS; w++) tile_last_slot = max(tile_last_slot, s_wmax[w]);

    // ---- forward sweep: running totals inside each run; edge partials go to shared memory
    RunAgg F[SEL_ITEMS];
    {
        RunAgg cur = {0ull, 0u, 0u};
#pragma unroll
        for (int k = 0; k < SEL_ITEMS; k++) {
            if (k < n_mine) {
                if ((heads >> k) & 1) cur = RunAgg{0ull, 0u, 0u};
                agg_add(cur, cand_of(e[k]), flags_of(e[k]));
                F[k] = cur;
            }
        }
    }
    // leading partial = F[first_head - 1], trailing partial = F[n_mine - 1] (when the thread holds a head)
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; k++) {
        if (k == first_head - 1 && left >= 0) {
            atomicMax(&s_cand[left], F[k].cand);
            atomicAdd(&s_cnt[left], F[k].cnt);
            if (!PAIRS) atomicOr(&s_fl[left], F[k].fl);
        }
        if (k == n_mine - 1 && last_head >= 0) {
            atomicMax(&s_cand[tid], F[k].cand);
            atomicAdd(&s_cnt[tid], F[k].cnt);
            if (!PAIRS) atomicOr(&s_fl[tid], F[k].fl);
        }
    }

    // ---- the tile's last run may spill into the following tiles: the owner follows it
    const bool spill_possible = count == SEL_TILE && base + SEL_TILE < n && tile_last_slot >= 0;
    E128 last_key_entry;
    if (spill_possible) {      // uniform over the CTA
        last_key_entry = s_e[pad_index(SEL_TILE - 1) + 1];
        uint32_t pos = base + SEL_TILE;
        while (true) {
            const uint32_t j = pos + tid;
            const bool in = j < n;
            bool match = false;
            E128 v;
            if (in) {
                v = ldg_entry(P.sorted + j);
                match = key_eq(v, last_key_entry, key_lo);
            }
            if (match) {
                atomicMax(&s_cand[tile_last_slot], cand_of(v));
                atomicAdd(&s_cnt[tile_last_slot], 1u);
                if (!PAIRS) atomicOr(&s_fl[tile_last_slot], flags_of(v));
            } else {
                // the array is sorted by key, so the matching entries are a prefix: the smallest
                // non-matching index is the end of the run
                atomicMin(&s_ext_end, in ? j : n);
            }
            if (!__syncthreads_and(match ? 1 : 0)) break;
            pos += SEL_THREADS;
        }
    }
    __syncthreads();
    const uint32_t ext_end = spill_possible ? s_ext_end : base + count;

    // ---- backward sweep: every entry learns its run's totals, then the verdict
    {
        RunAgg tot = {0ull, 0u, 0u};
        const RunAgg lead = left >= 0 ? RunAgg{s_cand[left], s_cnt[left], s_fl[left]} : RunAgg{0ull, 0u, 0u};
        const RunAgg trail = RunAgg{s_cand[tid], s_cnt[tid], s_fl[tid]};
#pragma unroll
        for (int k = SEL_ITEMS - 1; k >= 0; k--) {
            if (k < n_mine) {
                const bool run_ends_here = k == n_mine - 1 || ((heads >> (k + 1)) & 1);
                if (run_ends_here) tot = F[k];
                const bool leading = k < first_head, trailing = k >= last_head && last_head >= 0;
                RunAgg r = tot;
                if (leading) r = lead;
                if (trailing) r = trail;
                // (a leading entry with no head to its left in this tile is owned by an earlier tile)
                if (r.cnt >= 2 && !(leading && left < 0)) decide(e[k], cand_of(e[k]), r);
            }
        }
    }
    if (spill_possible) {
        const RunAgg r = RunAgg{s_cand[tile_last_slot], s_cnt[tile_last_slot], s_fl[tile_last_slot]};
        for (uint32_t j = base + SEL_TILE + tid; j < ext_end; j += SEL_THREADS) {
            const E128 v = ldg_entry(P.sorted + j);
            decide(v, cand_of(v), r);
        }
    }
    for (int o = 16; o; o >>= 1) marks += __shfl_xor_sync(0xFFFFFFFFu, marks, o);
    if (lane == 0 && marks) atomicAdd(&P.counters[CNT_MARKS], marks);
}

template <int KIND>
static int launch_select(const SelectParams &P, cudaStream_t stream, uint64_t *launches) {
    if (P.n_max == 0) return 0;
    // per device and per call: the attribute belongs to the current device's copy of the function
    OGE_CUDA_TRY(cudaFuncSetAttribute(select_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SEL_SMEM));
    const uint32_t grid = (P.n_max + SEL_TILE - 1) / SEL_TILE;
    select_kernel<KIND><<<grid, SEL_THREADS, SEL_SMEM, stream>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_select_pairs(const SelectParams &P, bool far, cudaStream_t stream, uint64_t *launches) {
    return far ? launch_select<2>(P, stream, launches) : launch_select<1>(P, stream, launches);
}
int launch_select_frags(const SelectParams &P, cudaStream_t stream, uint64_t *launches) { return launch_select<0>(P, stream, launches); }

}  // namespace oge
