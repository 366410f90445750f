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
// One CTA per 2048-entry tile.  Runs are delimited by head flags, numbered by a block scan
// and reduced with shared-memory atomics (thread-local pre-folding keeps long runs cheap).
// A run that starts in a tile is owned by that tile even when it spills into the next ones:
// the owner keeps reading until the key changes; a tile's leading entries that continue an
// earlier run are skipped.  HBM traffic: the sorted entries once (+ the spill) and one byte
// per mark.
#include "kernels.cuh"

namespace oge {

constexpr int SEL_THREADS = 256;
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;      // 2048
constexpr int SEL_PAD = SEL_TILE + SEL_TILE / 8 + 2;   // one skew slot per 8 entries: conflict-free blocked reads

constexpr uint32_t RUN_HAS_PAIRED = 1u, RUN_HAS_UNPAIRED = 2u;

// dynamic shared memory: entries (padded) + 4 per-run arrays
constexpr size_t SEL_SMEM = (size_t) SEL_PAD * sizeof(E128) + (size_t) SEL_TILE * 4 * 4;

__device__ __forceinline__ int pad_index(int j) { return j + (j >> 3); }

__device__ __forceinline__ E128 ldg_entry(const E128 *p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    E128 e;
    e.lo = v.x;
    e.hi = v.y;
    return e;
}

__device__ __forceinline__ bool key_eq(const E128 &a, const E128 &b, int key_lo) {
    E128 x = bits_from(a, key_lo), y = bits_from(b, key_lo);
    return x.lo == y.lo && x.hi == y.hi;
}

template <bool PAIRS>
__global__ void __launch_bounds__(SEL_THREADS) select_kernel(SelectParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    E128 *s_e = reinterpret_cast<E128 *>(smem_raw);
    int *s_max = reinterpret_cast<int *>(smem_raw + (size_t) SEL_PAD * sizeof(E128));
    uint32_t *s_best = reinterpret_cast<uint32_t *>(s_max + SEL_TILE);
    uint32_t *s_cnt = s_best + SEL_TILE;
    uint32_t *s_flags = s_cnt + SEL_TILE;
    __shared__ uint32_t s_wsum[SEL_THREADS / 32];
    __shared__ uint32_t s_ext_end;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = P.n_dev ? min(*P.n_dev, P.n_max) : P.n_max;
    const uint32_t base = blockIdx.x * SEL_TILE;
    if (base >= n) return;
    const uint32_t count = min((uint32_t) SEL_TILE, n - base);
    const KeyLayout &L = P.kl;
    const int key_lo = PAIRS ? L.p_coord2 : L.f_orient;
    const int idx_pos = PAIRS ? L.p_idx : L.f_idx;

    // ---- load the tile (striped, coalesced) + its predecessor into shared memory
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; k++) {
        int j = k * SEL_THREADS + tid;
        if ((uint32_t) j < count) s_e[pad_index(j) + 1] = ldg_entry(P.sorted + base + j);
        s_max[j] = INT_MIN;
        s_best[j] = 0xFFFFFFFFu;
        s_cnt[j] = 0;
        s_flags[j] = 0;
    }
    if (tid == 0) {
        if (base > 0) s_e[0] = ldg_entry(P.sorted + base - 1);
        s_ext_end = 0xFFFFFFFFu;
    }
    __syncthreads();

    // ---- blocked view: thread t owns entries [8t, 8t+8); head flags
    E128 e[SEL_ITEMS];
    uint32_t heads = 0;      // bit k: entry k starts a run
    int n_mine = 0;
    {
        E128 prev;
        int j0 = tid * SEL_ITEMS;
        if (j0 == 0) prev = s_e[0];
        else prev = s_e[pad_index(j0 - 1) + 1];
#pragma unroll
        for (int k = 0; k < SEL_ITEMS; k++) {
            int j = j0 + k;
            if ((uint32_t) j < count) {
                e[k] = s_e[pad_index(j) + 1];
                bool head = (base + j == 0) || !key_eq(e[k], prev, key_lo);
                heads |= (head ? 1u : 0u) << k;
                prev = e[k];
                n_mine = k + 1;
            }
        }
    }
    // ---- run ids: exclusive block scan of the head counts
    uint32_t hc = __popc(heads), x = hc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    uint32_t run_base = x - hc;
    for (int w = 0; w < warp; w++) run_base += s_wsum[w];
    uint32_t total_runs = 0;
    for (int w = 0; w < SEL_THREADS / 32; w++) total_runs += s_wsum[w];

    // run id of entry k = run_base + popc(heads & ((2 << k) - 1)) - 1   (-1: continues an earlier tile's run)
    auto run_of = [&](int k) { return (int) (run_base + __popc(heads & ((2u << k) - 1))) - 1; };
    auto score_of = [&](const E128 &v) { return (int) (int16_t) (uint16_t) (v.lo & 0xFFFFu); };
    // global ordinals (< 2^32); [idx_base, idx_base + n_records) are this rank's records
    auto idx_of = [&](const E128 &v) { return (uint32_t) bits_get(v, idx_pos, L.idx_bits); };
    auto mark = [&](uint32_t g) {
        const uint64_t l = (uint64_t) g - P.idx_base;      // wraps for ordinals below the base
        if (l < P.n_records) P.dup[l] = 1;
        else {
            uint32_t at = atomicAdd(&P.counters[CNT_FOREIGN_MARKS], 1u);
            if (at < P.foreign_cap) P.foreign_marks[at] = g;
        }
    };
    auto mate_of = [&](uint32_t g1) -> uint32_t {
        const uint64_t l = (uint64_t) g1 - P.idx_base;
        if (l < P.n_records) return P.mate_of[l];
        uint32_t lo = 0, hi = P.n_fm;      // first couple with idx1 >= g1
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if ((uint32_t) (P.fm[mid] >> 32) < g1) lo = mid + 1;
            else hi = mid;
        }
        return (uint32_t) P.fm[lo];
    };
    auto paired_of = [&](const E128 &v) { return PAIRS ? true : bits_get(v, L.f_paired, 1) != 0; };

    // ---- pass 1: max score, count, flags per run (thread-local folding of consecutive entries)
    {
        int cur = -2, mx = INT_MIN;
        uint32_t c = 0, fl = 0;
        auto flush = [&]() {
            if (cur >= 0) {
                atomicMax(&s_max[cur], mx);
                atomicAdd(&s_cnt[cur], c);
                if (!PAIRS) atomicOr(&s_flags[cur], fl);
            }
        };
#pragma unroll
        for (int k = 0; k < SEL_ITEMS; k++) {
            if (k < n_mine) {
                int r = run_of(k);
                if (r != cur) {
                    flush();
                    cur = r; mx = INT_MIN; c = 0; fl = 0;
                }
                mx = max(mx, score_of(e[k]));
                c++;
                fl |= paired_of(e[k]) ? RUN_HAS_PAIRED : RUN_HAS_UNPAIRED;
            }
        }
        flush();
    }

    // ---- the tile's last run may spill into the following tiles: the owner follows it
    const bool spill_possible = count == SEL_TILE && base + SEL_TILE < n && total_runs > 0;
    const int last_run = (int) total_runs - 1;
    E128 last_key_entry;
    if (spill_possible) {      // uniform over the CTA
        last_key_entry = s_e[pad_index(SEL_TILE - 1) + 1];
        uint32_t pos = base + SEL_TILE;
        while (true) {
            uint32_t j = pos + tid;
            bool in = j < n, match = false;
            E128 v;
            if (in) {
                v = ldg_entry(P.sorted + j);
                match = key_eq(v, last_key_entry, key_lo);
            }
            if (match) {
                atomicMax(&s_max[last_run], score_of(v));
                atomicAdd(&s_cnt[last_run], 1u);
                if (!PAIRS) atomicOr(&s_flags[last_run], paired_of(v) ? RUN_HAS_PAIRED : RUN_HAS_UNPAIRED);
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
    // first index past the tile's last run (== base + count when there is nothing to follow)
    const uint32_t ext_end = spill_possible ? s_ext_end : base + count;

    // ---- pass 2: smallest index among the entries holding the run's max score
    {
        int cur = -2;
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < SEL_ITEMS; k++) {
            if (k < n_mine) {
                int r = run_of(k);
                if (r != cur) {
                    if (cur >= 0 && best != 0xFFFFFFFFu) atomicMin(&s_best[cur], best);
                    cur = r; best = 0xFFFFFFFFu;
                }
                if (r >= 0 && score_of(e[k]) == s_max[r]) best = min(best, idx_of(e[k]));
            }
        }
        if (cur >= 0 && best != 0xFFFFFFFFu) atomicMin(&s_best[cur], best);
        if (spill_possible) {
            for (uint32_t j = base + SEL_TILE + tid; j < ext_end; j += SEL_THREADS) {
                E128 v = ldg_entry(P.sorted + j);
                if (score_of(v) == s_max[last_run]) atomicMin(&s_best[last_run], idx_of(v));
            }
        }
    }
    __syncthreads();

    // ---- pass 3: mark
    uint32_t marks = 0;
    auto decide = [&](const E128 &v, int r) {
        uint32_t c = s_cnt[r];
        if (c < 2) return;
        bool is_best = score_of(v) == s_max[r] && idx_of(v) == s_best[r];
        if (PAIRS) {
            if (!is_best) {
                const uint32_t i1 = idx_of(v);
                mark(i1);
                mark(mate_of(i1));
                marks += 2;
            }
        } else {
            uint32_t fl = s_flags[r];
            if (!(fl & RUN_HAS_UNPAIRED)) return;
            bool mark_it = (fl & RUN_HAS_PAIRED) ? !paired_of(v) : !is_best;
            if (mark_it) {
                mark(idx_of(v));
                marks += 1;
            }
        }
    };
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; k++) {
        if (k < n_mine) {
            int r = run_of(k);
            if (r >= 0) decide(e[k], r);
        }
    }
    if (spill_possible) {
        for (uint32_t j = base + SEL_TILE + tid; j < ext_end; j += SEL_THREADS) decide(ldg_entry(P.sorted + j), last_run);
    }
    for (int o = 16; o; o >>= 1) marks += __shfl_xor_sync(0xFFFFFFFFu, marks, o);
    if (lane == 0 && marks) atomicAdd(&P.counters[CNT_MARKS], marks);
}

static int launch_select(const SelectParams &P, bool pairs, cudaStream_t stream, uint64_t *launches) {
    if (P.n_max == 0) return 0;
    static bool configured = false;
    if (!configured) {
        OGE_CUDA_TRY(cudaFuncSetAttribute(select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SEL_SMEM));
        OGE_CUDA_TRY(cudaFuncSetAttribute(select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SEL_SMEM));
        configured = true;
    }
    uint32_t grid = (P.n_max + SEL_TILE - 1) / SEL_TILE;
    if (pairs) select_kernel<true><<<grid, SEL_THREADS, SEL_SMEM, stream>>>(P);
    else select_kernel<false><<<grid, SEL_THREADS, SEL_SMEM, stream>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_select_pairs(const SelectParams &P, cudaStream_t stream, uint64_t *launches) { return launch_select(P, true, stream, launches); }
int launch_select_frags(const SelectParams &P, cudaStream_t stream, uint64_t *launches) { return launch_select(P, false, stream, launches); }

}  // namespace oge
