// K3: LSD "onesweep" radix sort of 16-byte end entries, sm_100a.
//
// Replaces ogeSortMt(..., compareReadEnds()) (reference util/thread_pool.h:359-397,
// algorithms/mark_duplicates.cpp:262-271).  The comparison order of ReadEnds::compare
// (util/picard_structures.h:56-68) is not reproduced -- only the grouping of equal keys is
// needed downstream, and the survivor choice is order-independent (select.cu).
//
// Structure (one read of the input for all histograms, then one read + one write per pass):
//   rs_histogram      every pass's 256-bin digit histogram in one sweep (shared-memory bins)
//   rs_scan           exclusive scan of each histogram -> global digit offsets
//   rs_pass_v2        per 4096-entry tile: 128-bit coalesced loads, warp-private digit counters
//                     ranked with match.any, chained-scan (decoupled look-back) across tiles,
//                     reorder through shared memory, coalesced 128-bit stores
// The sort is stable per pass (warp-striped tile order + tile-ordered look-back), which LSD needs.
// HBM-bound integer work: no tensor cores.  Algorithmic bytes per entry: 16 (histogram) +
// 32 per pass.
#include "radix_sort.cuh"

namespace oge {

constexpr uint32_t LB_FLAG_AGG = 1u << 30;      // tile aggregate published
constexpr uint32_t LB_FLAG_INC = 2u << 30;      // inclusive prefix published
constexpr uint32_t LB_VALUE_MASK = (1u << 30) - 1;
constexpr int LB_WINDOW = 8;

SortPlan make_sort_plan(int bit_lo, int bit_hi) {
    SortPlan p;
    p.n_pass = 0;
    for (int s = bit_lo; s < bit_hi && p.n_pass < RS_MAX_PASSES; s += RS_RADIX_BITS) {
        p.shift[p.n_pass] = s;
        p.bits[p.n_pass] = (bit_hi - s) < RS_RADIX_BITS ? (bit_hi - s) : RS_RADIX_BITS;
        p.n_pass++;
    }
    return p;
}

static inline uint64_t tiles_of(uint64_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// scratch: [hist: MAXP*256 u32][offsets: MAXP*256 u32][tile counters: MAXP u32 (+pad)][status: 2 x tiles*256 u32]
// (two status arrays: pass p works on one while its tiles zero the other for pass p+1)
size_t sort_scratch_bytes(uint64_t n) {
    return (size_t) (2 * RS_MAX_PASSES * RS_RADIX + 64) * 4 + (size_t) 2 * tiles_of(n) * RS_RADIX * 4;
}

__device__ __forceinline__ E128 ld_entry(const E128 *p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    E128 e;
    e.lo = v.x;
    e.hi = v.y;
    return e;
}

__device__ __forceinline__ void st_entry(E128 *p, const E128 &e) {
    *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(e.lo, e.hi);
}

// ------------------------------------------------------------------------------------------------
// All digit histograms in one read of the entries.
constexpr int HIST_THREADS = 512;
constexpr int HIST_UNROLL = 4;

__global__ void __launch_bounds__(HIST_THREADS) rs_histogram(const E128 *__restrict__ in, uint32_t n_max,
                                                              const uint32_t *__restrict__ n_dev, SortPlan plan,
                                                              uint32_t *__restrict__ ghist) {
    __shared__ uint32_t sh[RS_MAX_PASSES * RS_RADIX];
    const uint32_t n = n_dev ? min(*n_dev, n_max) : n_max;
    const int n_pass = plan.n_pass, bit_lo = plan.shift[0];
    const int width = plan.shift[n_pass - 1] + plan.bits[n_pass - 1] - bit_lo;      // key bits, 1..128
    const uint64_t mask_lo = width >= 64 ? ~0ull : (1ull << width) - 1;
    const uint64_t mask_hi = width >= 128 ? ~0ull : (width > 64 ? (1ull << (width - 64)) - 1 : 0ull);
    for (int i = threadIdx.x; i < n_pass * RS_RADIX; i += HIST_THREADS) sh[i] = 0;
    __syncthreads();

    const uint32_t chunk = HIST_THREADS * HIST_UNROLL;
    const uint32_t n_chunks = (n + chunk - 1) / chunk;
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        E128 e[HIST_UNROLL];
        bool ok[HIST_UNROLL];
#pragma unroll
        for (int k = 0; k < HIST_UNROLL; k++) {
            uint32_t i = c * chunk + k * HIST_THREADS + threadIdx.x;
            ok[k] = i < n;
            if (ok[k]) e[k] = ld_entry(in + i);
        }
#pragma unroll
        for (int k = 0; k < HIST_UNROLL; k++) {
            // the key, shifted down to bit 0 and cut at its width: digit p is then simply byte p
            E128 key = bits_from(e[k], bit_lo);
            key.lo &= mask_lo;
            key.hi &= mask_hi;
            const uint32_t w[4] = {(uint32_t) key.lo, (uint32_t) (key.lo >> 32), (uint32_t) key.hi, (uint32_t) (key.hi >> 32)};
#pragma unroll
            for (int p = 0; p < RS_MAX_PASSES; p++) {
                if (p < n_pass) {      // uniform
                    uint32_t d = ok[k] ? ((w[p >> 2] >> (8 * (p & 3))) & 0xFFu) : 0xFFFFFFFFu;
                    // coordinate-sorted input makes the high digits warp-uniform: one add per warp then
                    int uniform;
                    __match_all_sync(0xFFFFFFFFu, d, &uniform);
                    if (uniform) {
                        if ((threadIdx.x & 31) == 0 && ok[k]) atomicAdd(&sh[p * RS_RADIX + d], 32u);
                    } else if (ok[k]) {
                        atomicAdd(&sh[p * RS_RADIX + d], 1u);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.n_pass * RS_RADIX; i += HIST_THREADS) {
        uint32_t v = sh[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

// One CTA of 256 threads per pass: exclusive scan of the 256 bins.
__global__ void __launch_bounds__(RS_RADIX) rs_scan(const uint32_t *__restrict__ ghist, uint32_t *__restrict__ goff) {
    __shared__ uint32_t wsum[RS_RADIX / 32];
    const int p = blockIdx.x, d = threadIdx.x, lane = d & 31, w = d >> 5;
    uint32_t v = ghist[p * RS_RADIX + d], x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int i = 0; i < w; i++) base += wsum[i];
    goff[p * RS_RADIX + d] = base + x - v;
}

// Lanes holding the same digit, from one ballot per digit bit (cheaper than match.any on the
// 8-bit digits: the ncu capture of the first version had every FLO behind a MATCH.ANY stalled).
// Out-of-range lanes carry 0xFFFFFFFF and are kept apart by a ninth ballot.
__device__ __forceinline__ uint32_t match_digit(uint32_t d, int bits) {
    uint32_t peers = __ballot_sync(0xFFFFFFFFu, d != 0xFFFFFFFFu);
    if (d == 0xFFFFFFFFu) peers = ~peers;
#pragma unroll
    for (int b = 0; b < RS_RADIX_BITS; b++) {
        if (b < bits) {
            bool bit = (d >> b) & 1;
            uint32_t v = __ballot_sync(0xFFFFFFFFu, bit);
            peers &= bit ? v : ~v;
        }
    }
    return peers;
}

// ------------------------------------------------------------------------------------------------
// The pass kernel (second generation; the first one, one 2048-entry tile per 256-thread CTA with 128-bit variable
// shifts, is in the history of this file and in profiles/r1_first_generation_kernels.txt).  It was issue-bound (ncu: 60 % issue slots busy, ~2000
// warp instructions per warp and tile, dram at 25 %), so this one is written for instruction count:
//   * the digit comes out of a 32-bit word pair chosen at compile time (WORD = shift / 32):
//     one funnel shift + one AND instead of a 128-bit variable shift, and it is computed once
//   * full tiles take a path without any bounds predicate
//   * the warp ranking is eight ballots per entry (one match.any per entry was measured: 2.55 vs 3.07 TB/s,
//     profiles/r1_sort_variants_ab.json); the group leader bumps the warp-private counter with a single
//     shared-memory atomic
//   * 32-bit scatter deltas
//   * THREADS = 256 or 512: tile = THREADS * 8 entries; a larger tile halves the look-backs per
//     entry and doubles the length of the contiguous runs written to HBM
template <int WORD>
__device__ __forceinline__ uint32_t digit_word(const E128 &e, int sh, uint32_t mask) {
    uint32_t a, b;
    if (WORD == 0) { a = (uint32_t) e.lo; b = (uint32_t) (e.lo >> 32); }
    else if (WORD == 1) { a = (uint32_t) (e.lo >> 32); b = (uint32_t) e.hi; }
    else if (WORD == 2) { a = (uint32_t) e.hi; b = (uint32_t) (e.hi >> 32); }
    else { a = (uint32_t) (e.hi >> 32); b = 0; }
    return __funnelshift_r(a, b, sh) & mask;
}

// lanes of the warp holding the same 8-bit digit (all lanes valid)
__device__ __forceinline__ uint32_t peers_of(uint32_t d) {
    uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < RS_RADIX_BITS; b++) {
        uint32_t v;
        asm("{\n .reg .pred p;\n .reg .b32 t;\n and.b32 t, %1, %2;\n setp.ne.u32 p, t, 0;\n vote.sync.ballot.b32 %0, p, 0xffffffff;\n"
            " @!p not.b32 %0, %0;\n}\n"
            : "=r"(v)
            : "r"(d), "r"(1u << b));
        peers &= v;
    }
    return peers;
}

template <int THREADS>
struct PassCfg {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * RS_ITEMS;
    static constexpr size_t SMEM = (size_t) TILE * sizeof(E128) + (size_t) WARPS * RS_RADIX * 4;
    static constexpr int CTAS_PER_SM = THREADS == 256 ? 4 : 2;
};

template <int THREADS, int WORD>
__global__ void __launch_bounds__(THREADS, PassCfg<THREADS>::CTAS_PER_SM)
rs_pass_v2(const E128 *__restrict__ in, E128 *__restrict__ out, uint32_t n_max, const uint32_t *__restrict__ n_dev, int shift,
           int bits, const uint32_t *__restrict__ digit_offset, uint32_t *__restrict__ status, uint32_t *__restrict__ status_next,
           uint32_t *__restrict__ tile_counter, int dbg) {
    using Cfg = PassCfg<THREADS>;
    constexpr int WARPS = Cfg::WARPS, TILE = Cfg::TILE;
    static_assert(THREADS >= RS_RADIX, "one thread per digit in the scan phase");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    E128 *s_sorted = reinterpret_cast<E128 *>(smem_raw);
    uint32_t *s_warp_cnt = reinterpret_cast<uint32_t *>(smem_raw + (size_t) TILE * sizeof(E128));
    __shared__ uint32_t s_tile_base[RS_RADIX];
    __shared__ int s_delta[RS_RADIX];
    __shared__ uint32_t s_scan[RS_RADIX / 32];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = n_dev ? min(*n_dev, n_max) : n_max;
    const uint32_t n_tiles = (n + TILE - 1) / TILE;
    const uint32_t mask = (1u << bits) - 1;
    const int sh = shift & 31;
    uint32_t *wc = s_warp_cnt + warp * RS_RADIX;

    // tile ids in launch order: every predecessor of a tile is running or done (look-back cannot starve)
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4 *>(wc)[lane] = z;
        reinterpret_cast<uint4 *>(wc)[lane + 32] = z;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= n_tiles) return;
    const uint32_t base = tile * TILE;
    const bool full = base + TILE <= n;
    const uint32_t count = full ? (uint32_t) TILE : n - base;

    // pull the tile that the CTA replacing this one will work on from HBM into L2 now, so that its
    // loads pay L2 latency instead of DRAM latency (gridDim tiles = one wave ahead is too far for
    // nothing: the L2 holds 126 MB, a wave is < 20 MB)
    if (tid == 0 && !(dbg & 8)) {
        const uint32_t ahead = tile + (uint32_t) (dbg >> 8);
        if (ahead < n_tiles) {
            const uint32_t bytes = min((uint32_t) TILE, n - ahead * TILE) * (uint32_t) sizeof(E128);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + (size_t) ahead * TILE), "r"(bytes) : "memory");
        }
    }

    // ---- load (warp-striped: 512 B contiguous per warp instruction) + digits
    E128 e[RS_ITEMS];
    uint32_t d[RS_ITEMS], rank[RS_ITEMS];
    const uint32_t wslot = warp * (32 * RS_ITEMS) + lane;      // slot of item 0 inside the tile
    const E128 *src = in + base + wslot;
    if (full) {
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k++) e[k] = ld_entry(src + k * 32);
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k++) d[k] = digit_word<WORD>(e[k], sh, mask);
    } else {
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k++) {
            d[k] = 0xFFFFFFFFu;
            if (wslot + k * 32 < count) {
                e[k] = ld_entry(src + k * 32);
                d[k] = digit_word<WORD>(e[k], sh, mask);
            }
        }
    }

    // ---- rank inside the warp
    const uint32_t lt_mask = (1u << lane) - 1;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        uint32_t peers;
        if (full) peers = peers_of(d[k]);
        else peers = match_digit(d[k], RS_RADIX_BITS);
        const int leader = 31 - __clz(peers);
        uint32_t old = 0;
        if (lane == leader && (full || d[k] != 0xFFFFFFFFu)) old = atomicAdd(&wc[d[k]], (uint32_t) __popc(peers));
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        rank[k] = old + __popc(peers & lt_mask);
    }
    __syncthreads();

    // ---- per digit: exclusive scan over warps, tile total, publish the aggregate, block scan
    uint32_t total = 0;
    uint32_t *my_status = status + (size_t) tile * RS_RADIX + tid;
    if (tid < RS_RADIX) {
        status_next[(size_t) tile * RS_RADIX + tid] = 0;      // the next pass finds its status array cleared
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            uint32_t c = s_warp_cnt[w * RS_RADIX + tid];
            s_warp_cnt[w * RS_RADIX + tid] = total;
            total += c;
        }
        {
            uint32_t word = (tile == 0 ? LB_FLAG_INC : LB_FLAG_AGG) | total;
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my_status), "r"(word) : "memory");
        }
        uint32_t x = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_scan[warp] = x;
        asm volatile("bar.sync 1, 256;" ::: "memory");      // the eight digit warps only
        uint32_t wprefix = 0;
#pragma unroll
        for (int i = 0; i < RS_RADIX / 32; i++) wprefix += i < warp ? s_scan[i] : 0u;
        s_tile_base[tid] = wprefix + x - total;
    }

    // decoupled look-back over the predecessors' status words (LB_WINDOW independent loads per round
    // trip), then the scatter delta of this digit.  It runs AFTER the shared-memory reorder: the
    // aggregate above is already visible to the successors, and the later this tile looks back the
    // more of its predecessors have published, so fewer polls come back empty.
    auto look_back = [&]() {
        uint32_t excl = 0;
        if (tile > 0 && !(dbg & 1)) {
            const uint32_t *q = my_status - RS_RADIX;      // predecessor's word for this digit
            uint32_t left = tile;                          // predecessors not yet folded in
            while (true) {
                uint32_t s[LB_WINDOW];
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++) {
                    s[j] = 0;
                    if ((uint32_t) j < left) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(s[j]) : "l"(q - j * RS_RADIX) : "memory");
                }
                bool done = false;
                int used = 0;
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++) {
                    uint32_t f = s[j] >> 30;
                    if (!done && used == j && f != 0) {
                        excl += s[j] & LB_VALUE_MASK;
                        used = j + 1;
                        done = f == 2;
                    }
                }
                if (done) break;
                q -= used * RS_RADIX;
                left -= used;
                if (used < LB_WINDOW) __nanosleep(20);
            }
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my_status), "r"(LB_FLAG_INC | (excl + total)) : "memory");
        }
        s_delta[tid] = (dbg & 2) ? (int) base : (int) (digit_offset[tid] + excl) - (int) s_tile_base[tid];
    };
    const bool late = !(dbg & 4);
    if (!late && tid < RS_RADIX) look_back();
    __syncthreads();

    // ---- reorder through shared memory
    if (full) {
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k++) st_entry(s_sorted + (s_tile_base[d[k]] + wc[d[k]] + rank[k]), e[k]);
    } else {
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k++)
            if (d[k] != 0xFFFFFFFFu) st_entry(s_sorted + (s_tile_base[d[k]] + wc[d[k]] + rank[k]), e[k]);
    }
    if (late && tid < RS_RADIX) look_back();
    __syncthreads();

    // ---- store: consecutive threads write consecutive sorted slots; a digit's run is contiguous in the output
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t j = k * THREADS + tid;
        if (full || j < count) {
            ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(s_sorted + j);
            E128 ev;
            ev.lo = v.x;
            ev.hi = v.y;
            st_entry(out + ((int) j + s_delta[digit_word<WORD>(ev, sh, mask)]), ev);
        }
    }
}

// variant: 0 = 256-thread CTAs (2048-entry tiles), 2 = 512-thread CTAs (4096-entry tiles, default)
static int g_sort_variant = 2;
static int g_sort_prefetch = 0;      // tiles ahead (0 = one wave)
void radix_sort_set_prefetch(int tiles) { g_sort_prefetch = tiles; }
void radix_sort_set_variant(int v) { g_sort_variant = v; }
int radix_sort_get_variant() { return g_sort_variant; }

template <int THREADS>
static cudaError_t launch_pass_v2(int word, uint32_t grid, cudaStream_t stream, const E128 *src, E128 *dst, uint32_t n,
                                  const uint32_t *n_dev, int shift, int bits, const uint32_t *goff, uint32_t *status, uint32_t *status_next, uint32_t *tc,
                                  int dbg) {
    const size_t smem = PassCfg<THREADS>::SMEM;
    switch (word) {
        case 0: rs_pass_v2<THREADS, 0><<<grid, THREADS, smem, stream>>>(src, dst, n, n_dev, shift, bits, goff, status, status_next, tc, dbg); break;
        case 1: rs_pass_v2<THREADS, 1><<<grid, THREADS, smem, stream>>>(src, dst, n, n_dev, shift, bits, goff, status, status_next, tc, dbg); break;
        case 2: rs_pass_v2<THREADS, 2><<<grid, THREADS, smem, stream>>>(src, dst, n, n_dev, shift, bits, goff, status, status_next, tc, dbg); break;
        default: rs_pass_v2<THREADS, 3><<<grid, THREADS, smem, stream>>>(src, dst, n, n_dev, shift, bits, goff, status, status_next, tc, dbg); break;
    }
    return cudaGetLastError();
}

template <int THREADS>
static cudaError_t init_pass_v2() {
    const int smem = (int) PassCfg<THREADS>::SMEM;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(rs_pass_v2<THREADS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(rs_pass_v2<THREADS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(rs_pass_v2<THREADS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(rs_pass_v2<THREADS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

int radix_sort_init() {
    OGE_CUDA_TRY((init_pass_v2<256>()));
    OGE_CUDA_TRY((init_pass_v2<512>()));
    return 0;
}

int radix_sort_128(E128 *a, E128 *b, uint64_t n, const uint32_t *n_dev, int bit_lo, int bit_hi, void *scratch,
                   cudaStream_t stream, E128 **result, uint64_t *launches, PassTimer *timer) {
    *result = a;
    if (n == 0 || bit_hi <= bit_lo) return 0;
    if (n >= (1ull << 30)) return fail_msg(-6, "radix sort: more than 2^30-1 entries");
    SortPlan plan = make_sort_plan(bit_lo, bit_hi);
    uint32_t *hist = (uint32_t *) scratch;
    uint32_t *goff = hist + RS_MAX_PASSES * RS_RADIX;
    uint32_t *tile_counters = goff + RS_MAX_PASSES * RS_RADIX;
    uint32_t *status = tile_counters + 64;
    const uint64_t tiles = tiles_of(n);

    OGE_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t) (2 * RS_MAX_PASSES * RS_RADIX + 64) * 4, stream));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t chunks = (n + HIST_THREADS * HIST_UNROLL - 1) / (HIST_THREADS * HIST_UNROLL);
    uint32_t hgrid = (uint32_t) (chunks < (uint64_t) sms * 4 ? chunks : (uint64_t) sms * 4);
    rs_histogram<<<hgrid, HIST_THREADS, 0, stream>>>(a, (uint32_t) n, n_dev, plan, hist);
    rs_scan<<<plan.n_pass, RS_RADIX, 0, stream>>>(hist, goff);
    *launches += 2;

    E128 *src = a, *dst = b;
    // dbg bits 0-2 are measurement knobs (bits 0, 1 give wrong results: -DOGE_TESTING builds only); bits 8.. = prefetch distance in tiles
#ifdef OGE_TESTING
    int dbg = (g_sort_variant >> 4) & 15;
#else
    int dbg = 0;
#endif
    const int variant = g_sort_variant & 2;
    const uint64_t vtiles = variant ? (n + PassCfg<512>::TILE - 1) / PassCfg<512>::TILE : tiles;
    {
        int pf = g_sort_prefetch > 0 ? g_sort_prefetch : sms * (variant ? 2 : 4);
        dbg |= pf << 8;
    }
    for (int p = 0; p < plan.n_pass; p++) {
        uint32_t *st = status + (size_t) (p & 1) * tiles * RS_RADIX, *st_next = status + (size_t) ((p + 1) & 1) * tiles * RS_RADIX;
        if (p == 0) OGE_CUDA_TRY(cudaMemsetAsync(st, 0, (size_t) tiles * RS_RADIX * 4, stream));
        const bool timed = timer && timer->used < timer->cap;
        if (timed) cudaEventRecord(timer->pool[2 * timer->used], stream);
        const int shift = plan.shift[p], bits = plan.bits[p];
        uint32_t *tc = tile_counters + p;
        const uint32_t *go = goff + p * RS_RADIX;
        if (variant) OGE_CUDA_TRY((launch_pass_v2<512>(shift >> 5, (uint32_t) vtiles, stream, src, dst, (uint32_t) n, n_dev, shift, bits, go, st, st_next, tc, dbg)));
        else OGE_CUDA_TRY((launch_pass_v2<256>(shift >> 5, (uint32_t) vtiles, stream, src, dst, (uint32_t) n, n_dev, shift, bits, go, st, st_next, tc, dbg)));
        if (timed) {
            cudaEventRecord(timer->pool[2 * timer->used + 1], stream);
            timer->used++;
            timer->bytes += n * 2 * sizeof(E128);
        }
        *launches += 1;
        E128 *t = src;
        src = dst;
        dst = t;
    }
    OGE_CUDA_TRY(cudaGetLastError());
    *result = src;
    return 0;
}

}  // namespace oge
