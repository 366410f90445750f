// f3: coordinate sort of the resident records in front of the duplicate marking (`openge mergesort -M`).
//
// Replaces ReadSorter (reference algorithms/read_sorter.cpp:203-205,248-: runs of 200 000 records sorted with
// ogeSortMt(..., Sort::ByPosition()), spilled to temp files, merged through a multiset of the same comparator) by
// one device sort of everything that is resident.  Order = Sort::ByPosition (util/bamtools/Sort.h:108-133):
//   refID (records without one, refID -1, last and equivalent to each other), position, strand (forward first),
//   name (std::string compare), flag; what the reference breaks by the ADDRESSES of its heap objects (exact copies, and
//   the order inside the unplaced tail) stays in input order here (the same convention the test suite's CPU restatement uses).
// Shape:
//   1  one 16-byte entry per record, key (ref', pos + 2^31, strand) above the ordinal; K3 (stable LSD onesweep) on the key
//   2  records that tie on that key (duplicates, mostly) are refined by name without a comparison sort: they are taken
//      out into a list keyed (run, 8 name bytes big-endian), radix-sorted, the runs split where the chunk differs, and
//      so on chunk by chunk until nothing ties or the names end; a last round does the same with the flag word.  LSD
//      stability keeps full ties in input order.  The refined list goes back into the tied positions.
//   3  record lengths in sorted order -> exclusive scan -> byte gather into a second record buffer, which becomes the
//      context's record array; the permutation stays available (oge_gpu_dedup_sort_order).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ctx.cuh"
#include "pairing.cuh"

namespace oge {

constexpr int CS_THREADS = 256;
constexpr int CS_SCAN_ITEMS = 4;
constexpr int CS_SCAN_TILE = CS_THREADS * CS_SCAN_ITEMS;

// ---- entry layouts --------------------------------------------------------------------------------
// main list:  lo[0,32) ordinal | bit 32 strand | bits [33,65) pos + 2^31 | bits [65, 65 + rb) ref'
// tied list:  lo[0,32) ordinal | bits [32,96) name chunk (big-endian) | bits [96,128) run
__device__ __forceinline__ uint64_t cs_key_lo(const E128 &e) { return e.lo >> 32; }      // strand + low 31 bits of pos'
__device__ __forceinline__ bool cs_same_key(const E128 &a, const E128 &b) { return (a.lo >> 32) == (b.lo >> 32) && a.hi == b.hi; }

__global__ void __launch_bounds__(CS_THREADS) cs_keys_kernel(const uint8_t *__restrict__ rec, const uint64_t *__restrict__ off, uint64_t n,
                                                             int ref_bits, E128 *__restrict__ out) {
    const uint64_t i = (uint64_t) blockIdx.x * CS_THREADS + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = rec + off[i];
    const int32_t ref = (int32_t) ldg_u32_unaligned(p + 4);
    E128 e;
    if (ref == -1) {      // unplaced: behind every reference, equivalent to each other (Sort.h:119-120)
        e.lo = i;
        e.hi = (((1ull << ref_bits) - 1) << 1);
    } else {
        const uint32_t pos = ldg_u32_unaligned(p + 8) ^ 0x80000000u;      // signed order as unsigned
        const uint32_t flag = ldg_u32_unaligned(p + 16) >> 16;
        const uint64_t strand = (flag >> 4) & 1;
        e.lo = i | (strand << 32) | ((uint64_t) (pos & 0x7FFFFFFFu) << 33);
        e.hi = (uint64_t) (pos >> 31) | ((uint64_t) (uint32_t) ref << 1);
    }
    reinterpret_cast<ulonglong2 *>(out)[i] = make_ulonglong2(e.lo, e.hi);
}

// tied[k] = 1: sorted position k shares its key with a neighbour (and is not in the unplaced tail); head[k] = 1: first of its run
__global__ void __launch_bounds__(CS_THREADS) cs_tied_kernel(const E128 *__restrict__ s, uint64_t n, int ref_bits, uint32_t *__restrict__ tied,
                                                             uint32_t *__restrict__ head) {
    const uint64_t k = (uint64_t) blockIdx.x * CS_THREADS + threadIdx.x;
    if (k >= n) return;
    const E128 e = s[k];
    const bool unplaced = (e.hi >> 1) == ((1ull << ref_bits) - 1);
    bool same_prev = false, same_next = false;
    if (!unplaced) {
        if (k > 0) same_prev = cs_same_key(s[k - 1], e);
        if (k + 1 < n) same_next = cs_same_key(s[k + 1], e);
    }
    tied[k] = (same_prev || same_next) ? 1u : 0u;
    head[k] = (!same_prev && same_next) ? 1u : 0u;
}

// ---- exclusive scan of 32-bit values into 64-bit sums (three launches) --------------------------------
template <class F>
__global__ void __launch_bounds__(CS_THREADS) cs_scan_sums_kernel(F f, uint64_t n, uint64_t *__restrict__ bsum) {
    __shared__ uint64_t s_w[CS_THREADS / 32];
    const uint64_t base = (uint64_t) blockIdx.x * CS_SCAN_TILE;
    uint64_t v = 0;
#pragma unroll
    for (int q = 0; q < CS_SCAN_ITEMS; q++) {
        const uint64_t i = base + (uint64_t) q * CS_THREADS + threadIdx.x;
        if (i < n) v += f(i);
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < CS_THREADS / 32; w++) t += s_w[w];
        bsum[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan of the block sums in place; total -> bsum[n_blocks]
__global__ void __launch_bounds__(1024) cs_scan_blocks_kernel(uint64_t *__restrict__ bsum, uint64_t n_blocks) {
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n_blocks; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < n_blocks ? bsum[i] : 0;
        uint64_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint64_t w = s_w[threadIdx.x], xs = w;
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, xs, o);
                if (threadIdx.x >= o) xs += y;
            }
            s_w[threadIdx.x] = xs - w;
        }
        __syncthreads();
        const uint64_t excl = s_carry + s_w[threadIdx.x >> 5] + x - v;
        if (i < n_blocks) bsum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[n_blocks] = s_carry;
}

template <class F, class T>
__global__ void __launch_bounds__(CS_THREADS) cs_scan_apply_kernel(F f, uint64_t n, const uint64_t *__restrict__ bsum, T *__restrict__ out) {
    // items are taken q-major (i = base + q * THREADS + tid), so the order of the sums is: all of q = 0, then q = 1, ...
    __shared__ uint64_t s_w[CS_SCAN_ITEMS][CS_THREADS / 32];
    const uint64_t base = (uint64_t) blockIdx.x * CS_SCAN_TILE;
    uint64_t v[CS_SCAN_ITEMS], x[CS_SCAN_ITEMS];
#pragma unroll
    for (int q = 0; q < CS_SCAN_ITEMS; q++) {
        const uint64_t i = base + (uint64_t) q * CS_THREADS + threadIdx.x;
        v[q] = i < n ? (uint64_t) f(i) : 0ull;
        x[q] = v[q];
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x[q], o);
            if ((threadIdx.x & 31) >= o) x[q] += y;
        }
        if ((threadIdx.x & 31) == 31) s_w[q][threadIdx.x >> 5] = x[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int q = 0; q < CS_SCAN_ITEMS; q++)
            for (int w = 0; w < CS_THREADS / 32; w++) {
                const uint64_t c = s_w[q][w];
                s_w[q][w] = t;
                t += c;
            }
    }
    __syncthreads();
    const uint64_t b0 = bsum[blockIdx.x];
#pragma unroll
    for (int q = 0; q < CS_SCAN_ITEMS; q++) {
        const uint64_t i = base + (uint64_t) q * CS_THREADS + threadIdx.x;
        if (i < n) out[i] = (T) (b0 + s_w[q][threadIdx.x >> 5] + x[q] - v[q]);
    }
}

struct ArrU32 {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return p[i]; }
};
// length of the record at sorted position k
struct RecLen {
    const E128 *s;
    const uint64_t *off;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const {
        const uint32_t o = (uint32_t) s[k].lo;
        return (uint32_t) (off[o + 1] - off[o]);
    }
};

template <class F, class T>
static int cs_scan(F f, uint64_t n, uint64_t *bsum, T *out, uint64_t *total_host, cudaStream_t st, uint64_t *launches) {
    if (n == 0) { if (total_host) *total_host = 0; return 0; }
    const uint64_t nb = (n + CS_SCAN_TILE - 1) / CS_SCAN_TILE;
    cs_scan_sums_kernel<<<(uint32_t) nb, CS_THREADS, 0, st>>>(f, n, bsum);
    cs_scan_blocks_kernel<<<1, 1024, 0, st>>>(bsum, nb);
    cs_scan_apply_kernel<<<(uint32_t) nb, CS_THREADS, 0, st>>>(f, n, bsum, out);
    *launches += 3;
    OGE_CUDA_TRY(cudaGetLastError());
    if (total_host) {
        OGE_CUDA_TRY(cudaMemcpyAsync(total_host, bsum + nb, 8, cudaMemcpyDeviceToHost, st));
        OGE_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---- the tied list ------------------------------------------------------------------------------------
// t[tpos[k]] = (ordinal, run) of every tied position k; run = number of run heads at or before k (1-based, monotone)
__global__ void __launch_bounds__(CS_THREADS) cs_take_tied_kernel(const E128 *__restrict__ s, uint64_t n, const uint32_t *__restrict__ tied,
                                                                  const uint32_t *__restrict__ tpos, const uint32_t *__restrict__ hpos,
                                                                  const uint32_t *__restrict__ head, E128 *__restrict__ t,
                                                                  const uint8_t *__restrict__ rec, const uint64_t *__restrict__ off,
                                                                  uint32_t *__restrict__ max_name) {
    const uint64_t k = (uint64_t) blockIdx.x * CS_THREADS + threadIdx.x;
    uint32_t l = 0;
    if (k < n && tied[k]) {
        const uint32_t ord = (uint32_t) s[k].lo;
        E128 e;
        e.lo = ord;
        e.hi = (uint64_t) (hpos[k] + head[k]) << 32;      // exclusive count of heads before k, + 1 when k is one
        reinterpret_cast<ulonglong2 *>(t)[tpos[k]] = make_ulonglong2(e.lo, e.hi);
        const uint32_t ln = rec[off[ord] + 12];
        l = ln ? ln - 1 : 0;
    }
    for (int o = 16; o; o >>= 1) l = max(l, __shfl_xor_sync(0xFFFFFFFFu, l, o));
    if ((threadIdx.x & 31) == 0 && l) atomicMax(max_name, l);
}

// chunk j of the name (bytes [8j, 8j + 8), big-endian, zero beyond the name) or, mode 1, the flag word.
// Elements of runs that are already singletons need no chunk (nothing can change their place): zero.
__global__ void __launch_bounds__(CS_THREADS) cs_chunk_kernel(E128 *__restrict__ t, uint32_t n_t, const uint8_t *__restrict__ rec,
                                                              const uint64_t *__restrict__ off, uint32_t j, int mode) {
    const uint32_t m = blockIdx.x * CS_THREADS + threadIdx.x;
    if (m >= n_t) return;
    E128 e = t[m];
    const uint32_t run = (uint32_t) (e.hi >> 32);
    bool alone = true;
    if (m > 0 && (uint32_t) (t[m - 1].hi >> 32) == run) alone = false;
    if (m + 1 < n_t && (uint32_t) (t[m + 1].hi >> 32) == run) alone = false;
    uint64_t chunk = 0;
    if (!alone) {
        const uint8_t *p = rec + off[(uint32_t) e.lo];
        if (mode == 1) {
            chunk = (uint64_t) p[18] | ((uint64_t) p[19] << 8);
        } else {
            const uint32_t ln = p[12], nlen = ln ? ln - 1 : 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const uint32_t at = 8 * j + b;
                chunk = (chunk << 8) | (at < nlen ? (uint64_t) p[36 + at] : 0ull);
            }
        }
    }
    e.lo = (e.lo & 0xFFFFFFFFull) | (chunk << 32);
    e.hi = (e.hi & 0xFFFFFFFF00000000ull) | (chunk >> 32);
    reinterpret_cast<ulonglong2 *>(t)[m] = make_ulonglong2(e.lo, e.hi);
}

// after the sort by (run, chunk): newhead[m] = 1 where (run, chunk) changes; counts the elements that still tie
__global__ void __launch_bounds__(CS_THREADS) cs_split_kernel(const E128 *__restrict__ t, uint32_t n_t, uint32_t *__restrict__ newhead,
                                                              uint32_t *__restrict__ unresolved) {
    const uint32_t m = blockIdx.x * CS_THREADS + threadIdx.x;
    uint32_t same = 0;
    if (m < n_t) {
        if (m > 0) {
            const E128 a = t[m - 1], b = t[m];
            same = (a.hi == b.hi && (a.lo >> 32) == (b.lo >> 32)) ? 1u : 0u;
        }
        newhead[m] = same ? 0u : 1u;
    }
    const uint32_t any = __ballot_sync(0xFFFFFFFFu, same);
    if ((threadIdx.x & 31) == 0 && any) atomicAdd(unresolved, (uint32_t) __popc(any));
}

// run <- 1 + number of heads before m (the scan gave the exclusive count; a head adds itself)
__global__ void __launch_bounds__(CS_THREADS) cs_rerun_kernel(E128 *__restrict__ t, uint32_t n_t, const uint32_t *__restrict__ hpos,
                                                              const uint32_t *__restrict__ newhead) {
    const uint32_t m = blockIdx.x * CS_THREADS + threadIdx.x;
    if (m >= n_t) return;
    E128 e = t[m];
    e.hi = (e.hi & 0xFFFFFFFFull) | ((uint64_t) (hpos[m] + newhead[m]) << 32);
    reinterpret_cast<ulonglong2 *>(t)[m] = make_ulonglong2(e.lo, e.hi);
}

// the refined ordinals go back into the tied positions; perm[k] = ordinal at sorted position k
__global__ void __launch_bounds__(CS_THREADS) cs_perm_kernel(const E128 *__restrict__ s, uint64_t n, const uint32_t *__restrict__ tied,
                                                             const uint32_t *__restrict__ tpos, const E128 *__restrict__ t,
                                                             uint32_t *__restrict__ perm) {
    const uint64_t k = (uint64_t) blockIdx.x * CS_THREADS + threadIdx.x;
    if (k >= n) return;
    perm[k] = (t && tied[k]) ? (uint32_t) t[tpos[k]].lo : (uint32_t) s[k].lo;
}

struct PermLen {
    const uint32_t *perm;
    const uint64_t *off;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const {
        const uint32_t o = perm[k];
        return (uint32_t) (off[o + 1] - off[o]);
    }
};

// one warp per record: bytes of record perm[k] to new_off[k]
__global__ void __launch_bounds__(CS_THREADS) cs_gather_kernel(const uint8_t *__restrict__ rec, const uint64_t *__restrict__ off,
                                                               const uint32_t *__restrict__ perm, const uint64_t *__restrict__ new_off, uint64_t n,
                                                               uint8_t *__restrict__ out) {
    const uint64_t k = ((uint64_t) blockIdx.x * CS_THREADS + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (k >= n) return;
    const uint32_t o = perm[k];
    const uint8_t *src = rec + off[o];
    uint8_t *dst = out + new_off[k];
    const uint32_t len = (uint32_t) (off[o + 1] - off[o]);
    // head bytes up to a 4-byte boundary of the destination, then words (source read unaligned), then the tail
    const uint32_t lead = min(len, (uint32_t) ((4 - ((uintptr_t) dst & 3)) & 3));
    if (lane < lead) dst[lane] = src[lane];
    const uint32_t words = (len - lead) >> 2;
    for (uint32_t w = lane; w < words; w += 32)
        *reinterpret_cast<uint32_t *>(dst + lead + 4 * w) = ldg_u32_unaligned(src + lead + 4 * w);
    const uint32_t done = lead + 4 * words;
    if (done + lane < len) dst[done + lane] = src[done + lane];
}

}  // namespace oge

using namespace oge;

extern "C" int oge_gpu_dedup_sort(oge_gpu_dedup_ctx *c) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "sort: null context");
    if (c->sh.on) return fail_msg(OGE_ERR_STATE, "sort: not available on a range-sharded context");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream;
    const uint64_t n = c->n;
    c->ran = false;
    c->sorted = true;
    c->sort_stats[0] = c->sort_stats[1] = c->sort_stats[2] = 0;
    if (n == 0) return OGE_OK;
    if (n >= (1ull << 30)) return fail_msg(OGE_ERR_TOO_LARGE, "sort: more than 2^30-1 records");
    OGE_CUDA_TRY(cudaEventRecord(c->copy_done, c->copy_stream));
    OGE_CUDA_TRY(cudaStreamWaitEvent(s, c->copy_done, 0));
    OGE_CUDA_TRY(cudaEventRecord(c->ev[8], s));
    int rc;
    uint64_t launches = 0;
    // every refID below the all-ones code of "unplaced"; with no reference count in the config the whole 32-bit field is
    // used, where -1 IS the all-ones code
    int ref_bits = 1;
    if (c->cfg.n_ref <= 0) ref_bits = 32;
    else while ((1ull << ref_bits) - 1 <= (uint64_t) c->cfg.n_ref) ref_bits++;
    // work arrays: the dedup run's own buffers are free at this point
    if ((rc = c->frag.reserve(n, false, s)) || (rc = c->sortbuf.reserve(n, false, s)) ||
        (rc = c->scratch.reserve(std::max(sort_scratch_bytes(n), compact_scratch_bytes(n)), false, s)) ||
        (rc = c->cs_u32.reserve(4 * n + 16, false, s)) || (rc = c->cs_bsum.reserve(n / CS_SCAN_TILE + 4, false, s)) ||
        (rc = c->perm.reserve(n, false, s)) || (rc = c->off2.reserve(n + 1, false, s)) || (rc = c->rec2.reserve(c->rec_bytes + 256, false, s)))
        return rc;
    const uint32_t grid = (uint32_t) ((n + CS_THREADS - 1) / CS_THREADS);
    cs_keys_kernel<<<grid, CS_THREADS, 0, s>>>(c->recs(), c->off.p, n, ref_bits, c->frag.p);
    launches++;
    E128 *S = nullptr;
    if ((rc = radix_sort_128(c->frag.p, c->sortbuf.p, n, nullptr, 32, 65 + ref_bits, c->scratch.p, s, &S, &launches))) return rc;
    E128 *spare = S == c->frag.p ? c->sortbuf.p : c->frag.p;
    uint32_t *tied = c->cs_u32.p, *head = tied + n, *tpos = head + n, *hpos = tpos + n, *cnt = hpos + n;      // cnt: [0] still tied, [1] longest name
    cs_tied_kernel<<<grid, CS_THREADS, 0, s>>>(S, n, ref_bits, tied, head);
    launches++;
    uint64_t n_tied = 0, n_runs = 0;
    if ((rc = cs_scan(ArrU32{tied}, n, c->cs_bsum.p, tpos, &n_tied, s, &launches))) return rc;
    if ((rc = cs_scan(ArrU32{head}, n, c->cs_bsum.p, hpos, &n_runs, s, &launches))) return rc;
    E128 *T = nullptr;
    uint32_t rounds = 0;
    if (n_tied) {
        // the tied list lives in the spare half of the main sort's buffers (n_tied <= n); its ping-pong buffer in `pair`
        if ((rc = c->pair.reserve(n_tied, false, s))) return rc;
        OGE_CUDA_TRY(cudaMemsetAsync(cnt, 0, 8, s));
        cs_take_tied_kernel<<<grid, CS_THREADS, 0, s>>>(S, n, tied, tpos, hpos, head, spare, c->recs(), c->off.p, cnt + 1);
        launches++;
        uint32_t h_cnt[2] = {0, 0};
        OGE_CUDA_TRY(cudaMemcpyAsync(h_cnt, cnt, 8, cudaMemcpyDeviceToHost, s));
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
        const uint32_t name_rounds = (h_cnt[1] + 7) / 8;
        int run_bits = 1;
        while ((1ull << run_bits) <= n_tied + 1) run_bits++;
        const uint32_t tgrid = (uint32_t) ((n_tied + CS_THREADS - 1) / CS_THREADS);
        uint32_t *nh = head, *nhp = hpos;      // per-round arrays over the tied list: the main list's head arrays are done with
        T = spare;
        E128 *T2 = c->pair.p;
        for (uint32_t j = 0; j <= name_rounds; j++) {      // name chunks, then the flag word
            const int mode = j == name_rounds ? 1 : 0;
            cs_chunk_kernel<<<tgrid, CS_THREADS, 0, s>>>(T, (uint32_t) n_tied, c->recs(), c->off.p, j, mode);
            launches++;
            E128 *res = nullptr;
            if ((rc = radix_sort_128(T, T2, n_tied, nullptr, 32, 96 + run_bits, c->scratch.p, s, &res, &launches))) return rc;
            if (res != T) { T2 = T; T = res; }
            OGE_CUDA_TRY(cudaMemsetAsync(cnt, 0, 4, s));
            cs_split_kernel<<<tgrid, CS_THREADS, 0, s>>>(T, (uint32_t) n_tied, nh, cnt);
            launches++;
            if ((rc = cs_scan(ArrU32{nh}, n_tied, c->cs_bsum.p, nhp, (uint64_t *) nullptr, s, &launches))) return rc;
            cs_rerun_kernel<<<tgrid, CS_THREADS, 0, s>>>(T, (uint32_t) n_tied, nhp, nh);
            launches++;
            OGE_CUDA_TRY(cudaMemcpyAsync(h_cnt, cnt, 4, cudaMemcpyDeviceToHost, s));
            OGE_CUDA_TRY(cudaStreamSynchronize(s));
            rounds++;
            if (h_cnt[0] == 0) break;      // nothing ties any more
        }
    }
    cs_perm_kernel<<<grid, CS_THREADS, 0, s>>>(S, n, tied, tpos, T, c->perm.p);
    launches++;
    // ---- records into sorted order
    uint64_t total = 0;
    if ((rc = cs_scan(PermLen{c->perm.p, c->off.p}, n, c->cs_bsum.p, c->off2.p, &total, s, &launches))) return rc;
    if (total != c->rec_bytes) return fail_msg(OGE_ERR_STATE, "sort: record lengths sum to %llu, the context holds %llu bytes", (unsigned long long) total, (unsigned long long) c->rec_bytes);
    OGE_CUDA_TRY(cudaMemcpyAsync(c->off2.p + n, &total, 8, cudaMemcpyHostToDevice, s));
    cs_gather_kernel<<<(uint32_t) ((n * 32 + CS_THREADS - 1) / CS_THREADS), CS_THREADS, 0, s>>>(c->recs(), c->off.p, c->perm.p, c->off2.p, n, c->rec2.p);
    launches++;
    OGE_CUDA_TRY(cudaGetLastError());
    OGE_CUDA_TRY(cudaEventRecord(c->ev[9], s));
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    std::swap(c->rec.p, c->rec2.p);
    std::swap(c->rec.cap, c->rec2.cap);
    std::swap(c->off.p, c->off2.p);
    std::swap(c->off.cap, c->off2.cap);
    c->rec_lead = 0;
    c->sort_stats[0] = n_tied;
    c->sort_stats[1] = rounds;
    c->sort_stats[2] = launches;
    c->ms_sort_records = ms_between(c->ev[8], c->ev[9]);
    return OGE_OK;
}

extern "C" int oge_gpu_dedup_sort_order(oge_gpu_dedup_ctx *c, uint32_t *perm_out, uint64_t n) {
    if (!c || (n && !perm_out)) return fail_msg(OGE_ERR_INVALID_ARG, "sort_order: null argument");
    if (n != c->n || c->perm.cap < n) return fail_msg(OGE_ERR_STATE, "sort_order: call oge_gpu_dedup_sort first (the context holds %llu records)", (unsigned long long) c->n);
    if (n == 0) return OGE_OK;
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    OGE_CUDA_TRY(cudaMemcpyAsync(perm_out, c->perm.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OGE_OK;
}

extern "C" int oge_gpu_dedup_sort_stats(oge_gpu_dedup_ctx *c, uint64_t *n_tied, uint64_t *rounds, uint64_t *launches, float *ms) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "sort_stats: null context");
    if (n_tied) *n_tied = c->sort_stats[0];
    if (rounds) *rounds = c->sort_stats[1];
    if (launches) *launches = c->sort_stats[2];
    if (ms) *ms = c->ms_sort_records;
    return OGE_OK;
}
