// K5: flag write.  Replaces the rewrite loop of MarkDuplicates::runInternal (reference
// algorithms/mark_duplicates.cpp:443-465) and SetIsDuplicate (util/bamtools/BamAlignment.cpp:600-603):
// every primary record gets bit 0x400 set if its ordinal is in the duplicate set and CLEARED
// otherwise; secondary records (0x100) are left untouched.  The new flag word goes to a compact
// u16 array (what the host reads back) and is scattered into the device-resident record only
// where it changed.  Also here: the compaction behind pull() with remove_duplicates (:456-458).
#include "kernels.cuh"

namespace oge {

constexpr int FL_THREADS = 256;
constexpr int FL_ITEMS = 8;      // records per thread: one 16-byte load of flags, one 8-byte load of marks

__device__ __forceinline__ uint32_t flag_of(const uint4 &v, int k) {
    const uint32_t w = k < 2 ? v.x : (k < 4 ? v.y : (k < 6 ? v.z : v.w));
    return (k & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// The kernel is latency-bound by construction (flag -> mark -> offset -> store is a dependent chain
// per record), so every thread takes eight records with all of its loads issued up front.
__global__ void __launch_bounds__(FL_THREADS) flags_kernel(FlagParams P) {
    const uint64_t i0 = ((uint64_t) blockIdx.x * FL_THREADS + threadIdx.x) * FL_ITEMS;
    uint32_t n_dup = 0;
    if (i0 + FL_ITEMS <= P.n) {
        const uint4 fi = *reinterpret_cast<const uint4 *>(P.flag_in + i0);
        const uint2 dm = *reinterpret_cast<const uint2 *>(P.dup + i0);
        const bool quiet_hit = P.quiet_index_bug && i0 == 0 && P.counters[CNT_MARKS] > 0;   // every index is 0
        uint32_t out[FL_ITEMS], changed = 0;
#pragma unroll
        for (int k = 0; k < FL_ITEMS; k++) {
            const uint32_t f = flag_of(fi, k);
            uint32_t nf = f;
            if (!(f & 0x100)) {
                bool d;
                if (P.quiet_index_bug) d = quiet_hit && k == 0;
                else d = (((k < 4 ? dm.x : dm.y) >> (8 * (k & 3))) & 0xFFu) != 0;
                nf = d ? (f | 0x400u) : (f & ~0x400u);
                n_dup += d;
                // duplicates are always (re)written so that a re-run over the resident records does the same work
                if (nf != f || d) changed |= 1u << k;
            }
            out[k] = nf;
        }
        *reinterpret_cast<uint4 *>(P.flag_out + i0) =
            make_uint4(out[0] | (out[1] << 16), out[2] | (out[3] << 16), out[4] | (out[5] << 16), out[6] | (out[7] << 16));
        // scatter into the resident records: all offsets first, then the stores
        uint64_t o[FL_ITEMS];
#pragma unroll
        for (int k = 0; k < FL_ITEMS; k++)
            if (changed & (1u << k)) o[k] = P.off[i0 + k];
#pragma unroll
        for (int k = 0; k < FL_ITEMS; k++)
            if (changed & (1u << k)) {
                uint8_t *p = P.rec + o[k] + 18;
                p[0] = (uint8_t) (out[k] & 0xFF);
                p[1] = (uint8_t) (out[k] >> 8);
            }
    } else {
        for (uint64_t i = i0; i < P.n; i++) {      // the last, partial group
            const uint32_t f = P.flag_in[i];
            uint32_t nf = f;
            bool d = false;
            if (!(f & 0x100)) {
                if (P.quiet_index_bug) d = (i == 0) && P.counters[CNT_MARKS] > 0;
                else d = P.dup[i] != 0;
                nf = d ? (f | 0x400u) : (f & ~0x400u);
                n_dup += d;
            }
            P.flag_out[i] = (uint16_t) nf;
            if (nf != f || d) {
                uint8_t *p = P.rec + P.off[i] + 18;
                p[0] = (uint8_t) (nf & 0xFF);
                p[1] = (uint8_t) (nf >> 8);
            }
        }
    }
    for (int o = 16; o; o >>= 1) n_dup += __shfl_xor_sync(0xFFFFFFFFu, n_dup, o);
    __shared__ uint32_t s_dup;
    if (threadIdx.x == 0) s_dup = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && n_dup) atomicAdd(&s_dup, n_dup);
    __syncthreads();
    if (threadIdx.x == 0 && s_dup) atomicAdd(&P.counters[CNT_DUPS], s_dup);
}

int launch_flags(const FlagParams &P, cudaStream_t stream, uint64_t *launches) {
    if (P.n == 0) return 0;
    const uint64_t per_cta = (uint64_t) FL_THREADS * FL_ITEMS;
    flags_kernel<<<(uint32_t) ((P.n + per_cta - 1) / per_cta), FL_THREADS, 0, stream>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Compaction for pull(): keep[i] = !(remove && flag has 0x400) (:456: anything flagged after the
// rewrite is dropped, including secondary records that arrived flagged).
// Two-level scan of (kept records, kept bytes): block sums -> scan of block sums -> final.
constexpr int CP_THREADS = 256;
constexpr int CP_ITEMS = 8;
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;

struct U2 {
    uint64_t cnt, bytes;
};

__device__ __forceinline__ U2 block_scan_excl(U2 v, U2 *total, U2 *s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    U2 x = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t a = __shfl_up_sync(0xFFFFFFFFu, x.cnt, o), b = __shfl_up_sync(0xFFFFFFFFu, x.bytes, o);
        if (lane >= o) { x.cnt += a; x.bytes += b; }
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    U2 base = {0, 0}, tot = {0, 0};
    for (int w = 0; w < CP_THREADS / 32; w++) {
        if (w < warp) { base.cnt += s_w[w].cnt; base.bytes += s_w[w].bytes; }
        tot.cnt += s_w[w].cnt; tot.bytes += s_w[w].bytes;
    }
    *total = tot;
    U2 r = {base.cnt + x.cnt - v.cnt, base.bytes + x.bytes - v.bytes};
    __syncthreads();
    return r;
}

__device__ __forceinline__ bool kept(const uint16_t *flag_out, uint64_t i, int remove_dups) {
    return !(remove_dups && (flag_out[i] & 0x400));
}

__global__ void __launch_bounds__(CP_THREADS) compact_sums(const uint64_t *off, uint64_t n, const uint16_t *flag_out,
                                                           int remove_dups, U2 *block_sums) {
    __shared__ U2 s_w[CP_THREADS / 32];
    U2 v = {0, 0};
    uint64_t i0 = (uint64_t) blockIdx.x * CP_TILE + (uint64_t) threadIdx.x * CP_ITEMS;
    for (int k = 0; k < CP_ITEMS; k++) {
        uint64_t i = i0 + k;
        if (i < n && kept(flag_out, i, remove_dups)) { v.cnt++; v.bytes += off[i + 1] - off[i]; }
    }
    U2 tot;
    block_scan_excl(v, &tot, s_w);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(CP_THREADS) compact_scan_blocks(U2 *block_sums, uint32_t n_blocks, uint32_t *counters,
                                                                  uint64_t *totals) {
    __shared__ U2 s_w[CP_THREADS / 32];
    U2 carry = {0, 0};
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += CP_THREADS) {
        uint32_t b = b0 + threadIdx.x;
        U2 v = {0, 0};
        if (b < n_blocks) v = block_sums[b];
        U2 tot;
        U2 ex = block_scan_excl(v, &tot, s_w);
        if (b < n_blocks) block_sums[b] = U2{carry.cnt + ex.cnt, carry.bytes + ex.bytes};
        carry.cnt += tot.cnt;
        carry.bytes += tot.bytes;
    }
    if (threadIdx.x == 0) {
        totals[0] = carry.cnt;
        totals[1] = carry.bytes;
        counters[CNT_KEPT] = (uint32_t) carry.cnt;
    }
}

// one warp per record copies its bytes to the compacted position
__global__ void __launch_bounds__(CP_THREADS) compact_copy(const uint8_t *rec, const uint64_t *off, uint64_t n,
                                                           const uint16_t *flag_out, int remove_dups, const U2 *block_sums,
                                                           uint8_t *out_rec, uint64_t *out_off, const uint64_t *totals) {
    __shared__ U2 s_w[CP_THREADS / 32];
    __shared__ uint64_t s_src[CP_TILE], s_dst[CP_TILE];
    __shared__ uint32_t s_len[CP_TILE];
    U2 v = {0, 0};
    uint64_t i0 = (uint64_t) blockIdx.x * CP_TILE + (uint64_t) threadIdx.x * CP_ITEMS;
    uint32_t keepmask = 0;
    for (int k = 0; k < CP_ITEMS; k++) {
        uint64_t i = i0 + k;
        if (i < n && kept(flag_out, i, remove_dups)) { v.cnt++; v.bytes += off[i + 1] - off[i]; keepmask |= 1u << k; }
    }
    U2 tot;
    U2 ex = block_scan_excl(v, &tot, s_w);
    U2 base = block_sums[blockIdx.x];
    uint64_t rank = base.cnt + ex.cnt, pos = base.bytes + ex.bytes;
    for (int k = 0; k < CP_ITEMS; k++) {
        int slot = threadIdx.x * CP_ITEMS + k;
        s_len[slot] = 0;
        if (keepmask & (1u << k)) {
            uint64_t i = i0 + k, len = off[i + 1] - off[i];
            s_src[slot] = off[i];
            s_dst[slot] = pos;
            s_len[slot] = (uint32_t) len;
            out_off[rank] = pos;
            rank++;
            pos += len;
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out_off[totals[0]] = totals[1];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int slot = warp; slot < CP_TILE; slot += CP_THREADS / 32) {
        uint32_t len = s_len[slot];
        const uint8_t *src = rec + s_src[slot];
        uint8_t *dst = out_rec + s_dst[slot];
        for (uint32_t b = lane; b < len; b += 32) dst[b] = src[b];
    }
}

size_t compact_scratch_bytes(uint64_t n) { return ((n + CP_TILE - 1) / CP_TILE + 2) * sizeof(U2) + 64; }

int launch_compact(const uint8_t *rec, const uint64_t *off, uint64_t n, const uint16_t *flag_out, int remove_dups,
                   uint8_t *out_rec, uint64_t *out_off, uint64_t *scratch, uint32_t *counters, cudaStream_t stream,
                   uint64_t *launches) {
    uint64_t *totals = scratch;                 // [2]
    U2 *block_sums = reinterpret_cast<U2 *>(scratch + 2);
    if (n == 0) {
        OGE_CUDA_TRY(cudaMemsetAsync(totals, 0, 16, stream));
        OGE_CUDA_TRY(cudaMemsetAsync(out_off, 0, 8, stream));
        return 0;
    }
    uint32_t blocks = (uint32_t) ((n + CP_TILE - 1) / CP_TILE);
    compact_sums<<<blocks, CP_THREADS, 0, stream>>>(off, n, flag_out, remove_dups, block_sums);
    compact_scan_blocks<<<1, CP_THREADS, 0, stream>>>(block_sums, blocks, counters, totals);
    compact_copy<<<blocks, CP_THREADS, 0, stream>>>(rec, off, n, flag_out, remove_dups, block_sums, out_rec, out_off, totals);
    *launches += 3;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
