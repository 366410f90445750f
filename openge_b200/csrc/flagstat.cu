// Flag statistics over the resident records: the counting loop of the reference's
// Statistics module (algorithms/statistics.cpp:77-162, flag predicates util/bamtools/BamAlignment.cpp:444-516)
// as device reductions over what the dedup run left in HBM.
//   flagstat_count    one pass over flag_out (2 B/record): twelve counters
//   flagstat_tiles    the "Sorted:" verdict (:89-101).  The reference's scan is a little state machine:
//                     records with refID == -1 or pos == -1 are skipped; a contig id going down is a
//                     violation; a position going down inside a contig is one too, EXCEPT against the first
//                     record of the contig (last_position is reset to -1 when the contig changes and that
//                     record does not set it).  With p = previous considered record and pp = the one before:
//                         violation(j)  <=>  ref[p] > ref[j]
//                                        or  ref[p] == ref[j] and ref[pp] == ref[p] and pos[p] > pos[j]
//                     which is local, so every 1024-record tile checks its interior and leaves its first two
//                     and last two considered records for
//   flagstat_combine  one thread walking the tile summaries in order to check the seams.
// Not on the timed path: run on demand by oge_gpu_dedup_flagstats.
#include "kernels.cuh"

namespace oge {

constexpr int FS_THREADS = 256;
constexpr int FS_ITEMS = 8;

__global__ void __launch_bounds__(FS_THREADS) flagstat_count_kernel(const uint16_t *__restrict__ flags, uint64_t n,
                                                                    unsigned long long *__restrict__ out) {
    uint32_t c[FS_N_COUNTERS];
#pragma unroll
    for (int k = 0; k < FS_N_COUNTERS; k++) c[k] = 0;
    const uint64_t stride = (uint64_t) gridDim.x * FS_THREADS * FS_ITEMS;
    for (uint64_t i0 = ((uint64_t) blockIdx.x * FS_THREADS + threadIdx.x) * FS_ITEMS; i0 < n; i0 += stride) {
        uint32_t f[FS_ITEMS];
        if (i0 + FS_ITEMS <= n) {
            const uint4 v = *reinterpret_cast<const uint4 *>(flags + i0);
            f[0] = v.x & 0xFFFFu; f[1] = v.x >> 16; f[2] = v.y & 0xFFFFu; f[3] = v.y >> 16;
            f[4] = v.z & 0xFFFFu; f[5] = v.z >> 16; f[6] = v.w & 0xFFFFu; f[7] = v.w >> 16;
        } else {
#pragma unroll
            for (int k = 0; k < FS_ITEMS; k++) f[k] = i0 + k < n ? (uint32_t) flags[i0 + k] : 0xFFFF0000u;   // bit 16+: not a record
        }
#pragma unroll
        for (int k = 0; k < FS_ITEMS; k++) {
            const uint32_t w = f[k];
            if (w >> 16) continue;
            const bool mapped = !(w & 0x4), paired = (w & 0x1) != 0;
            c[FS_READS] += 1;
            c[FS_MAPPED] += mapped;
            c[FS_REVERSE] += (w >> 4) & 1;
            c[FS_FORWARD] += !((w >> 4) & 1);
            c[FS_FAILED_QC] += (w >> 9) & 1;
            c[FS_DUPLICATES] += (w >> 10) & 1;
            c[FS_PAIRED] += paired;
            c[FS_PROPER_PAIR] += paired && (w & 0x2);
            c[FS_BOTH_MAPPED] += paired && mapped && !(w & 0x8);
            c[FS_SINGLETONS] += paired && mapped && (w & 0x8);
            c[FS_FIRST_MATE] += paired && (w & 0x40);
            c[FS_SECOND_MATE] += paired && (w & 0x80);
        }
    }
    __shared__ unsigned long long s[FS_N_COUNTERS];
    if (threadIdx.x < FS_N_COUNTERS) s[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FS_N_COUNTERS; k++) {
        uint32_t v = c[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s[k], (unsigned long long) v);
    }
    __syncthreads();
    if (threadIdx.x < FS_N_COUNTERS && s[threadIdx.x]) atomicAdd(&out[threadIdx.x], s[threadIdx.x]);
}

// ---- sortedness ---------------------------------------------------------------------------------
constexpr int ST_THREADS = 1024;      // one record per thread, one tile per CTA

struct __align__(8) RefPos {
    int32_t ref, pos;
};

__device__ __forceinline__ bool st_violation(RefPos pp, bool have_pp, RefPos p, RefPos j) {
    if (p.ref > j.ref) return true;
    return p.ref == j.ref && have_pp && pp.ref == p.ref && p.pos > j.pos;
}

__device__ __forceinline__ uint32_t ld_u32_bytes(const uint8_t *p) {
    return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
}

__global__ void __launch_bounds__(ST_THREADS) flagstat_tiles_kernel(const uint8_t *__restrict__ rec, const uint64_t *__restrict__ off,
                                                                    uint64_t n, SortTile *__restrict__ tiles) {
    __shared__ RefPos s_rp[ST_THREADS];
    __shared__ int s_prev[ST_THREADS];           // previous considered record inside the tile, or -1
    __shared__ int s_warp_last[ST_THREADS / 32];
    __shared__ int s_bad;
    __shared__ RefPos s_first, s_second;      // the tile's first two considered records: their predecessors lie outside
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint64_t i = (uint64_t) blockIdx.x * ST_THREADS + t;
    RefPos me = {-1, -1};
    if (i < n) {
        const uint8_t *p = rec + off[i] + 4;      // refID, pos (util/bam_deserializer.h:176-178); records are not aligned
        me.ref = (int32_t) ld_u32_bytes(p);
        me.pos = (int32_t) ld_u32_bytes(p + 4);
    }
    const bool valid = i < n && me.ref != -1 && me.pos != -1;      // statistics.cpp:89
    s_rp[t] = me;
    if (t == 0) s_bad = 0;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
    const uint32_t below = m & ((1u << lane) - 1);
    int prev = below ? warp * 32 + (31 - __clz(below)) : -1;
    if (lane == 0) s_warp_last[warp] = m ? warp * 32 + (31 - __clz(m)) : -1;
    __syncthreads();
    if (prev < 0)
        for (int w = warp - 1; w >= 0; w--)
            if (s_warp_last[w] >= 0) { prev = s_warp_last[w]; break; }
    s_prev[t] = prev;
    __syncthreads();
    if (valid && prev < 0) s_first = me;
    if (valid && prev >= 0) {
        const int pp = s_prev[prev];
        const RefPos p = s_rp[prev];
        if (pp < 0) s_second = me;
        // with pp outside the tile only the contig test is decidable here; the seam check covers the rest
        if (st_violation(pp >= 0 ? s_rp[pp] : p, pp >= 0, p, me)) s_bad = 1;
    }
    __syncthreads();
    if (t == 0) {
        SortTile T;
        T.n_valid = 0;
        T.bad = s_bad;
        int last = -1;
        for (int w = ST_THREADS / 32 - 1; w >= 0; w--)
            if (s_warp_last[w] >= 0) { last = s_warp_last[w]; break; }
        T.last_ref = T.last_pos = T.last2_ref = T.last2_pos = T.first_ref = T.first_pos = T.second_ref = T.second_pos = -1;
        if (last >= 0) {
            T.n_valid = 1;
            T.first_ref = s_first.ref;
            T.first_pos = s_first.pos;
            T.last_ref = s_rp[last].ref;
            T.last_pos = s_rp[last].pos;
            const int l2 = s_prev[last];
            if (l2 >= 0) {
                T.n_valid = 2;
                T.last2_ref = s_rp[l2].ref;
                T.last2_pos = s_rp[l2].pos;
                T.second_ref = s_second.ref;
                T.second_pos = s_second.pos;
            }
        }
        tiles[blockIdx.x] = T;
    }
}

__global__ void flagstat_combine_kernel(const SortTile *__restrict__ tiles, uint32_t n_tiles, unsigned long long *__restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    bool bad = false, have1 = false, have2 = false;
    RefPos l1 = {-1, -1}, l2 = {-1, -1};      // last and second-last considered record so far
    for (uint32_t b = 0; b < n_tiles && !bad; b++) {
        const SortTile T = tiles[b];
        if (T.bad) { bad = true; break; }
        if (T.n_valid == 0) continue;
        const RefPos f = {T.first_ref, T.first_pos};
        if (have1 && st_violation(l2, have2, l1, f)) { bad = true; break; }
        if (T.n_valid >= 2) {
            const RefPos s = {T.second_ref, T.second_pos};
            if (st_violation(l1, have1, f, s)) { bad = true; break; }
            l2 = RefPos{T.last2_ref, T.last2_pos};
            l1 = RefPos{T.last_ref, T.last_pos};
            have1 = have2 = true;
        } else {
            l2 = l1;
            have2 = have1;
            l1 = f;
            have1 = true;
        }
    }
    out[FS_SORTED] = bad ? 0ull : 1ull;
}

size_t flagstat_scratch_bytes(uint64_t n) {
    return (FS_N_OUT + 1) * sizeof(unsigned long long) + ((n + ST_THREADS - 1) / ST_THREADS + 1) * sizeof(SortTile);
}

// out_dev: FS_N_OUT u64 (counters, then the sorted verdict); tiles follow it in the same scratch block
int launch_flagstats(const uint8_t *rec, const uint64_t *off, const uint16_t *flags, uint64_t n, void *scratch, int sms,
                     cudaStream_t stream, uint64_t *launches) {
    unsigned long long *out = reinterpret_cast<unsigned long long *>(scratch);
    SortTile *tiles = reinterpret_cast<SortTile *>(out + FS_N_OUT + 1);
    OGE_CUDA_TRY(cudaMemsetAsync(out, 0, (FS_N_OUT + 1) * sizeof(unsigned long long), stream));
    const uint32_t n_tiles = (uint32_t) ((n + ST_THREADS - 1) / ST_THREADS);
    if (n) {
        const uint64_t per_cta = (uint64_t) FS_THREADS * FS_ITEMS;
        const uint64_t want = (n + per_cta - 1) / per_cta;
        const uint32_t grid = (uint32_t) (want < (uint64_t) sms * 8 ? want : (uint64_t) sms * 8);      // grid-stride: a multiple of the SM count
        flagstat_count_kernel<<<grid, FS_THREADS, 0, stream>>>(flags, n, out);
        flagstat_tiles_kernel<<<n_tiles, ST_THREADS, 0, stream>>>(rec, off, n, tiles);
        *launches += 2;
    }
    flagstat_combine_kernel<<<1, 32, 0, stream>>>(tiles, n_tiles, out);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
