// K2: mate join.  Replaces ReadEndsMap + the pairing block of buildSortedReadEndLists
// (reference util/picard_structures.h:82-109, algorithms/mark_duplicates.cpp:209-246).
//
// The reference walks the file once with a string-keyed map: the first sighting of a key
// RG + ":" + name is stored, the second removes it and forms a pair, a third is stored again,
// and so on.  For a key seen k times the sightings therefore pair up (1,2), (3,4), ... in file
// order.  On the device:
//   mate_insert   every map-eligible record adds itself to an open-addressing table slot chosen
//                 by the 64-bit hash of the key bytes: val += (1 << 32) + ordinal
//   mate_resolve  a slot with exactly two arrivals is the ordinary case: mate = sum - self.
//                 The earlier record of the two confirms the match by comparing the read-group
//                 code and the name bytes, builds the pair entry (flip rule :226-243, orientation
//                 :169-178, short score sum :245) and appends it.  Anything else -- more than two
//                 arrivals, or a hash-equal couple that fails the comparison -- is sent to
//   mate_complex  the exact path: those records sorted by (hash, ordinal), one thread per hash
//                 value replays the reference's toggle map with full byte comparison of the keys.
// Random 16-byte slot traffic; mates of a coordinate-sorted file sit a few hundred records
// apart, so the second touch of a slot is an L2 hit.
#include "kernels.cuh"

namespace oge {

constexpr int JOIN_THREADS = 256;

__device__ __forceinline__ uint64_t slot_of(uint64_t h, uint64_t n_slots) { return __umul64hi(h, n_slots); }

__global__ void __launch_bounds__(JOIN_THREADS) mate_insert_kernel(JoinParams P) {
    uint64_t i = (uint64_t) blockIdx.x * JOIN_THREADS + threadIdx.x;
    if (i >= P.n) return;
    uint64_t h = P.hk[i];
    if (!h) return;
    uint64_t s = slot_of(h, P.n_slots);
    while (true) {
        unsigned long long *kp = reinterpret_cast<unsigned long long *>(&P.table[s].key);
        unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(kp);
        if (k == h) break;
        if (k == 0) {
            unsigned long long old = atomicCAS(kp, 0ull, (unsigned long long) h);
            if (old == 0 || old == h) break;
        }
        if (++s == P.n_slots) s = 0;
    }
    atomicAdd(reinterpret_cast<unsigned long long *>(&P.table[s].val), (1ull << 32) + (uint32_t) i);
}

// ---- exact key comparison -----------------------------------------------------------------------
// (shared with the slow path)  RG value location by the same tag walk as endbuild.cu.
__device__ int find_rg_global(const uint8_t *tags, uint32_t n, uint32_t *len);   // defined below

struct KeyView {
    const uint8_t *rg, *name;
    uint32_t rg_len, name_len;
};

__device__ KeyView key_view(const uint8_t *rec, const uint64_t *off, uint64_t i) {
    const uint8_t *p = rec + off[i];
    uint32_t rec_len = (uint32_t) (off[i + 1] - off[i]);
    uint32_t l_name = p[12];
    uint32_t n_cig = (uint32_t) p[16] | ((uint32_t) p[17] << 8);
    uint32_t l_seq = (uint32_t) p[20] | ((uint32_t) p[21] << 8) | ((uint32_t) p[22] << 16) | ((uint32_t) p[23] << 24);
    uint32_t o_tags = 36 + l_name + 4 * n_cig + ((l_seq + 1) >> 1) + l_seq;
    KeyView v;
    v.name = p + 36;
    v.name_len = l_name ? l_name - 1 : 0;
    uint32_t rl;
    int at = find_rg_global(p + o_tags, rec_len - o_tags, &rl);
    v.rg = at >= 0 ? p + o_tags + at : p;
    v.rg_len = at >= 0 ? rl : 0;
    return v;
}

__device__ __forceinline__ uint8_t key_byte(const KeyView &v, uint32_t i) {
    return i < v.rg_len ? v.rg[i] : (i == v.rg_len ? (uint8_t) ':' : v.name[i - v.rg_len - 1]);
}

__device__ bool key_equal(const KeyView &a, const KeyView &b) {
    uint32_t la = a.rg_len + 1 + a.name_len, lb = b.rg_len + 1 + b.name_len;
    if (la != lb) return false;
    for (uint32_t i = 0; i < la; i++)
        if (key_byte(a, i) != key_byte(b, i)) return false;
    return true;
}

__device__ int find_rg_global(const uint8_t *tags, uint32_t n, uint32_t *len) {
    uint32_t parsed = 0;
    *len = 0;
    while (parsed < n) {
        if (n - parsed < 3) return -1;
        uint8_t t0 = tags[parsed], t1 = tags[parsed + 1], type = tags[parsed + 2];
        parsed += 3;
        if (t0 == 'R' && t1 == 'G') {
            uint32_t l = 0;
            while (parsed + l < n && tags[parsed + l]) l++;
            *len = l;
            return (int) parsed;
        }
        switch (type) {
            case 'A': case 'c': case 'C': parsed += 1; break;
            case 's': case 'S': parsed += 2; break;
            case 'f': case 'i': case 'I': parsed += 4; break;
            case 'Z': case 'H':
                while (parsed < n && tags[parsed]) parsed++;
                parsed++;
                break;
            case 'B': {
                if (parsed + 5 > n) return -1;
                uint8_t at = tags[parsed];
                int32_t cnt = (int32_t) ((uint32_t) tags[parsed + 1] | ((uint32_t) tags[parsed + 2] << 8) |
                                         ((uint32_t) tags[parsed + 3] << 16) | ((uint32_t) tags[parsed + 4] << 24));
                parsed += 5;
                long long skip;
                if (at == 'c' || at == 'C') skip = cnt;
                else if (at == 's' || at == 'S') skip = 2ll * cnt;
                else if (at == 'f' || at == 'i' || at == 'I') skip = 4ll * cnt;
                else return -1;
                if (skip < 0 || (long long) parsed + skip > (long long) n) return -1;
                parsed += (uint32_t) skip;
                break;
            }
            default: return -1;
        }
        if (parsed >= n) return -1;
        if (tags[parsed] == 0) return -1;
    }
    return -1;
}

// ---- pair entry ---------------------------------------------------------------------------------
// first = the record seen first in the file (the map's stored ReadEnds), second = the current one.
__device__ __forceinline__ E128 make_pair_entry(const KeyLayout &L, const E128 &first, const E128 &second,
                                                uint32_t *idx1_local, uint32_t *idx2_local, uint64_t idx_base) {
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
    bits_or(e, L.p_coord2, c2);
    bits_or(e, L.p_ref2, r2);
    bits_or(e, L.p_orient, (v1 << 1) | v2);      // getOrientationByte(read1Negative, read2Negative) (:169-178)
    bits_or(e, L.p_coord1, c1);
    bits_or(e, L.p_ref1, r1);
    bits_or(e, L.p_lib, lib);
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

__device__ __forceinline__ E128 complex_entry(uint64_t h, uint32_t ordinal) {
    E128 e;
    e.lo = (h << 32) | ordinal;
    e.hi = h >> 32;
    return e;
}

__global__ void __launch_bounds__(JOIN_THREADS) mate_resolve_kernel(JoinParams P) {
    const uint64_t i = (uint64_t) blockIdx.x * JOIN_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1;
    uint64_t h = i < P.n ? P.hk[i] : 0;

    bool emit = false, cplx_self = false, cplx_mate = false;
    uint32_t mate = 0;
    E128 ent;
    ent.lo = ent.hi = 0;
    uint32_t i1 = 0, i2 = 0;
    if (h) {
        uint64_t s = slot_of(h, P.n_slots);
        while (P.table[s].key != h)
            if (++s == P.n_slots) s = 0;
        uint64_t val = P.table[s].val;
        uint32_t cnt = (uint32_t) (val >> 32);
        if (cnt == 2) {
            mate = (uint32_t) val - (uint32_t) i;
            if (mate > (uint32_t) i) {      // this record came first in the file
                bool same = true;
                if (P.verify_names) {
                    uint32_t ra = P.rgcode[i], rb = P.rgcode[mate];
                    const uint8_t *pa = P.rec + P.off[i], *pb = P.rec + P.off[mate];
                    uint32_t la = pa[12], lb = pb[12];
                    same = ra == rb && ra != RGC_UNKNOWN && la == lb;
                    for (uint32_t j = 0; same && j + 1 < la; j++) same = pa[36 + j] == pb[36 + j];
                }
                if (same) {
                    E128 a = ld_frag(P.frag + i), b = ld_frag(P.frag + mate);
                    ent = make_pair_entry(P.kl, a, b, &i1, &i2, P.idx_base);
                    emit = true;
                } else {
                    cplx_self = cplx_mate = true;
                }
            }
        } else if (cnt > 2) {
            cplx_self = true;
        }
    }

    // ---- warp-aggregated appends
    uint32_t m = __ballot_sync(0xFFFFFFFFu, emit);
    if (m) {
        int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&P.counters[CNT_PAIRS], (uint32_t) __popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (emit) {
            reinterpret_cast<ulonglong2 *>(P.pair)[base + __popc(m & lt)] = make_ulonglong2(ent.lo, ent.hi);
            P.mate_of[i1] = i2;
        }
    }
    uint32_t nc = (cplx_self ? 1u : 0u) + (cplx_mate ? 1u : 0u);
    uint32_t any = __ballot_sync(0xFFFFFFFFu, nc != 0);
    if (any) {
        // exclusive prefix of nc over the warp
        uint32_t x = nc;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        uint32_t total = __shfl_sync(0xFFFFFFFFu, x, 31);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&P.counters[CNT_COMPLEX], total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0) + x - nc;
        if (cplx_self) reinterpret_cast<ulonglong2 *>(P.cplx)[base++] = make_ulonglong2(complex_entry(h, (uint32_t) i).lo, complex_entry(h, (uint32_t) i).hi);
        if (cplx_mate) {
            reinterpret_cast<ulonglong2 *>(P.cplx)[base] = make_ulonglong2(complex_entry(h, mate).lo, complex_entry(h, mate).hi);
            atomicAdd(&P.counters[CNT_HASH_MISMATCH], 1u);
        }
    }
}

// ---- exact slow path ------------------------------------------------------------------------------
// sorted: complex entries ordered by (hash, ordinal).  One thread per distinct hash value
// replays the reference's toggle map over that segment with byte-exact key comparison.
__global__ void __launch_bounds__(JOIN_THREADS) mate_complex_kernel(JoinParams P, const E128 *__restrict__ sorted,
                                                                    uint32_t n_cplx, uint8_t *__restrict__ state) {
    uint32_t j = blockIdx.x * JOIN_THREADS + threadIdx.x;
    if (j >= n_cplx) return;
    E128 e = sorted[j];
    uint64_t h = (e.lo >> 32) | (e.hi << 32);
    if (j > 0) {
        E128 q = sorted[j - 1];
        if (((q.lo >> 32) | (q.hi << 32)) == h) return;      // not a segment head
    }
    atomicAdd(&P.counters[CNT_COMPLEX_SEGS], 1u);
    uint32_t end = j;
    while (end < n_cplx) {
        E128 q = sorted[end];
        if (((q.lo >> 32) | (q.hi << 32)) != h) break;
        state[end] = 0;
        end++;
    }
    for (uint32_t a = j; a < end; a++) {
        uint32_t ra = (uint32_t) sorted[a].lo;
        KeyView ka = key_view(P.rec, P.off, ra);
        int found = -1;
        for (uint32_t b = j; b < a; b++) {
            if (!state[b]) continue;
            uint32_t rb = (uint32_t) sorted[b].lo;
            KeyView kb = key_view(P.rec, P.off, rb);
            if (key_equal(ka, kb)) {
                found = (int) b;
                break;
            }
        }
        if (found < 0) {
            state[a] = 1;       // tmp.put (:222-223)
        } else {
            state[found] = 0;   // tmp.remove (:216)
            uint32_t rb = (uint32_t) sorted[found].lo;
            E128 first = ld_frag(P.frag + rb), second = ld_frag(P.frag + ra);
            uint32_t i1, i2;
            E128 ent = make_pair_entry(P.kl, first, second, &i1, &i2, P.idx_base);
            uint32_t pos = atomicAdd(&P.counters[CNT_PAIRS], 1u);
            reinterpret_cast<ulonglong2 *>(P.pair)[pos] = make_ulonglong2(ent.lo, ent.hi);
            P.mate_of[i1] = i2;
        }
    }
}

int launch_mate_insert(const JoinParams &P, cudaStream_t stream, uint64_t *launches) {
    if (P.n == 0) return 0;
    mate_insert_kernel<<<(uint32_t) ((P.n + JOIN_THREADS - 1) / JOIN_THREADS), JOIN_THREADS, 0, stream>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_mate_resolve(const JoinParams &P, cudaStream_t stream, uint64_t *launches) {
    if (P.n == 0) return 0;
    mate_resolve_kernel<<<(uint32_t) ((P.n + JOIN_THREADS - 1) / JOIN_THREADS), JOIN_THREADS, 0, stream>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_mate_complex(const JoinParams &P, const E128 *sorted_cplx, uint32_t n_cplx, uint8_t *state, cudaStream_t stream,
                        uint64_t *launches) {
    if (n_cplx == 0) return 0;
    mate_complex_kernel<<<(n_cplx + JOIN_THREADS - 1) / JOIN_THREADS, JOIN_THREADS, 0, stream>>>(P, sorted_cplx, n_cplx, state);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
