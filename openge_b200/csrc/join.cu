// K2: mate join.  Replaces ReadEndsMap + the pairing block of buildSortedReadEndLists
// (reference util/picard_structures.h:82-109, algorithms/mark_duplicates.cpp:209-246).
//
// The reference walks the file once with a string-keyed map: the first sighting of a key
// RG + ":" + name is stored, the second removes it and forms a pair, a third is stored again,
// and so on.  For a key seen k times the sightings therefore pair up (1,2), (3,4), ... in file
// order.  On the device, one pass over the records:
//   mate_join     every map-eligible record claims/finds the open-addressing slot of its 64-bit key
//                 hash and adds (1 << 32) + ordinal + 1 to the slot's counter word.  The record that
//                 gets back an arrival count of 1 is the second of its name: the sum field is its
//                 mate's ordinal.  It confirms the match by comparing read-group code and name
//                 bytes, builds the pair entry (flip rule :226-243, orientation :169-178, short
//                 score sum :245), appends it and leaves (first, second, pair position) in the slot.
//                 An arrival count of 2 or more means the name is not a plain pair: the record goes
//                 to the exact path, and the third arrival also lists the slot, so that
//   mate_fixup    retracts the provisional pair of such a slot and sends its two records after the others;
//   mate_complex  the exact path: those records sorted by (hash, ordinal), one thread per hash
//                 value replays the reference's toggle map with full byte comparison of the keys.
//                 Hash-equal couples whose names differ take the same path.
// One random 32-byte slot per record (two atomics); mates of a coordinate-sorted file sit a few
// hundred records apart, so the second touch of a slot and the mate's name are L2 hits.
#include <stdlib.h>

#include "kernels.cuh"
#include "pairing.cuh"

namespace oge {

constexpr int JOIN_THREADS = 256;

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

// name bytes from NAME_TAG_BYTES on (the part the tags do not cover) of records a and b equal?
// l_name is known equal.  Reads past the name stay inside the record buffer (cigar/bases/quals follow).
__device__ bool name_tails_equal(const uint8_t *pa, const uint8_t *pb, uint32_t l_name) {
    const uint32_t n = l_name ? l_name - 1 : 0;
    uint32_t j = NAME_TAG_BYTES;
    for (; j + 4 <= n; j += 4)
        if (ldg_u32_unaligned(pa + 36 + j) != ldg_u32_unaligned(pb + 36 + j)) return false;
    for (; j < n; j++)
        if (pa[36 + j] != pb[36 + j]) return false;
    return true;
}

// ITEMS records per thread, staged so that the dependent chain hash -> slot -> counter -> tags -> ends
// of one record overlaps with the others': the kernel is bound by memory latency, not bandwidth.
// LIST: the records to join are P.list[0 .. n_list) (what the fused end-build could not settle inside its CTAs).
template <int ITEMS, bool LIST>
__global__ void __launch_bounds__(JOIN_THREADS, ITEMS == 1 ? 8 : 4) mate_join_kernel(JoinParams P) {
    const uint64_t i0 = (uint64_t) blockIdx.x * (JOIN_THREADS * ITEMS) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1;

    uint64_t h[ITEMS], s[ITEMS];
    uint32_t rec_i[ITEMS];
    unsigned long long key0[ITEMS], old[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint64_t j = i0 + (uint64_t) k * JOIN_THREADS;
        const bool valid = j < (LIST ? (uint64_t) P.n_list : P.n);
        rec_i[k] = valid ? (LIST ? P.list[j] : (uint32_t) j) : 0u;
        h[k] = valid ? P.hk[rec_i[k]] : 0;
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {      // first probe of every record in flight together
        s[k] = slot_of(h[k], P.n_slots);
        key0[k] = 0;
        if (h[k]) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(key0[k]) : "l"(&P.table[s[k]].key) : "memory");
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {      // claim or find the key's slot and count this arrival
        old[k] = 0;
        if (!h[k]) continue;
        const unsigned long long inc = (1ull << 32) + rec_i[k] + 1u;
        unsigned long long key = key0[k];
        while (true) {
            if (key == 0) {
                // an empty slot: claim it and count the arrival with ONE 16-byte compare-and-swap of (key, val);
                // the L2 atomic units are what bounds this kernel, so the first arrival of a name costs one
                // atomic instead of a claim plus an add
                unsigned long long ok, ov;
                asm volatile(
                    "{\n .reg .b128 cmp, nv, ov;\n mov.b128 cmp, {%2, %3};\n mov.b128 nv, {%4, %5};\n"
                    " atom.global.relaxed.gpu.cas.b128 ov, [%6], cmp, nv;\n mov.b128 {%0, %1}, ov;\n}\n"
                    : "=l"(ok), "=l"(ov)
                    : "l"(0ull), "l"(0ull), "l"((unsigned long long) h[k]), "l"(inc), "l"(&P.table[s[k]])
                    : "memory");
                if (ok == 0 && ov == 0) break;      // claimed: first arrival, old[k] stays 0
                key = ok;                           // somebody was faster: their key is in the slot now
            }
            if (key == h[k]) {
                old[k] = atomicAdd(reinterpret_cast<unsigned long long *>(&P.table[s[k]].val), inc);
                break;
            }
            if (++s[k] == P.n_slots) s[k] = 0;
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(key) : "l"(&P.table[s[k]].key) : "memory");
        }
    }

#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint64_t i = rec_i[k];
        bool emit = false, far = false, cplx_self = false, cplx_other = false, list_slot = false;
        uint32_t other = 0, i1 = 0, i2 = 0;
        E128 ent;
        ent.lo = ent.hi = 0;
        if (h[k]) {
            const uint32_t arrivals = (uint32_t) (old[k] >> 32);
            if (arrivals == 1) {      // the second of its name: the sum field is the first one's ordinal
                other = (uint32_t) old[k] - 1u;
                bool same = true;
                if (P.verify_names) {
                    const uint4 *ta = reinterpret_cast<const uint4 *>(P.tag + i), *tb = reinterpret_cast<const uint4 *>(P.tag + other);
                    const uint4 a0 = ta[0], a1 = ta[1], b0 = tb[0], b1 = tb[1];
                    same = a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w && a1.x == b1.x && a1.y == b1.y &&
                           a1.z == b1.z && a1.w == b1.w && (a0.x & 0xFFFFu) != RGC_UNKNOWN;
                    const uint32_t l_name = (a0.x >> 16) & 0xFFu;
                    if (same && l_name > NAME_TAG_BYTES + 1) same = name_tails_equal(P.rec + P.off[i], P.rec + P.off[other], l_name);
                }
                if (same) {
                    const uint32_t first = min((uint32_t) i, other), second = max((uint32_t) i, other);      // file order
                    ent = make_pair_entry(P.kl, ld_frag(P.frag + first), ld_frag(P.frag + second), &i1, &i2, P.idx_base, &far);
                    emit = true;
                } else {
                    cplx_self = cplx_other = true;      // two names, one hash (or read groups the header does not list)
                }
            } else if (arrivals >= 2) {
                cplx_self = true;
                list_slot = arrivals == 2;
            }
        }

        // ---- warp-aggregated appends (near pairs; far pairs are rare and append one by one)
        uint32_t m = __ballot_sync(0xFFFFFFFFu, emit && !far);
        uint32_t pair_pos = SLOT_NO_PAIR;
        if (m) {
            int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&P.counters[CNT_PAIRS], (uint32_t) __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (emit && !far) {
                pair_pos = base + __popc(m & lt);
                reinterpret_cast<ulonglong2 *>(P.pair)[pair_pos] = make_ulonglong2(ent.lo, ent.hi);
            }
        }
        if (emit) {
            if (far) {
                const uint32_t at = atomicAdd(&P.counters[CNT_PAIRS_FAR], 1u);
                reinterpret_cast<ulonglong2 *>(P.pair_far)[at] = make_ulonglong2(ent.lo, ent.hi);
                pair_pos = at | SLOT_PAIR_FAR;
            }
            P.mate_of[i1] = (uint32_t) (i2 + P.idx_base);
        }
        if (emit || cplx_other) {      // what mate_fixup needs should a third record of this name turn up
            P.table[s[k]].who = ((uint64_t) (uint32_t) i << 32) | other;
            P.table[s[k]].pair_pos = pair_pos;
        }
        uint32_t nc = (cplx_self ? 1u : 0u) + (cplx_other ? 1u : 0u);
        uint32_t any = __ballot_sync(0xFFFFFFFFu, nc != 0);
        if (any) {
            uint32_t x = nc;      // exclusive prefix of nc over the warp
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += y;
            }
            uint32_t total = __shfl_sync(0xFFFFFFFFu, x, 31);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&P.counters[CNT_COMPLEX], total);
            base = __shfl_sync(0xFFFFFFFFu, base, 0) + x - nc;
            if (cplx_self) {
                E128 c = complex_entry(h[k], (uint32_t) i);
                reinterpret_cast<ulonglong2 *>(P.cplx)[base++] = make_ulonglong2(c.lo, c.hi);
            }
            if (cplx_other) {
                E128 c = complex_entry(h[k], other);
                reinterpret_cast<ulonglong2 *>(P.cplx)[base] = make_ulonglong2(c.lo, c.hi);
                atomicAdd(&P.counters[CNT_HASH_MISMATCH], 1u);
            }
        }
        if (list_slot) P.cplx_slots[atomicAdd(&P.counters[CNT_COMPLEX_SLOTS], 1u)] = (uint32_t) s[k];
    }
}

// One thread per slot that saw a third arrival: its first two records follow the others to the exact
// path, and the pair they formed provisionally is retracted (an all-ones entry sorts behind every
// real key; the host shortens the pair list by the number of retractions).
__global__ void __launch_bounds__(JOIN_THREADS) mate_fixup_kernel(JoinParams P, uint32_t n_slots_listed) {
    uint32_t j = blockIdx.x * JOIN_THREADS + threadIdx.x;
    if (j >= n_slots_listed) return;
    const MateSlot &sl = P.table[P.cplx_slots[j]];
    const uint32_t a = (uint32_t) (sl.who >> 32), b = (uint32_t) sl.who;
    if (sl.pair_pos != SLOT_NO_PAIR) {      // else: a hash-mismatched couple, already on the exact path
        const bool far = (sl.pair_pos & SLOT_PAIR_FAR) != 0;
        reinterpret_cast<ulonglong2 *>(far ? P.pair_far : P.pair)[sl.pair_pos & ~SLOT_PAIR_FAR] = make_ulonglong2(~0ull, ~0ull);
        atomicAdd(&P.counters[far ? CNT_FAR_RETRACTED : CNT_PAIRS_RETRACTED], 1u);
        uint32_t base = atomicAdd(&P.counters[CNT_COMPLEX], 2u);
        E128 ca = complex_entry(sl.key, a), cb = complex_entry(sl.key, b);
        reinterpret_cast<ulonglong2 *>(P.cplx)[base] = make_ulonglong2(ca.lo, ca.hi);
        reinterpret_cast<ulonglong2 *>(P.cplx)[base + 1] = make_ulonglong2(cb.lo, cb.hi);
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
            bool far;
            E128 ent = make_pair_entry(P.kl, first, second, &i1, &i2, P.idx_base, &far);
            uint32_t pos = atomicAdd(&P.counters[far ? CNT_PAIRS_FAR : CNT_PAIRS], 1u);
            reinterpret_cast<ulonglong2 *>(far ? P.pair_far : P.pair)[pos] = make_ulonglong2(ent.lo, ent.hi);
            P.mate_of[i1] = (uint32_t) (i2 + P.idx_base);
        }
    }
}

int launch_mate_join(const JoinParams &P, cudaStream_t stream, uint64_t *launches) {
    const uint64_t count = P.list ? (uint64_t) P.n_list : P.n;
    if (count == 0) return 0;
    // one record per thread: two and four were measured and do not help
    const uint32_t grid = (uint32_t) ((count + JOIN_THREADS - 1) / JOIN_THREADS);
    if (P.list) mate_join_kernel<1, true><<<grid, JOIN_THREADS, 0, stream>>>(P);
    else mate_join_kernel<1, false><<<grid, JOIN_THREADS, 0, stream>>>(P);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- fused form: the check pass -----------------------------------------------------------------
// A pair formed inside a CTA of the fused end-build is final iff no other record of its key exists
// among the leftovers, i.e. iff its hash is not in the mate table the global join has just built.
// A hit retracts the pair (all-ones entry: sorts behind every key; counted) and sends its two records
// to the exact path, together with whatever the slot held: a lone record (arrivals == 1: pushed here),
// a couple (arrivals == 2: the slot is listed for mate_fixup, which retracts their pair and pushes
// them), or a name the join itself found complex (arrivals >= 3: already there).  Every hit adds two
// to the slot's arrival count, so that exactly one visitor does that.
__global__ void __launch_bounds__(JOIN_THREADS) pair_check_kernel(JoinParams P, const uint64_t *__restrict__ pair_hk, uint32_t n_pairs,
                                                                  int far) {
    const uint32_t pos = blockIdx.x * JOIN_THREADS + threadIdx.x;
    if (pos >= n_pairs) return;
    const uint64_t h = pair_hk[pos];
    if (!h) return;
    uint64_t s = slot_of(h, P.n_slots);
    while (true) {
        unsigned long long key;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(key) : "l"(&P.table[s].key) : "memory");
        if (key == 0) return;      // no leftover of this name: the pair stands
        if (key == h) break;
        if (++s == P.n_slots) s = 0;
    }
    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&P.table[s].val), 2ull << 32);
    const uint32_t arrivals = (uint32_t) (old >> 32);
    E128 *list = far ? P.pair_far : P.pair;
    const E128 ent = ld_frag(list + pos);
    const uint32_t i1 = (uint32_t) (bits_get(ent, P.kl.p_idx, P.kl.idx_bits) - P.idx_base);
    const uint32_t i2 = (uint32_t) (P.mate_of[i1] - P.idx_base);
    reinterpret_cast<ulonglong2 *>(list)[pos] = make_ulonglong2(~0ull, ~0ull);
    atomicAdd(&P.counters[far ? CNT_FAR_RETRACTED : CNT_PAIRS_RETRACTED], 1u);
    const uint32_t extra = arrivals == 1 ? 1u : 0u;
    uint32_t base = atomicAdd(&P.counters[CNT_COMPLEX], 2u + extra);
    E128 c = complex_entry(h, i1);
    reinterpret_cast<ulonglong2 *>(P.cplx)[base++] = make_ulonglong2(c.lo, c.hi);
    c = complex_entry(h, i2);
    reinterpret_cast<ulonglong2 *>(P.cplx)[base++] = make_ulonglong2(c.lo, c.hi);
    if (extra) {
        c = complex_entry(h, (uint32_t) old - 1u);
        reinterpret_cast<ulonglong2 *>(P.cplx)[base] = make_ulonglong2(c.lo, c.hi);
    }
    if (arrivals == 2) P.cplx_slots[atomicAdd(&P.counters[CNT_COMPLEX_SLOTS], 1u)] = (uint32_t) s;
}

int launch_pair_check(const JoinParams &P, const uint64_t *pair_hk, uint32_t n_pairs, bool far, cudaStream_t stream, uint64_t *launches) {
    if (n_pairs == 0) return 0;
    pair_check_kernel<<<(n_pairs + JOIN_THREADS - 1) / JOIN_THREADS, JOIN_THREADS, 0, stream>>>(P, pair_hk, n_pairs, far ? 1 : 0);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_mate_fixup(const JoinParams &P, uint32_t n_slots_listed, cudaStream_t stream, uint64_t *launches) {
    if (n_slots_listed == 0) return 0;
    mate_fixup_kernel<<<(n_slots_listed + JOIN_THREADS - 1) / JOIN_THREADS, JOIN_THREADS, 0, stream>>>(P, n_slots_listed);
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
