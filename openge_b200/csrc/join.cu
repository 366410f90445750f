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

#include "ctx.cuh"
#include "keyview.cuh"
#include "pairing.cuh"

namespace oge {

constexpr int JOIN_THREADS = 256;

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
// sorted: complex entries ordered by (hash, ordinal).  The entries of one hash value (a segment) replay the
// reference's toggle map (tmp.put / tmp.remove, mark_duplicates.cpp:216-223) with byte-exact key comparison.  The map
// holds at most ONE unmatched sighting per distinct key, and the keys of a segment share a 64-bit hash, so the live
// set of a segment is a handful of entries at most (one, unless hashes collide): the replay keeps it in registers
// and costs O(k) comparisons for k sightings (the byte map `state` is the fall-back should more than CPLX_LIVE
// distinct keys ever share a hash).  A short segment is replayed by one thread; a long one (stripped or constant read
// names put a whole file into one segment) by a CTA: if all its keys are equal -- adjacent comparisons in parallel --
// the sightings pair up (1,2), (3,4), ... in file order and the pairs are emitted in parallel.
constexpr int CPLX_LIVE = 8;
constexpr uint32_t CPLX_LONG = 512;

__device__ __forceinline__ uint64_t cplx_hash(const E128 &e) { return (e.lo >> 32) | (e.hi << 32); }

__device__ __forceinline__ void cplx_emit(const JoinParams &P, uint32_t r_first, uint32_t r_second) {
    const E128 first = ld_frag(P.frag + r_first), second = ld_frag(P.frag + r_second);
    uint32_t i1, i2;
    bool far;
    const E128 ent = make_pair_entry(P.kl, first, second, &i1, &i2, P.idx_base, &far);
    const uint32_t pos = atomicAdd(&P.counters[far ? CNT_PAIRS_FAR : CNT_PAIRS], 1u);
    reinterpret_cast<ulonglong2 *>(far ? P.pair_far : P.pair)[pos] = make_ulonglong2(ent.lo, ent.hi);
    P.mate_of[i1] = (uint32_t) (i2 + P.idx_base);
}

// the toggle over sorted[j, end), sequentially
__device__ void cplx_replay_segment(const JoinParams &P, const E128 *__restrict__ sorted, uint32_t j, uint32_t end, uint8_t *__restrict__ state) {
    uint32_t live[CPLX_LIVE];
    int n_live = 0;
    bool overflow = false;
    for (uint32_t a = j; a < end; a++) {
        const uint32_t ra = (uint32_t) sorted[a].lo;
        const KeyView ka = key_view(P.rec, P.off, ra);
        int found = -1;
        if (!overflow) {
            for (int t = 0; t < n_live; t++)
                if (key_equal(ka, key_view(P.rec, P.off, (uint32_t) sorted[live[t]].lo))) { found = (int) live[t]; live[t] = live[--n_live]; break; }
        } else {
            for (uint32_t b2 = j; b2 < a; b2++)
                if (state[b2] && key_equal(ka, key_view(P.rec, P.off, (uint32_t) sorted[b2].lo))) { found = (int) b2; break; }
        }
        if (found < 0) {      // tmp.put (:222-223)
            if (overflow) state[a] = 1;
            else if (n_live < CPLX_LIVE) live[n_live++] = a;
            else {      // more distinct keys under one hash than the registers hold: from here on the byte map is the live set
                overflow = true;
                for (uint32_t x = j; x < a; x++) state[x] = 0;
                for (int t = 0; t < n_live; t++) state[live[t]] = 1;
                state[a] = 1;
            }
        } else {              // tmp.remove (:216)
            if (overflow) state[found] = 0;
            cplx_emit(P, (uint32_t) sorted[found].lo, ra);
        }
    }
}

__global__ void __launch_bounds__(JOIN_THREADS) mate_complex_kernel(JoinParams P, const E128 *__restrict__ sorted,
                                                                    uint32_t n_cplx, uint8_t *__restrict__ state, uint2 *__restrict__ long_segs) {
    uint32_t j = blockIdx.x * JOIN_THREADS + threadIdx.x;
    if (j >= n_cplx) return;
    const uint64_t h = cplx_hash(sorted[j]);
    if (j > 0 && cplx_hash(sorted[j - 1]) == h) return;      // not a segment head
    atomicAdd(&P.counters[CNT_COMPLEX_SEGS], 1u);
    // end of the segment: first entry with another hash (binary search: the list is sorted by hash)
    uint32_t lo = j + 1, hi = n_cplx;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (cplx_hash(sorted[mid]) == h) lo = mid + 1; else hi = mid;
    }
    const uint32_t end = lo;
    if (end - j > CPLX_LONG) {
        long_segs[atomicAdd(&P.counters[CNT_SCRATCH0], 1u)] = make_uint2(j, end);
        return;
    }
    cplx_replay_segment(P, sorted, j, end, state);
}

// one CTA per long segment
__global__ void __launch_bounds__(JOIN_THREADS) mate_complex_long_kernel(JoinParams P, const E128 *__restrict__ sorted, uint8_t *__restrict__ state,
                                                                         const uint2 *__restrict__ long_segs) {
    const uint2 seg = long_segs[blockIdx.x];
    const uint32_t j = seg.x, end = seg.y;
    __shared__ int s_differ;
    if (threadIdx.x == 0) s_differ = 0;
    __syncthreads();
    for (uint32_t a = j + 1 + threadIdx.x; a < end && !s_differ; a += JOIN_THREADS)
        if (!key_equal(key_view(P.rec, P.off, (uint32_t) sorted[a].lo), key_view(P.rec, P.off, (uint32_t) sorted[a - 1].lo))) s_differ = 1;
    __syncthreads();
    if (s_differ) {      // several keys under one hash in a long segment: the sequential toggle
        if (threadIdx.x == 0) cplx_replay_segment(P, sorted, j, end, state);
        return;
    }
    const uint32_t n_pairs = (end - j) / 2;      // one key: sightings (1,2), (3,4), ... in file order; an odd last one stays unmatched
    for (uint32_t t = threadIdx.x; t < n_pairs; t += JOIN_THREADS) cplx_emit(P, (uint32_t) sorted[j + 2 * t].lo, (uint32_t) sorted[j + 2 * t + 1].lo);
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


// ---- the windowed join -----------------------------------------------------------------------------
// A coordinate-sorted file keeps the two reads of a pair a few (to a few hundred) records apart, so the join does
// not need a table over the whole file: every CTA walks a CONTIGUOUS range of records in tiles and keeps the
// reads whose mate it has not met yet in a shared-memory table (8-way buckets: a key word per way = low hash word,
// 12 bytes of payload).  Per tile:
//   phase 1   every map-eligible record scans its bucket: its key word is there -> FOUND that way; else it
//             claims the first empty way with a shared-memory CAS (a failed CAS that returns its own key
//             word is a FOUND as well: the mate got there first) and writes its payload -> INSERTED
//   phase 2   a FOUND record takes the entry (first taker only) and confirms the match by comparing the two
//             32-byte name tags (and the name tails in the records for names longer than a tag); it builds
//             the pair entry exactly as the global join does (flip rule mark_duplicates.cpp:226-243,
//             orientation :169-178, score :245), appends it (warp-aggregated) and frees the way.
// Everything else -- bucket full, a second taker (name seen three times at once), equal hash but different key
// bytes, entries nobody came for within LJ_HORIZON records or by the end of the CTA's range -- goes on the
// `left` list, which the global join above takes as its input (LIST form).  Pairing here is by arrival, not by
// file order, which is only right for names seen exactly twice; that is what pair_check_kernel establishes
// afterwards (a name with any record on the `left` list has its pairs retracted and is replayed in file order by
// the exact path).  Within one CTA an entry is always the most recent unmatched sighting of its name, so names
// without leftovers and without a double take were paired as the reference's toggle map pairs them
// (util/picard_structures.h:87-96).
// HBM traffic: the hash of every record once (8 B), tag and end entry of every paired record once (48 B), the
// pair entry and its hash (24 B per pair): about 70 B per record, streamed; the table never leaves the SM.
struct __align__(4) LjPay {
    uint32_t h_hi, ord, cnt;      // high hash word, local ordinal, takers so far
};
static_assert(sizeof(LjPay) + 4 == LJ_ENTRY_BYTES, "LJ_ENTRY_BYTES");

__device__ __forceinline__ void lj_leftover(uint32_t *counter, uint32_t *left, uint32_t ord) {
    const uint32_t act = __activemask(), lane = threadIdx.x & 31;
    const int leader = __ffs(act) - 1;
    uint32_t base = 0;
    if ((int) lane == leader) base = atomicAdd(counter, (uint32_t) __popc(act));
    base = __shfl_sync(act, base, leader);
    left[base + __popc(act & ((1u << lane) - 1))] = ord;
}

// K2a: phases 1 and 2 (take) over the hashes alone.  A taker writes the couple (its ordinal, the entry's ordinal,
// the key hash) into the CTA's own stretch of the couple list -- both records lie in the CTA's range, so the
// stretch never holds more than half the range's records and needs no cross-CTA counter.  Nothing but hk[] is read
// from global memory: one latency per tile, small register footprint, six CTAs per SM.
__global__ void __launch_bounds__(LJ_THREADS, LJ_CTAS_PER_SM) local_match_kernel(const __grid_constant__ JoinParams P,
                                                                                 const __grid_constant__ LocalJoinParams J) {
    extern __shared__ __align__(16) uint8_t lj_smem[];      // [key words: n_entries][payloads: n_entries]
    __shared__ uint32_t s_couples;
    const int tid = threadIdx.x;
    const uint32_t n_entries = J.n_buckets * LJ_WAYS;
    uint32_t *keys = reinterpret_cast<uint32_t *>(lj_smem);
    LjPay *pay = reinterpret_cast<LjPay *>(keys + n_entries);
    for (uint32_t j = tid; j < n_entries; j += LJ_THREADS) keys[j] = 0;
    if (tid == 0) s_couples = 0;
    const uint64_t t0 = (uint64_t) blockIdx.x * J.tiles_per_cta;
    const uint64_t n_tiles = (P.n + LJ_TILE - 1) / LJ_TILE;
    const uint64_t t1 = t0 + J.tiles_per_cta < n_tiles ? t0 + J.tiles_per_cta : n_tiles;
    uint4 *couples = J.couples + (uint64_t) blockIdx.x * J.couples_per_cta;
    __syncthreads();

    uint32_t k = 0;
    uint64_t h_next[LJ_ITEMS];      // the next tile's hashes are requested one tile ahead
#pragma unroll
    for (int q = 0; q < LJ_ITEMS; q++) {
        const uint64_t i = t0 * LJ_TILE + (uint64_t) q * LJ_THREADS + tid;
        h_next[q] = t0 < t1 && i < P.n ? P.hk[i] : 0ull;
    }
    for (uint64_t tile = t0; tile < t1; tile++, k++) {
        const uint64_t r0 = tile * LJ_TILE;
        uint64_t h[LJ_ITEMS];
        int way[LJ_ITEMS];            // FOUND: the way; INSERTED: -2; overflow: -3; not in the map: -1
        uint32_t slot0[LJ_ITEMS];
#pragma unroll
        for (int q = 0; q < LJ_ITEMS; q++) {
            h[q] = h_next[q];
            const uint64_t i = r0 + LJ_TILE + (uint64_t) q * LJ_THREADS + tid;
            h_next[q] = tile + 1 < t1 && i < P.n ? P.hk[i] : 0ull;
        }
        // ---- phase 1: find the mate's entry, or leave one
#pragma unroll
        for (int q = 0; q < LJ_ITEMS; q++) {
            way[q] = -1;
            slot0[q] = 0;
            if (!h[q]) continue;
            const uint32_t k32 = (uint32_t) h[q], h_hi = (uint32_t) (h[q] >> 32);
            const uint32_t ord = (uint32_t) (r0 + (uint64_t) q * LJ_THREADS + tid);
            slot0[q] = __umulhi(h_hi, J.n_buckets) * LJ_WAYS;
            const uint4 ka = *reinterpret_cast<const uint4 *>(keys + slot0[q]), kb = *reinterpret_cast<const uint4 *>(keys + slot0[q] + 4);
            const uint32_t kk[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
            int first_empty = -1, w_found = -1;
#pragma unroll
            for (int w = 7; w >= 0; w--) {
                if (kk[w] == k32) w_found = w;
                if (kk[w] == 0) first_empty = w;
            }
            if (w_found >= 0) { way[q] = w_found; continue; }
            way[q] = -3;
            for (int w = first_empty; w >= 0 && w < LJ_WAYS; w++) {
                const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(keys + slot0[q] + w);
                if (cur == k32) { way[q] = w; break; }
                if (cur != 0) continue;
                const uint32_t old = atomicCAS(keys + slot0[q] + w, 0u, k32);
                if (old == 0) {
                    LjPay &e = pay[slot0[q] + w];
                    e.h_hi = h_hi; e.ord = ord; e.cnt = 0;
                    way[q] = -2;
                    break;
                }
                if (old == k32) { way[q] = w; break; }
            }
        }
        __syncthreads();      // payloads of this tile's insertions are visible

        // ---- phase 2: take the entry (first taker only); the names are compared by K2b
#pragma unroll
        for (int q = 0; q < LJ_ITEMS; q++) {
            if (!h[q] || way[q] == -2) continue;
            const uint32_t i = (uint32_t) (r0 + (uint64_t) q * LJ_THREADS + tid);
            bool took = false;
            if (way[q] >= 0) {
                LjPay &e = pay[slot0[q] + way[q]];
                if (e.h_hi == (uint32_t) (h[q] >> 32) && atomicAdd(&e.cnt, 1u) == 0) {
                    const uint32_t other = e.ord;
                    keys[slot0[q] + way[q]] = 0;      // the way is free again (nobody reads key words before the next barrier)
                    couples[atomicAdd(&s_couples, 1u)] = make_uint4(i, other, (uint32_t) h[q], (uint32_t) (h[q] >> 32));
                    took = true;
                }
            }
            if (!took) lj_leftover(P.counters + CNT_LEFT, J.left, i);
        }

        // ---- entries nobody came for leave for the global join
        if ((k % LJ_SWEEP_TILES) == LJ_SWEEP_TILES - 1) {
            __syncthreads();
            for (uint32_t j = tid; j < n_entries; j += LJ_THREADS) {
                if (keys[j] && (uint32_t) r0 - pay[j].ord > LJ_HORIZON) {
                    keys[j] = 0;
                    lj_leftover(P.counters + CNT_LEFT, J.left, pay[j].ord);
                }
            }
        }
        __syncthreads();      // phase 2 (and the sweep) are complete before the next tile's phase 1
    }
    // ---- end of the CTA's range: whatever is still waiting leaves
    for (uint32_t j = tid; j < n_entries; j += LJ_THREADS)
        if (keys[j]) lj_leftover(P.counters + CNT_LEFT, J.left, pay[j].ord);
    if (tid == 0) J.couple_count[blockIdx.x] = s_couples;
}

// K2b: one thread per couple.  Confirms the match by comparing the two 32-byte name tags (and the name tails in
// the records for names longer than a tag), builds the pair entry and appends it (warp-aggregated); a couple that
// fails the comparison -- one hash, two keys, or read groups the header does not list -- goes to the global join.
__global__ void __launch_bounds__(JOIN_THREADS) local_emit_kernel(const __grid_constant__ JoinParams P, const __grid_constant__ LocalJoinParams J,
                                                                  uint32_t n_ctas) {
    const uint64_t j = (uint64_t) blockIdx.x * JOIN_THREADS + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    const uint32_t cta = (uint32_t) (j / J.couples_per_cta);
    const bool live = cta < n_ctas && (uint32_t) (j - (uint64_t) cta * J.couples_per_cta) < J.couple_count[cta];
    bool emit = false, far = false;
    uint32_t i1 = 0, i2 = 0;
    uint64_t h = 0;
    E128 ent;
    ent.lo = ent.hi = 0;
    if (live) {
        const uint4 c = J.couples[j];
        const uint32_t i = c.x, other = c.y;
        h = ((uint64_t) c.w << 32) | c.z;
        const uint4 *ta = reinterpret_cast<const uint4 *>(P.tag + i), *tb = reinterpret_cast<const uint4 *>(P.tag + other);
        const uint4 a0 = ta[0], a1 = ta[1], b0 = tb[0], b1 = tb[1];
        const E128 fa = ld_frag(P.frag + i), fb = ld_frag(P.frag + other);
        bool same = a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w && a1.x == b1.x && a1.y == b1.y &&
                    a1.z == b1.z && a1.w == b1.w && (a0.x & 0xFFFFu) != RGC_UNKNOWN;
        const uint32_t l_name = (a0.x >> 16) & 0xFFu;
        if (same && l_name > NAME_TAG_BYTES + 1) same = name_tails_equal(P.rec + P.off[i], P.rec + P.off[other], l_name);
        if (same) {
            const bool self_first = i < other;      // file order
            ent = make_pair_entry(P.kl, self_first ? fa : fb, self_first ? fb : fa, &i1, &i2, P.idx_base, &far);
            emit = true;
        } else {
            lj_leftover(P.counters + CNT_LEFT, J.left, i);
            lj_leftover(P.counters + CNT_LEFT, J.left, other);
        }
    }
    // ---- CTA-aggregated append: one reservation in the near-pair list per CTA (a counter at one address serves about
    //      one atomic per nanosecond; per warp that was a third of this kernel's time).  Far pairs are rare and append one by one.
    __shared__ uint32_t s_warp_cnt[JOIN_THREADS / 32], s_base;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, emit && !far);
    if (lane == 0) s_warp_cnt[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < JOIN_THREADS / 32; w++) {
            const uint32_t c = s_warp_cnt[w];
            s_warp_cnt[w] = total;
            total += c;
        }
        s_base = total ? atomicAdd(&P.counters[CNT_PAIRS], total) : 0u;
    }
    __syncthreads();
    if (emit && !far) {
        const uint32_t at = s_base + s_warp_cnt[threadIdx.x >> 5] + __popc(m & lt);
        if (at < J.pair_cap) {
            reinterpret_cast<ulonglong2 *>(J.pair)[at] = make_ulonglong2(ent.lo, ent.hi);
            J.pair_hk[at] = h;
        } else atomicOr(&P.counters[CNT_ERR], DEV_ERR_CAPACITY);
    }
    if (emit) {
        if (far) {
            const uint32_t at = atomicAdd(&P.counters[CNT_PAIRS_FAR], 1u);
            if (at < J.far_cap) {
                reinterpret_cast<ulonglong2 *>(J.pair_far)[at] = make_ulonglong2(ent.lo, ent.hi);
                J.pair_far_hk[at] = h;
            } else atomicOr(&P.counters[CNT_ERR], DEV_ERR_CAPACITY);
        }
        J.mate_of[i1] = (uint32_t) (i2 + P.idx_base);
    }
}

uint32_t local_join_max_grid(int sms) { return (uint32_t) sms * LJ_CTAS_PER_SM; }

// tiles per CTA and grid for n records on `sms` SMs: contiguous tile ranges, at least 8 tiles each (pairs across a
// range boundary go through the global join)
void local_join_shape(uint64_t n, int sms, uint32_t *grid_out, uint32_t *tiles_per_cta) {
    const uint64_t n_tiles = (n + LJ_TILE - 1) / LJ_TILE;
    uint64_t grid = local_join_max_grid(sms);
    if (grid > (n_tiles + 7) / 8) grid = (n_tiles + 7) / 8;
    if (grid < 1) grid = 1;
    const uint64_t tpc = (n_tiles + grid - 1) / grid;
    grid = (n_tiles + tpc - 1) / tpc;
    *grid_out = (uint32_t) grid;
    *tiles_per_cta = (uint32_t) tpc;
}

int launch_local_join(const JoinParams &P, LocalJoinParams J, int sms, cudaStream_t stream, uint64_t *launches, oge_gpu_dedup_ctx *c) {
    if (P.n == 0) return 0;
    uint32_t grid = 0, tpc = 0;
    local_join_shape(P.n, sms, &grid, &tpc);
    J.n_buckets = LJ_BUCKETS;
    J.tiles_per_cta = tpc;
    J.couples_per_cta = tpc * (LJ_TILE / 2);
    const size_t smem = (size_t) LJ_BUCKETS * LJ_WAYS * LJ_ENTRY_BYTES;
    OGE_CUDA_TRY(cudaFuncSetAttribute(local_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    if (c) c->k_begin(OGE_K_MATCH, stream);
    local_match_kernel<<<grid, LJ_THREADS, smem, stream>>>(P, J);
    if (c) { c->k_end(stream); c->k_begin(OGE_K_EMIT, stream); }
    const uint64_t slots = (uint64_t) grid * J.couples_per_cta;
    local_emit_kernel<<<(uint32_t) ((slots + JOIN_THREADS - 1) / JOIN_THREADS), JOIN_THREADS, 0, stream>>>(P, J, grid);
    if (c) c->k_end(stream);
    *launches += 2;
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

int launch_mate_complex(const JoinParams &P, const E128 *sorted_cplx, uint32_t n_cplx, uint8_t *state, void *long_segs /* n_cplx / CPLX_LONG + 1 uint2 */,
                        cudaStream_t stream, uint64_t *launches) {
    if (n_cplx == 0) return 0;
    OGE_CUDA_TRY(cudaMemsetAsync(P.counters + CNT_SCRATCH0, 0, 4, stream));
    mate_complex_kernel<<<(n_cplx + JOIN_THREADS - 1) / JOIN_THREADS, JOIN_THREADS, 0, stream>>>(P, sorted_cplx, n_cplx, state, (uint2 *) long_segs);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    uint32_t n_long = 0;
    OGE_CUDA_TRY(cudaMemcpyAsync(&n_long, P.counters + CNT_SCRATCH0, 4, cudaMemcpyDeviceToHost, stream));
    OGE_CUDA_TRY(cudaStreamSynchronize(stream));
    if (n_long) {
        mate_complex_long_kernel<<<n_long, JOIN_THREADS, 0, stream>>>(P, sorted_cplx, state, (const uint2 *) long_segs);
        *launches += 1;
        OGE_CUDA_TRY(cudaGetLastError());
    }
    return 0;
}
size_t mate_complex_long_segs_bytes(uint32_t n_cplx) { return ((size_t) n_cplx / CPLX_LONG + 2) * sizeof(uint2); }

}  // namespace oge
