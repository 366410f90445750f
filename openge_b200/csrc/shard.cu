// Range-sharded dedup across GPUs: the device work between the exchanges.
//
// The reference has no counterpart (it is one process); what has to be preserved is that the
// result equals its single-stream run (`openge dedup --nosplit -v`):
//   * the pairing of ReadEndsMap (util/picard_structures.h:82-109, algorithms/mark_duplicates.cpp:209-246)
//     is a sequential toggle over the whole file, so every name that is not a plain couple inside
//     one shard is resolved over the union of all shards' sightings, in global file order;
//   * a duplicate group (areComparableForDuplicates, :402-414) must be evaluated in one place, so
//     every end entry lives on the rank that owns its key range.
// All lists that cross ranks are small (cross-shard mates, boundary ends, marks); the exchange
// delivers every rank's list to every rank and the receiving kernels keep what concerns them.
#include "keyview.cuh"
#include "pairing.cuh"

namespace oge {

constexpr int SH_THREADS = 256;

// rank owning the key range that holds (ref, coord): number of split keys <= the packed key
__device__ __forceinline__ int owner_of(const ShardParams &S, uint64_t packed) {
    int o = 0;
    for (int r = 0; r + 1 < S.world; r++) o += S.split[r] <= packed ? 1 : 0;
    return o;
}
__device__ __forceinline__ uint64_t frag_packed(const KeyLayout &L, const E128 &e) { return bits_get(e, L.f_coord, L.coord_bits + L.ref_bits); }
__device__ __forceinline__ uint64_t pair_packed(const KeyLayout &L, const E128 &e) { return bits_get(e, L.p_coord1, L.coord_bits + L.ref_bits); }
__device__ __forceinline__ bool is_dead(const E128 &e) { return (e.lo & e.hi) == ~0ull; }

// warp-aggregated append position (all lanes of the warp must call)
__device__ __forceinline__ uint32_t warp_append(bool want, uint32_t *counter) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, want);
    if (!m) return 0;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t) __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

// ---- phase 1: who gets published -------------------------------------------------------------------
// names seen once locally: their slot still holds one arrival
__global__ void __launch_bounds__(SH_THREADS) sh_singletons_kernel(const MateSlot *__restrict__ table, uint64_t n_slots,
                                                                   uint32_t *__restrict__ list, uint32_t *__restrict__ counters) {
    const uint64_t s = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    bool want = false;
    uint32_t ord = 0;
    if (s < n_slots) {
        const ulonglong2 kv = *reinterpret_cast<const ulonglong2 *>(&table[s]);      // key, val
        want = kv.x != 0 && (kv.y >> 32) == 1;
        ord = (uint32_t) kv.y - 1u;
    }
    const uint32_t at = warp_append(want, &counters[CNT_PUB]);
    if (want) list[at] = ord;
}

// records on the exact-path list (names seen three or more times locally, hash-equal couples)
__global__ void __launch_bounds__(SH_THREADS) sh_complex_kernel(const E128 *__restrict__ cplx, uint32_t n_cplx, uint32_t *__restrict__ list,
                                                                uint32_t *__restrict__ counters) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    const bool want = j < n_cplx;
    const uint32_t at = warp_append(want, &counters[CNT_PUB]);
    if (want) list[at] = (uint32_t) cplx[j].lo;
}

__global__ void __launch_bounds__(SH_THREADS) sh_gather_kernel(const uint32_t *__restrict__ list, uint32_t n_list, const E128 *__restrict__ frag,
                                                               const uint64_t *__restrict__ hk, const NameTag *__restrict__ tag,
                                                               PubEntry *__restrict__ out) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_list) return;
    const uint32_t i = list[j];
    PubEntry p;
    p.frag = ld_frag(frag + i);
    p.hk = hk[i];
    p.rsv = 0;
    p.tag = tag[i];
    out[j] = p;
}

// ---- phase 2: a name published elsewhere that is a complete couple here -------------------------------
// The couple is retracted (the sequential toggle may pair its records differently once the other
// shards' sightings are interleaved) and both records are published in the second round.
__global__ void __launch_bounds__(SH_THREADS) sh_probe_kernel(const PubEntry *__restrict__ pub, uint64_t n_pub, ShardParams S,
                                                              MateSlot *__restrict__ table, uint64_t n_slots, E128 *__restrict__ pair,
                                                              E128 *__restrict__ pair_far, uint32_t *__restrict__ list2) {
    const uint64_t j = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_pub) return;
    const uint64_t g = bits_get(pub[j].frag, S.kl.f_idx, S.kl.idx_bits);
    if (g - S.idx_base < S.n) return;      // one of mine
    const uint64_t h = pub[j].hk;
    uint64_t s = slot_of(h, n_slots);
    while (true) {
        const uint64_t k = table[s].key;
        if (k == 0) return;
        if (k == h) break;
        if (++s == n_slots) s = 0;
    }
    unsigned long long *vp = reinterpret_cast<unsigned long long *>(&table[s].val);
    const unsigned long long v = *vp;
    if ((v >> 32) != 2) return;                       // not a couple, or already retracted (bit 63 set)
    if (atomicCAS(vp, v, v | (1ull << 63)) != v) return;      // another entry of the same name got here first
    const uint32_t pos = table[s].pair_pos;
    if (pos == SLOT_NO_PAIR) return;                  // hash-equal records with different names: published in round 1
    const bool far = (pos & SLOT_PAIR_FAR) != 0;
    reinterpret_cast<ulonglong2 *>(far ? pair_far : pair)[pos & ~SLOT_PAIR_FAR] = make_ulonglong2(~0ull, ~0ull);
    atomicAdd(&S.counters[far ? CNT_FAR_RETRACTED : CNT_PAIRS_RETRACTED], 1u);
    const uint32_t at = atomicAdd(&S.counters[CNT_PUB], 2u);
    list2[at] = (uint32_t) (table[s].who >> 32);
    list2[at + 1] = (uint32_t) table[s].who;
}

// ---- phase 3: replay of the published set -------------------------------------------------------------
// sort entry: hi = key hash, lo = (global ordinal << 32) | position in the published array; sorted on bits
// [32, 96): by the low half of the hash, then by file order.  A segment of equal low halves may mix
// several hashes; the replay compares whole keys (and whole hashes), so that only costs comparisons.
__global__ void __launch_bounds__(SH_THREADS) sh_wbuild_kernel(const PubEntry *__restrict__ w, uint32_t n_w, KeyLayout L, E128 *__restrict__ out) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_w) return;
    E128 e;
    e.hi = w[j].hk;
    e.lo = (bits_get(w[j].frag, L.f_idx, L.idx_bits) << 32) | j;
    out[j] = e;
}

// The reference's map key is the byte string RG value + ":" + name (mark_duplicates.cpp:214), so
// ("a", ":x") and ("a:", "x") are the SAME key.  Rebuild that string from a tag: the read-group
// value comes back out of the host-resolved @RG table, the name out of the tag words.  Name bytes
// beyond the 29 a tag holds are not available here: two keys that agree on length, on everything
// a tag holds and on the 64-bit hash of the whole key (they are in the same hash segment) are taken as equal.
struct PubKey {
    const uint8_t *rg;
    uint32_t rg_len, name_len;
    const NameTag *t;
    bool opaque;      // read group not listed in the header: its bytes are unknown on this rank
};
__device__ __forceinline__ PubKey pub_key(const NameTag &t, const RgTable &rg) {
    PubKey k;
    const uint32_t code = t.w[0] & 0xFFFFu, l_name = (t.w[0] >> 16) & 0xFFu;
    k.t = &t;
    k.name_len = l_name ? l_name - 1 : 0;
    k.opaque = code == RGC_UNKNOWN;
    k.rg = rg.bytes;
    k.rg_len = 0;
    if (code < RGC_UNKNOWN && (int) code < rg.n) {
        k.rg = rg.bytes + rg.off[code];
        k.rg_len = rg.off[code + 1] - rg.off[code];
    }
    return k;
}
__device__ __forceinline__ int pub_key_byte(const PubKey &k, uint32_t i) {      // -1: beyond what the tag holds
    if (i < k.rg_len) return k.rg[i];
    if (i == k.rg_len) return ':';
    const uint32_t j = i - k.rg_len - 1;
    if (j >= NAME_TAG_BYTES) return -1;
    return j == 0 ? (int) (k.t->w[0] >> 24) : (int) ((k.t->w[(j + 3) >> 2] >> (8 * ((j + 3) & 3))) & 0xFFu);
}
__device__ bool pub_keys_equal(const NameTag &ta, const NameTag &tb, const RgTable &rg) {
    const PubKey a = pub_key(ta, rg), b = pub_key(tb, rg);
    if (a.opaque || b.opaque) {      // fall back to the tags themselves (plus the shared hash)
        bool same = true;
        for (int k = 0; k < 8; k++) same = same && ta.w[k] == tb.w[k];
        return same;
    }
    const uint32_t la = a.rg_len + 1 + a.name_len, lb = b.rg_len + 1 + b.name_len;
    if (la != lb) return false;
    for (uint32_t i = 0; i < la; i++) {
        const int x = pub_key_byte(a, i), y = pub_key_byte(b, i);
        if (x >= 0 && y >= 0 && x != y) return false;
    }
    return true;
}

// One thread per hash value: the reference's toggle (tmp.put / tmp.remove, mark_duplicates.cpp:216-223)
// over that hash's sightings in global file order, keys compared as pub_keys_equal does.  Every rank
// replays the same list and keeps the pairs whose key it owns.
__global__ void __launch_bounds__(SH_THREADS) sh_replay_kernel(const E128 *__restrict__ sorted, uint32_t n_w, const PubEntry *__restrict__ w,
                                                               uint8_t *__restrict__ state, ShardParams S, E128 *__restrict__ pair,
                                                               uint32_t pair_cap, E128 *__restrict__ pair_far, uint32_t far_cap,
                                                               uint32_t *__restrict__ mate_of, uint64_t *__restrict__ fm, uint32_t fm_cap,
                                                               RgTable rg) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_w) return;
    const uint32_t h = (uint32_t) sorted[j].hi;
    if (j > 0 && (uint32_t) sorted[j - 1].hi == h) return;      // not a segment head
    uint32_t end = j;
    while (end < n_w && (uint32_t) sorted[end].hi == h) state[end++] = 0;
    for (uint32_t a = j; a < end; a++) {
        const PubEntry &pa = w[(uint32_t) sorted[a].lo];
        if (a > j && sorted[a].lo >> 32 == sorted[a - 1].lo >> 32) continue;      // the same record published twice
        int found = -1;
        for (uint32_t b = j; b < a; b++) {
            if (!state[b]) continue;
            if (sorted[b].hi == sorted[a].hi && pub_keys_equal(pa.tag, w[(uint32_t) sorted[b].lo].tag, rg)) { found = (int) b; break; }
        }
        if (found < 0) { state[a] = 1; continue; }
        state[found] = 0;
        const PubEntry &pb = w[(uint32_t) sorted[found].lo];      // the earlier sighting
        uint32_t i1, i2;
        bool far;
        const E128 ent = make_pair_entry(S.kl, pb.frag, pa.frag, &i1, &i2, 0, &far);      // global ordinals
        if (owner_of(S, pair_packed(S.kl, ent)) != S.rank) continue;
        const uint32_t pos = atomicAdd(&S.counters[far ? CNT_PAIRS_FAR : CNT_PAIRS], 1u);
        if (pos < (far ? far_cap : pair_cap)) reinterpret_cast<ulonglong2 *>(far ? pair_far : pair)[pos] = make_ulonglong2(ent.lo, ent.hi);
        const uint64_t l1 = (uint64_t) i1 - S.idx_base;
        if (l1 < S.n) mate_of[l1] = i2;
        else {
            const uint32_t at = atomicAdd(&S.counters[CNT_FM], 1u);
            if (at < fm_cap) fm[at] = ((uint64_t) i1 << 32) | i2;
        }
    }
}

// ---- phase 4: entries whose key belongs to another rank -------------------------------------------------
// kind 0: fragment entries, 1: near pairs, 2: far pairs.  dry 0: routed entries leave the local list (all-ones =
// dead); 1: count only; 2: a copy leaves (fragment ends: K4 skips the runs whose key another rank owns).
__global__ void __launch_bounds__(SH_THREADS) sh_route_kernel(E128 *__restrict__ ents, uint64_t n_ents, int kind, ShardParams S,
                                                              const uint32_t *__restrict__ mate_of, const uint64_t *__restrict__ fm, uint32_t n_fm,
                                                              RouteEntry *__restrict__ out, uint32_t out_cap, int dry) {
    const uint64_t j = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    bool want = false;
    E128 e;
    e.lo = e.hi = 0;
    if (j < n_ents) {
        e = ld_frag(ents + j);
        if (!is_dead(e)) want = owner_of(S, kind ? pair_packed(S.kl, e) : frag_packed(S.kl, e)) != S.rank;
    }
    const uint32_t at = warp_append(want, &S.counters[CNT_ROUTE]);
    if (!want || dry == 1 || at >= out_cap) return;      // no room: the entry stays and is picked up by the next sweep
    if (kind) atomicAdd(&S.counters[kind == 1 ? CNT_SCRATCH0 : CNT_SCRATCH1], 1u);      // pair entries that leave
    RouteEntry r;
    r.e = e;
    r.kind = (uint32_t) kind;
    r.idx2 = 0;
    r.rsv = 0;
    if (kind) {
        const uint32_t i1 = (uint32_t) bits_get(e, S.kl.p_idx, S.kl.idx_bits);
        const uint64_t l1 = (uint64_t) i1 - S.idx_base;
        if (l1 < S.n) r.idx2 = mate_of[l1];
        else if (n_fm) {
            uint32_t lo = 0, hi = n_fm;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if ((uint32_t) (fm[mid] >> 32) < i1) lo = mid + 1; else hi = mid;
            }
            r.idx2 = (uint32_t) fm[lo];
        }
    }
    out[at] = r;
    if (dry == 0) reinterpret_cast<ulonglong2 *>(ents)[j] = make_ulonglong2(~0ull, ~0ull);      // dry == 2: a copy leaves, the entry stays
}

__global__ void __launch_bounds__(SH_THREADS) sh_receive_kernel(const RouteEntry *__restrict__ in, uint64_t n_in, ShardParams S, uint32_t kinds,
                                                                E128 *__restrict__ frag_extra, uint32_t frag_cap, E128 *__restrict__ pair,
                                                                uint32_t pair_cap, E128 *__restrict__ pair_far, uint32_t far_cap,
                                                                uint32_t *__restrict__ mate_of, uint64_t *__restrict__ fm, uint32_t fm_cap) {
    const uint64_t j = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_in) return;
    const RouteEntry r = in[j];
    if (!((kinds >> r.kind) & 1u)) return;      // not in this call
    if (owner_of(S, r.kind ? pair_packed(S.kl, r.e) : frag_packed(S.kl, r.e)) != S.rank) return;
    if (r.kind == 0) {
        const uint32_t at = atomicAdd(&S.counters[CNT_FRAG_EXTRA], 1u);
        if (at < frag_cap) reinterpret_cast<ulonglong2 *>(frag_extra)[at] = make_ulonglong2(r.e.lo, r.e.hi);
    } else {
        const bool far = r.kind == 2;
        const uint32_t pos = atomicAdd(&S.counters[far ? CNT_PAIRS_FAR : CNT_PAIRS], 1u);
        if (pos < (far ? far_cap : pair_cap)) reinterpret_cast<ulonglong2 *>(far ? pair_far : pair)[pos] = make_ulonglong2(r.e.lo, r.e.hi);
        const uint32_t i1 = (uint32_t) bits_get(r.e, S.kl.p_idx, S.kl.idx_bits);
        const uint64_t l1 = (uint64_t) i1 - S.idx_base;
        if (l1 < S.n) mate_of[l1] = r.idx2;
        else {
            const uint32_t at = atomicAdd(&S.counters[CNT_FM], 1u);
            if (at < fm_cap) fm[at] = ((uint64_t) i1 << 32) | r.idx2;
        }
    }
}

// foreign-mate couples <-> sortable 16-byte entries (sorted by idx1 = bits [32, 64))
__global__ void __launch_bounds__(SH_THREADS) sh_fm_pack_kernel(const uint64_t *__restrict__ fm, uint32_t n, E128 *__restrict__ out) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    if (j < n) { E128 e; e.lo = fm[j]; e.hi = 0; out[j] = e; }
}
__global__ void __launch_bounds__(SH_THREADS) sh_fm_unpack_kernel(const E128 *__restrict__ in, uint32_t n, uint64_t *__restrict__ fm) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    if (j < n) fm[j] = in[j].lo;
}

// ---- phase 6: marks decided on other ranks --------------------------------------------------------------
__global__ void __launch_bounds__(SH_THREADS) sh_apply_marks_kernel(const uint32_t *__restrict__ marks, uint64_t n_marks, uint64_t idx_base,
                                                                    uint64_t n, uint8_t *__restrict__ dup) {
    const uint64_t j = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_marks) return;
    const uint64_t l = (uint64_t) marks[j] - idx_base;
    if (l < n) dup[l] = 1;
}


// ======================================================================================================
// Owner-routed protocol: what crosses ranks goes only where it is needed.
//   published entries -> the rank that OWNS the name (key hash mod world), which alone replays it
//   key hashes of the published entries -> every rank (8 bytes each), for the retraction of local pairs
//   end entries whose key lies in another rank's range -> that rank; marks -> the rank that holds the record
// A published entry carries the whole key string RG value + ":" + name, so the owner's comparison is exact
// whatever the name length and whether or not the header lists the read group:
//   [frag E128 : 16][key hash : 8][key length : 4][reserved : 4][key bytes ... padded to the entry size]
// The entry size is agreed by all ranks from the longest key (oge_gpu_shard_key_bytes / _set_entry_bytes).

__device__ __forceinline__ const uint8_t *pub2_at(const uint8_t *base, uint32_t stride, uint64_t j) { return base + j * stride; }

__global__ void __launch_bounds__(SH_THREADS) sh_keylen_kernel(const uint8_t *__restrict__ rec, const uint64_t *__restrict__ off, uint64_t n,
                                                               const uint64_t *__restrict__ hk, uint32_t *__restrict__ out_max) {
    const uint64_t i = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    uint32_t l = 0;
    if (i < n && (!hk || hk[i])) {
        const KeyView v = key_view(rec, off, i);
        l = v.rg_len + 1 + v.name_len;
    }
    for (int o = 16; o; o >>= 1) l = max(l, __shfl_xor_sync(0xFFFFFFFFu, l, o));
    if ((threadIdx.x & 31) == 0 && l) atomicMax(out_max, l);
}

__global__ void __launch_bounds__(SH_THREADS) sh_gather2_kernel(const uint32_t *__restrict__ list, uint32_t n_list, const E128 *__restrict__ frag,
                                                                const uint64_t *__restrict__ hk, const uint8_t *__restrict__ rec,
                                                                const uint64_t *__restrict__ off, uint8_t *__restrict__ out, uint32_t stride,
                                                                uint64_t *__restrict__ hashes, uint32_t *__restrict__ err) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_list) return;
    const uint32_t i = list[j];
    uint8_t *o = out + (uint64_t) j * stride;
    const E128 f = ld_frag(frag + i);
    const KeyView v = key_view(rec, off, i);
    const uint32_t len = v.rg_len + 1 + v.name_len;
    reinterpret_cast<ulonglong2 *>(o)[0] = make_ulonglong2(f.lo, f.hi);
    reinterpret_cast<uint64_t *>(o)[2] = hk[i];
    reinterpret_cast<uint32_t *>(o)[6] = len;
    reinterpret_cast<uint32_t *>(o)[7] = 0;
    if (32 + len > stride) { atomicOr(err, DEV_ERR_CAPACITY); return; }
    for (uint32_t b = 0; b < stride - 32; b++) o[32 + b] = b < len ? key_byte(v, b) : (uint8_t) 0;
    if (hashes) hashes[j] = hk[i];
}

// ---- a set of 64-bit hashes (open addressing, 0 = empty) -----------------------------------------------
__global__ void __launch_bounds__(SH_THREADS) sh_set_build_kernel(const uint64_t *__restrict__ h, uint64_t n, unsigned long long *__restrict__ set,
                                                                  uint64_t n_slots) {
    const uint64_t j = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n) return;
    const unsigned long long k = h[j];
    if (!k) return;
    uint64_t s = slot_of(k, n_slots);
    while (true) {
        const unsigned long long old = atomicCAS(set + s, 0ull, k);
        if (old == 0 || old == k) return;
        if (++s == n_slots) s = 0;
    }
}
__device__ __forceinline__ bool sh_set_has(const unsigned long long *set, uint64_t n_slots, uint64_t k) {
    uint64_t s = slot_of(k, n_slots);
    while (true) {
        const unsigned long long v = set[s];
        if (v == k) return true;
        if (v == 0) return false;
        if (++s == n_slots) s = 0;
    }
}

// A couple the global join over the leftovers formed (its slot holds two arrivals) whose name some other rank published:
// retracted, both records published in the second round.
__global__ void __launch_bounds__(SH_THREADS) sh_probe_table_kernel(MateSlot *__restrict__ table, uint64_t n_slots, const unsigned long long *__restrict__ set,
                                                                    uint64_t set_slots, E128 *__restrict__ pair, E128 *__restrict__ pair_far,
                                                                    uint32_t *__restrict__ list2, uint32_t *__restrict__ counters) {
    const uint64_t s = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    if (s >= n_slots) return;
    const uint64_t k = table[s].key;
    if (!k || (table[s].val >> 32) != 2) return;      // not a couple (or complex / retracted: those were published in round 1)
    if (!sh_set_has(set, set_slots, k)) return;
    const uint32_t pos = table[s].pair_pos;
    if (pos == SLOT_NO_PAIR) return;                  // hash-equal records with different names: published in round 1
    table[s].val |= 1ull << 63;
    const bool far = (pos & SLOT_PAIR_FAR) != 0;
    reinterpret_cast<ulonglong2 *>(far ? pair_far : pair)[pos & ~SLOT_PAIR_FAR] = make_ulonglong2(~0ull, ~0ull);
    atomicAdd(&counters[far ? CNT_FAR_RETRACTED : CNT_PAIRS_RETRACTED], 1u);
    const uint32_t at = atomicAdd(&counters[CNT_PUB], 2u);
    list2[at] = (uint32_t) (table[s].who >> 32);
    list2[at + 1] = (uint32_t) table[s].who;
}

// The same for the pairs the windowed join settled inside its CTAs (they are in no table: their hashes are kept by list position).
__global__ void __launch_bounds__(SH_THREADS) sh_probe_pairs_kernel(const uint64_t *__restrict__ pair_hk, uint32_t n_pairs, const unsigned long long *__restrict__ set,
                                                                    uint64_t set_slots, E128 *__restrict__ list, int far, const uint32_t *__restrict__ mate_of,
                                                                    ShardParams S, uint32_t *__restrict__ list2) {
    const uint32_t pos = blockIdx.x * SH_THREADS + threadIdx.x;
    if (pos >= n_pairs) return;
    const uint64_t h = pair_hk[pos];
    if (!h || !sh_set_has(set, set_slots, h)) return;
    const E128 ent = ld_frag(list + pos);
    if (is_dead(ent)) return;
    const uint32_t i1 = (uint32_t) (bits_get(ent, S.kl.p_idx, S.kl.idx_bits) - S.idx_base);
    const uint32_t i2 = (uint32_t) (mate_of[i1] - S.idx_base);
    reinterpret_cast<ulonglong2 *>(list)[pos] = make_ulonglong2(~0ull, ~0ull);
    atomicAdd(&S.counters[far ? CNT_FAR_RETRACTED : CNT_PAIRS_RETRACTED], 1u);
    const uint32_t at = atomicAdd(&S.counters[CNT_PUB], 2u);
    list2[at] = i1;
    list2[at + 1] = i2;
}

// ---- replay at the name's owner ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SH_THREADS) sh_wbuild2_kernel(const uint8_t *__restrict__ w, uint32_t stride, uint32_t n_w, KeyLayout L,
                                                                E128 *__restrict__ out) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_w) return;
    const uint8_t *p = pub2_at(w, stride, j);
    const E128 f = ld_frag(reinterpret_cast<const E128 *>(p));
    E128 e;
    e.hi = reinterpret_cast<const uint64_t *>(p)[2];
    e.lo = (bits_get(f, L.f_idx, L.idx_bits) << 32) | j;
    out[j] = e;
}

__device__ __forceinline__ bool pub2_keys_equal(const uint8_t *a, const uint8_t *b) {
    const uint32_t la = reinterpret_cast<const uint32_t *>(a)[6], lb = reinterpret_cast<const uint32_t *>(b)[6];
    if (la != lb) return false;
    const uint32_t *wa = reinterpret_cast<const uint32_t *>(a + 32), *wb = reinterpret_cast<const uint32_t *>(b + 32);
    for (uint32_t k = 0; k < (la + 3) / 4; k++)      // the padding behind the key is zero in both
        if (wa[k] != wb[k]) return false;
    return true;
}

// One thread per hash value: the reference's toggle (tmp.put / tmp.remove, mark_duplicates.cpp:216-223) over that hash's
// sightings in global file order, keys compared byte by byte.  A pair whose key range this rank owns joins its lists; the
// others are handed to their owners (out, counters[CNT_ROUTE]).
__global__ void __launch_bounds__(SH_THREADS) sh_replay2_kernel(const E128 *__restrict__ sorted, uint32_t n_w, const uint8_t *__restrict__ w, uint32_t stride,
                                                                uint8_t *__restrict__ state, ShardParams S, E128 *__restrict__ pair,
                                                                uint32_t pair_cap, E128 *__restrict__ pair_far, uint32_t far_cap,
                                                                uint32_t *__restrict__ mate_of, uint64_t *__restrict__ fm, uint32_t fm_cap,
                                                                RouteEntry *__restrict__ out, uint32_t out_cap) {
    const uint32_t j = blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n_w) return;
    const uint32_t h = (uint32_t) sorted[j].hi;
    if (j > 0 && (uint32_t) sorted[j - 1].hi == h) return;      // not a segment head
    uint32_t end = j;
    while (end < n_w && (uint32_t) sorted[end].hi == h) end++;
    // the map holds at most one unmatched sighting per distinct key: a handful of live entries per segment, kept in
    // registers; the byte map is the fall-back should more than SH_LIVE distinct keys ever share the low hash word
    constexpr int SH_LIVE = 8;
    uint32_t live[SH_LIVE];
    int n_live = 0;
    bool overflow = false;
    for (uint32_t a = j; a < end; a++) {
        const uint8_t *pa = pub2_at(w, stride, (uint32_t) sorted[a].lo);
        if (a > j && sorted[a].lo >> 32 == sorted[a - 1].lo >> 32) continue;      // the same record published twice
        int found = -1;
        if (!overflow) {
            for (int t = 0; t < n_live; t++)
                if (sorted[live[t]].hi == sorted[a].hi && pub2_keys_equal(pa, pub2_at(w, stride, (uint32_t) sorted[live[t]].lo))) {
                    found = (int) live[t];
                    live[t] = live[--n_live];
                    break;
                }
        } else {
            for (uint32_t b = j; b < a; b++) {
                if (!state[b]) continue;
                if (sorted[b].hi == sorted[a].hi && pub2_keys_equal(pa, pub2_at(w, stride, (uint32_t) sorted[b].lo))) { found = (int) b; break; }
            }
        }
        if (found < 0) {
            if (overflow) state[a] = 1;
            else if (n_live < SH_LIVE) live[n_live++] = a;
            else {
                overflow = true;
                for (uint32_t x = j; x < a; x++) state[x] = 0;
                for (int t = 0; t < n_live; t++) state[live[t]] = 1;
                state[a] = 1;
            }
            continue;
        }
        if (overflow) state[found] = 0;
        const uint8_t *pb = pub2_at(w, stride, (uint32_t) sorted[found].lo);      // the earlier sighting
        uint32_t i1, i2;
        bool far;
        const E128 ent = make_pair_entry(S.kl, ld_frag(reinterpret_cast<const E128 *>(pb)), ld_frag(reinterpret_cast<const E128 *>(pa)), &i1, &i2, 0, &far);
        if (owner_of(S, pair_packed(S.kl, ent)) != S.rank) {
            const uint32_t at = atomicAdd(&S.counters[CNT_ROUTE], 1u);
            if (at < out_cap) {
                RouteEntry r;
                r.e = ent;
                r.idx2 = i2;
                r.kind = far ? 2u : 1u;
                r.rsv = 0;
                out[at] = r;
            }
            continue;
        }
        const uint32_t pos = atomicAdd(&S.counters[far ? CNT_PAIRS_FAR : CNT_PAIRS], 1u);
        if (pos < (far ? far_cap : pair_cap)) reinterpret_cast<ulonglong2 *>(far ? pair_far : pair)[pos] = make_ulonglong2(ent.lo, ent.hi);
        const uint64_t l1 = (uint64_t) i1 - S.idx_base;
        if (l1 < S.n) mate_of[l1] = i2;
        else {
            const uint32_t at = atomicAdd(&S.counters[CNT_FM], 1u);
            if (at < fm_cap) fm[at] = ((uint64_t) i1 << 32) | i2;
        }
    }
}

// ---- bucketing by destination rank: what a rank sends is ordered by destination, so that the exchange is one
//      all-to-all with uneven splits.  kind 0: published entries (owner of the name = hash mod world), 1: routed end
//      entries (owner of the key range), 2: marks (rank whose record range holds the ordinal; bases = world + 1 ordinals).
__device__ __forceinline__ int sh_dest(int kind, const uint8_t *item, const ShardParams &S, const uint64_t *bases) {
    if (kind == 0) return (int) (reinterpret_cast<const uint64_t *>(item)[2] % (uint64_t) S.world);
    if (kind == 1) {
        const RouteEntry *r = reinterpret_cast<const RouteEntry *>(item);
        return owner_of(S, r->kind ? pair_packed(S.kl, r->e) : frag_packed(S.kl, r->e));
    }
    const uint64_t g = *reinterpret_cast<const uint32_t *>(item);
    int o = 0;
    for (int r = 1; r < S.world; r++) o += bases[r] <= g ? 1 : 0;
    return o;
}

__global__ void __launch_bounds__(SH_THREADS) sh_bucket_count_kernel(const uint8_t *__restrict__ items, uint64_t n, uint32_t item_bytes, int kind,
                                                                     ShardParams S, const uint64_t *__restrict__ bases, uint32_t *__restrict__ count) {
    const uint64_t j = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n) return;
    atomicAdd(&count[sh_dest(kind, items + j * item_bytes, S, bases)], 1u);
}

// start[d] = first slot of destination d in `out`; fill[d] counts up from zero
__global__ void __launch_bounds__(SH_THREADS) sh_bucket_scatter_kernel(const uint8_t *__restrict__ items, uint64_t n, uint32_t item_bytes, int kind,
                                                                       ShardParams S, const uint64_t *__restrict__ bases, const uint32_t *__restrict__ start,
                                                                       uint32_t *__restrict__ fill, uint8_t *__restrict__ out) {
    const uint64_t j = (uint64_t) blockIdx.x * SH_THREADS + threadIdx.x;
    if (j >= n) return;
    const uint8_t *src = items + j * item_bytes;
    const int d = sh_dest(kind, src, S, bases);
    uint8_t *dst = out + (uint64_t) (start[d] + atomicAdd(&fill[d], 1u)) * item_bytes;
    for (uint32_t b = 0; b < item_bytes; b += 4) *reinterpret_cast<uint32_t *>(dst + b) = *reinterpret_cast<const uint32_t *>(src + b);
}

// ---- launchers --------------------------------------------------------------------------------------------
static inline uint32_t grid_for(uint64_t n) { return (uint32_t) ((n + SH_THREADS - 1) / SH_THREADS); }

#define SH_LAUNCH(n, call)                          \
    do {                                            \
        if ((n) > 0) {                              \
            call;                                   \
            *launches += 1;                         \
            OGE_CUDA_TRY(cudaGetLastError());       \
        }                                           \
    } while (0)

int launch_sh_singletons(const MateSlot *table, uint64_t n_slots, uint32_t *list, uint32_t *counters, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_slots, (sh_singletons_kernel<<<grid_for(n_slots), SH_THREADS, 0, s>>>(table, n_slots, list, counters)));
    return 0;
}
int launch_sh_complex(const E128 *cplx, uint32_t n_cplx, uint32_t *list, uint32_t *counters, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_cplx, (sh_complex_kernel<<<grid_for(n_cplx), SH_THREADS, 0, s>>>(cplx, n_cplx, list, counters)));
    return 0;
}
int launch_sh_gather(const uint32_t *list, uint32_t n_list, const E128 *frag, const uint64_t *hk, const NameTag *tag, PubEntry *out,
                     cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_list, (sh_gather_kernel<<<grid_for(n_list), SH_THREADS, 0, s>>>(list, n_list, frag, hk, tag, out)));
    return 0;
}
int launch_sh_probe(const PubEntry *pub, uint64_t n_pub, const ShardParams &S, MateSlot *table, uint64_t n_slots, E128 *pair, E128 *pair_far,
                    uint32_t *list2, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_pub, (sh_probe_kernel<<<grid_for(n_pub), SH_THREADS, 0, s>>>(pub, n_pub, S, table, n_slots, pair, pair_far, list2)));
    return 0;
}
int launch_sh_wbuild(const PubEntry *w, uint32_t n_w, const KeyLayout &L, E128 *out, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_w, (sh_wbuild_kernel<<<grid_for(n_w), SH_THREADS, 0, s>>>(w, n_w, L, out)));
    return 0;
}
int launch_sh_replay(const E128 *sorted, uint32_t n_w, const PubEntry *w, uint8_t *state, const ShardParams &S, E128 *pair, uint32_t pair_cap,
                     E128 *pair_far, uint32_t far_cap, uint32_t *mate_of, uint64_t *fm, uint32_t fm_cap, const RgTable &rg, cudaStream_t s,
                     uint64_t *launches) {
    SH_LAUNCH(n_w, (sh_replay_kernel<<<grid_for(n_w), SH_THREADS, 0, s>>>(sorted, n_w, w, state, S, pair, pair_cap, pair_far, far_cap, mate_of,
                                                                           fm, fm_cap, rg)));
    return 0;
}
int launch_sh_route(E128 *ents, uint64_t n_ents, int kind, const ShardParams &S, const uint32_t *mate_of, const uint64_t *fm, uint32_t n_fm,
                    RouteEntry *out, uint32_t out_cap, int dry, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_ents, (sh_route_kernel<<<grid_for(n_ents), SH_THREADS, 0, s>>>(ents, n_ents, kind, S, mate_of, fm, n_fm, out, out_cap, dry)));
    return 0;
}
int launch_sh_receive(const RouteEntry *in, uint64_t n_in, const ShardParams &S, uint32_t kinds, E128 *frag_extra, uint32_t frag_cap, E128 *pair,
                      uint32_t pair_cap, E128 *pair_far, uint32_t far_cap, uint32_t *mate_of, uint64_t *fm, uint32_t fm_cap, cudaStream_t s,
                      uint64_t *launches) {
    SH_LAUNCH(n_in, (sh_receive_kernel<<<grid_for(n_in), SH_THREADS, 0, s>>>(in, n_in, S, kinds, frag_extra, frag_cap, pair, pair_cap, pair_far,
                                                                            far_cap, mate_of, fm, fm_cap)));
    return 0;
}
int launch_sh_fm_pack(const uint64_t *fm, uint32_t n, E128 *out, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n, (sh_fm_pack_kernel<<<grid_for(n), SH_THREADS, 0, s>>>(fm, n, out)));
    return 0;
}
int launch_sh_fm_unpack(const E128 *in, uint32_t n, uint64_t *fm, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n, (sh_fm_unpack_kernel<<<grid_for(n), SH_THREADS, 0, s>>>(in, n, fm)));
    return 0;
}
int launch_sh_apply_marks(const uint32_t *marks, uint64_t n_marks, uint64_t idx_base, uint64_t n, uint8_t *dup, cudaStream_t s,
                          uint64_t *launches) {
    SH_LAUNCH(n_marks, (sh_apply_marks_kernel<<<grid_for(n_marks), SH_THREADS, 0, s>>>(marks, n_marks, idx_base, n, dup)));
    return 0;
}

int launch_sh_keylen(const uint8_t *rec, const uint64_t *off, uint64_t n, const uint64_t *hk, uint32_t *out_max, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n, (sh_keylen_kernel<<<grid_for(n), SH_THREADS, 0, s>>>(rec, off, n, hk, out_max)));
    return 0;
}
int launch_sh_gather2(const uint32_t *list, uint32_t n_list, const E128 *frag, const uint64_t *hk, const uint8_t *rec, const uint64_t *off, uint8_t *out,
                      uint32_t stride, uint64_t *hashes, uint32_t *err, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_list, (sh_gather2_kernel<<<grid_for(n_list), SH_THREADS, 0, s>>>(list, n_list, frag, hk, rec, off, out, stride, hashes, err)));
    return 0;
}
int launch_sh_set_build(const uint64_t *h, uint64_t n, unsigned long long *set, uint64_t n_slots, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n, (sh_set_build_kernel<<<grid_for(n), SH_THREADS, 0, s>>>(h, n, set, n_slots)));
    return 0;
}
int launch_sh_probe_table(MateSlot *table, uint64_t n_slots, const unsigned long long *set, uint64_t set_slots, E128 *pair, E128 *pair_far, uint32_t *list2,
                          uint32_t *counters, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_slots, (sh_probe_table_kernel<<<grid_for(n_slots), SH_THREADS, 0, s>>>(table, n_slots, set, set_slots, pair, pair_far, list2, counters)));
    return 0;
}
int launch_sh_probe_pairs(const uint64_t *pair_hk, uint32_t n_pairs, const unsigned long long *set, uint64_t set_slots, E128 *list, int far,
                          const uint32_t *mate_of, const ShardParams &S, uint32_t *list2, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_pairs, (sh_probe_pairs_kernel<<<grid_for(n_pairs), SH_THREADS, 0, s>>>(pair_hk, n_pairs, set, set_slots, list, far, mate_of, S, list2)));
    return 0;
}
int launch_sh_wbuild2(const uint8_t *w, uint32_t stride, uint32_t n_w, const KeyLayout &L, E128 *out, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n_w, (sh_wbuild2_kernel<<<grid_for(n_w), SH_THREADS, 0, s>>>(w, stride, n_w, L, out)));
    return 0;
}
int launch_sh_replay2(const E128 *sorted, uint32_t n_w, const uint8_t *w, uint32_t stride, uint8_t *state, const ShardParams &S, E128 *pair, uint32_t pair_cap,
                      E128 *pair_far, uint32_t far_cap, uint32_t *mate_of, uint64_t *fm, uint32_t fm_cap, RouteEntry *out, uint32_t out_cap, cudaStream_t s,
                      uint64_t *launches) {
    SH_LAUNCH(n_w, (sh_replay2_kernel<<<grid_for(n_w), SH_THREADS, 0, s>>>(sorted, n_w, w, stride, state, S, pair, pair_cap, pair_far, far_cap, mate_of,
                                                                            fm, fm_cap, out, out_cap)));
    return 0;
}
int launch_sh_bucket_count(const void *items, uint64_t n, uint32_t item_bytes, int kind, const ShardParams &S, const uint64_t *bases, uint32_t *count,
                           cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n, (sh_bucket_count_kernel<<<grid_for(n), SH_THREADS, 0, s>>>((const uint8_t *) items, n, item_bytes, kind, S, bases, count)));
    return 0;
}
int launch_sh_bucket_scatter(const void *items, uint64_t n, uint32_t item_bytes, int kind, const ShardParams &S, const uint64_t *bases, const uint32_t *start,
                             uint32_t *fill, void *out, cudaStream_t s, uint64_t *launches) {
    SH_LAUNCH(n, (sh_bucket_scatter_kernel<<<grid_for(n), SH_THREADS, 0, s>>>((const uint8_t *) items, n, item_bytes, kind, S, bases, start, fill,
                                                                              (uint8_t *) out)));
    return 0;
}

}  // namespace oge
