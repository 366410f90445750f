// K1: end-building.  One pass over the device-resident BAM records.
//
// Replaces, per record, buildReadEnds and everything it calls in the reference
// (algorithms/mark_duplicates.cpp under /root/reference/openge/src):
//   eligibility            :202-205      mapped && refID != -1 && primary
//   getReferenceLength     :44-61        sum of M D N = X
//   getUnclippedStart/End  :88-129       pos - leading S/H ; alignment end + trailing S/H
//   getScore               :135-144      short sum of quality bytes >= 15 (wraps mod 2^16)
//   getLibraryId/-Name     :282-318      RG tag -> @RG -> LB (host-resolved table), "Unknown Library"
//   pairing key            :210-214      RG + ":" + name  -> 64-bit hash of exactly those bytes
//   tag walk               util/bamtools/BamAlignment.cpp:270-294, 699-786
//
// Shape: persistent CTAs of 128 threads, several per SM.  A tile = 128 consecutive records = one
// contiguous byte range; the CTA fetches it (and the tile's 129 offsets) into shared memory with
// 1-D bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier), then one thread
// parses one record out of shared memory with aligned 32-bit ld.shared reads.  Each CTA has one
// stage; the copy latency is covered by the other CTAs resident on the SM.
// HBM-bound by design: the record bytes are read once (algorithmic: core + name + cigar + quals
// + tags up to RG; the packed bases ride along in the same sectors).
// Outputs per record (coalesced): 16 B end entry, 8 B name hash, 2 B flag, and for records that enter
// the mate map a 32 B pairing-key tag (read-group code + name).
#include <stdlib.h>

#include "kernels.cuh"
#include "pairing.cuh"

namespace oge {

// ---------------------------------------------------------------- PTX helpers (mbarrier + bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// ---------------------------------------------------------------- record readers
// A record is parsed through a reader so that the same code runs on the shared-memory stage
// (32-bit shared addresses, ld.shared) and, for tiles too large for the stage, on global memory.
// Every buffer read this way carries >= 16 B of slack past its last byte.
struct SharedRd {
    uint32_t base;      // shared-space byte address of the record's block_size field
    __device__ __forceinline__ uint32_t word(uint32_t a) const {      // aligned word at shared address a
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    }
    __device__ __forceinline__ uint32_t addr(uint32_t o) const { return base + o; }
    __device__ __forceinline__ uint32_t u8(uint32_t o) const {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + o));
        return v;
    }
};

struct GlobalRd {
    const uint8_t *base;
    __device__ __forceinline__ uint32_t word(uint64_t a) const { return *reinterpret_cast<const uint32_t *>(a); }
    __device__ __forceinline__ uint64_t addr(uint32_t o) const { return (uint64_t) (uintptr_t) base + o; }
    __device__ __forceinline__ uint32_t u8(uint32_t o) const { return base[o]; }
};

// unaligned little-endian 32-bit read: two aligned words + funnel shift (branch free)
template <class Rd>
__device__ __forceinline__ uint32_t rd_u32(const Rd &r, uint32_t o) {
    auto a = r.addr(o);
    auto wa = a & ~(decltype(a)) 3;
    return __funnelshift_r(r.word(wa), r.word(wa + 4), (uint32_t) (a & 3) * 8);
}
template <class Rd>
__device__ __forceinline__ uint32_t rd_u16(const Rd &r, uint32_t o) { return rd_u32(r, o) & 0xFFFFu; }

// ---------------------------------------------------------------- pairing-key hash
// A function of the byte string RG + ":" + name only (not of how it splits into RG and name:
// the reference's map key is the concatenation, mark_duplicates.cpp:214).  Bytes are packed
// into 32-bit words in stream order and mixed with two murmur3-style lanes.
struct KeyHasher {
    uint32_t h1, h2, carry, nb, len;      // carry: nb (0..3) pending bytes in the low end
    __device__ __forceinline__ void init() {
        h1 = 0x9E3779B9u; h2 = 0x85EBCA6Bu; carry = 0; nb = 0; len = 0;
    }
    __device__ __forceinline__ void mix(uint32_t k) {
        uint32_t k1 = k * 0xCC9E2D51u;
        k1 = __funnelshift_l(k1, k1, 15) * 0x1B873593u;
        h1 ^= k1;
        h1 = __funnelshift_l(h1, h1, 13) * 5u + 0xE6546B64u;
        uint32_t k2 = k * 0x239B961Bu;
        k2 = __funnelshift_l(k2, k2, 16) * 0xAB0E9789u;
        h2 ^= k2;
        h2 = __funnelshift_l(h2, h2, 17) * 5u + 0x561CCD1Bu;
    }
    __device__ __forceinline__ void push4(uint32_t w) {      // four bytes
        mix(carry | (w << (8 * nb)));
        carry = __funnelshift_rc(w, 0u, 32 - 8 * nb);       // nb == 0 -> shift 32 -> 0
        len += 4;
    }
    __device__ __forceinline__ void push_tail(uint32_t w, uint32_t n) {      // n in 1..3, bytes above n zero
        uint32_t v = carry | (w << (8 * nb));
        if (nb + n >= 4) {
            mix(v);
            carry = __funnelshift_rc(w, 0u, 32 - 8 * nb);
            nb = nb + n - 4;
        } else {
            carry = v;
            nb += n;
        }
        len += n;
    }
    // n bytes starting at record offset o
    template <class Rd>
    __device__ __forceinline__ void push_bytes(const Rd &r, uint32_t o, uint32_t n) {
        if (n == 0) return;
        auto a = r.addr(o);
        auto wa = a & ~(decltype(a)) 3;
        const uint32_t sh = (uint32_t) (a & 3) * 8;
        uint32_t cur = r.word(wa);
        const uint32_t full = n >> 2;
        for (uint32_t j = 0; j < full; j++) {
            wa += 4;
            uint32_t nxt = r.word(wa);
            push4(__funnelshift_r(cur, nxt, sh));
            cur = nxt;
        }
        const uint32_t rem = n & 3;
        if (rem) {
            uint32_t nxt = r.word(wa + 4);
            push_tail(__funnelshift_r(cur, nxt, sh) & ((1u << (8 * rem)) - 1), rem);
        }
    }
    __device__ __forceinline__ uint64_t finish() {
        if (nb) mix(carry);
        h1 ^= len; h2 ^= len;
        h1 += h2; h2 += h1;
        h1 ^= h1 >> 16; h1 *= 0x85EBCA6Bu; h1 ^= h1 >> 13; h1 *= 0xC2B2AE35u; h1 ^= h1 >> 16;
        h2 ^= h2 >> 16; h2 *= 0x85EBCA6Bu; h2 ^= h2 >> 13; h2 *= 0xC2B2AE35u; h2 ^= h2 >> 16;
        h1 += h2; h2 += h1;
        if (h2 == 0) h2 = 1;      // the low word is the key word of the in-CTA join table (0 = empty); h == 0 = "not in the mate map"
        return ((uint64_t) h1 << 32) | h2;
    }
};

// ---------------------------------------------------------------- RG tag walk
// FindTag + SkipToNextTag for "RG" (BamAlignment.cpp:270-294, 699-786); the value is taken as
// a NUL-terminated string whatever its type code (GetTag<string>, BamAlignment.h:575-606).
// `t0` = record offset of the tag block, n = its length.  Returns the value's record offset
// (or -1) and its length bounded by the record end.
template <class Rd>
__device__ int find_rg(const Rd &r, uint32_t t0, uint32_t n, uint32_t *len) {
    uint32_t parsed = 0;
    *len = 0;
    while (parsed < n) {
        if (n - parsed < 3) return -1;
        uint32_t t = rd_u32(r, t0 + parsed);      // name[0], name[1], type, first value byte
        uint32_t type = (t >> 16) & 0xFF;
        parsed += 3;
        if ((t & 0xFFFF) == ((uint32_t) 'R' | ((uint32_t) 'G' << 8))) {
            uint32_t l = 0;
            while (parsed + l < n && r.u8(t0 + parsed + l)) l++;
            *len = l;
            return (int) (t0 + parsed);
        }
        switch (type) {
            case 'A': case 'c': case 'C': parsed += 1; break;
            case 's': case 'S': parsed += 2; break;
            case 'f': case 'i': case 'I': parsed += 4; break;
            case 'Z': case 'H':
                while (parsed < n && r.u8(t0 + parsed)) parsed++;
                parsed++;
                break;
            case 'B': {
                if (parsed + 5 > n) return -1;
                uint32_t at = r.u8(t0 + parsed);
                int32_t cnt = (int32_t) rd_u32(r, t0 + parsed + 1);
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
            default: return -1;      // includes type == 0
        }
        if (parsed >= n) return -1;
        if (r.u8(t0 + parsed) == 0) return -1;
    }
    return -1;
}

// read-group code: index into the host-resolved @RG table, RGC_ABSENT for no tag / empty value,
// RGC_UNKNOWN for a value the header does not list.  `tb`/`to` point at the table (shared-memory
// copy when it is small, see the kernel).
template <class Rd>
__device__ __forceinline__ uint32_t rg_lookup(const uint8_t *tb, const uint32_t *to, int n_rg, const Rd &r, uint32_t o, uint32_t len) {
    if (len == 0) return RGC_ABSENT;
    for (int i = 0; i < n_rg; i++) {
        uint32_t a = to[i], b = to[i + 1];
        if (b - a != len) continue;
        uint32_t j = 0;
        while (j < len && tb[a + j] == r.u8(o + j)) j++;
        if (j == len) return (uint32_t) i;
    }
    return RGC_UNKNOWN;
}

// bytes >= 15 of w, summed (SWAR: no per-byte compare instruction exists on sm_100)
__device__ __forceinline__ uint32_t score4(uint32_t w, uint32_t acc) {
    uint32_t msb = (((w | 0x80808080u) - 0x0F0F0F0Fu) | w) & 0x80808080u;      // bit 7 of each byte: byte >= 15
    uint32_t m = (msb - (msb >> 7)) | msb;                                     // 0xFF per selected byte
    return __dp4a(w & m, 0x01010101u, acc);
}

struct RgSmem {
    const uint8_t *bytes;
    const uint32_t *off;
    const int16_t *lib;
};

// ---------------------------------------------------------------- one record
// What the fused kernel keeps of a record for the in-CTA join (the key arrays hk[] / tag[] are not written there).
struct EndOut {
    E128 ent;
    uint64_t hk;
    uint32_t rgc, l_name;
};

template <class Rd, bool KEYS>
__device__ __forceinline__ void build_end(const EndbuildParams &P, const RgSmem &rgt, const Rd &r, uint32_t rec_len, uint64_t i,
                                          uint32_t &err, bool &is_frag, bool &is_pe, bool &is_unpaired, EndOut &out) {
    E128 ent;
    ent.lo = ent.hi = ~0ull;
    uint64_t hk = 0;
    uint32_t rgc = RGC_ABSENT, flag = 0;
    is_frag = is_pe = is_unpaired = false;

    bool ok = rec_len >= 36;
    uint32_t l_name = 0, n_cig = 0, l_seq = 0, o_cig = 0, o_qual = 0, o_tags = 0;
    if (ok) {
        uint32_t block_size = rd_u32(r, 0);
        l_name = r.u8(12);
        uint32_t cf = rd_u32(r, 16);
        n_cig = cf & 0xFFFF;
        flag = cf >> 16;
        l_seq = rd_u32(r, 20);
        o_cig = 36 + l_name;
        uint64_t oq = (uint64_t) o_cig + 4ull * n_cig + (((uint64_t) l_seq + 1) >> 1);
        uint64_t ot = oq + l_seq;
        ok = (uint64_t) block_size + 4 == rec_len && ot <= rec_len;
        o_qual = (uint32_t) oq;
        o_tags = (uint32_t) ot;
    }
    if (!ok) {
        err |= DEV_ERR_BAD_RECORD;
    } else {
        int32_t ref = (int32_t) rd_u32(r, 4);
        if (!(flag & 0x4) && ref != -1 && !(flag & 0x100)) {          // mark_duplicates.cpp:202-205
            int32_t pos = (int32_t) rd_u32(r, 8);
            bool rev = (flag & 0x10) != 0;
            // ---- CIGAR: reference length + leading / trailing clip runs in one walk
            uint32_t lead = 0, trail = 0, reflen = 0;
            bool in_lead = true;
            for (uint32_t c = 0; c < n_cig; c++) {
                uint32_t w = rd_u32(r, o_cig + 4 * c), op = w & 0xF, len = w >> 4;
                if (op == 4 || op == 5) {
                    if (in_lead) lead += len;
                    trail += len;
                } else {
                    in_lead = false;
                    trail = 0;
                    if ((0x18Du >> op) & 1) reflen += len;      // M D N = X  (ops 0 2 3 7 8)
                }
            }
            int32_t coord = rev ? (int32_t) ((uint32_t) pos + reflen - 1u + trail) : (int32_t) ((uint32_t) pos - lead);

            // ---- score: sum of quality bytes >= 15, mod 2^16 (short accumulator in the reference)
            uint32_t s0 = 0, s1 = 0;
            {
                auto a = r.addr(o_qual);
                auto wa = a & ~(decltype(a)) 3;
                const uint32_t sh = (uint32_t) (a & 3) * 8;
                const uint32_t full = l_seq >> 2, rem = l_seq & 3;
                uint32_t cur = r.word(wa);
                uint32_t j = 0;
                for (; j + 4 <= full; j += 4) {
                    uint32_t w1 = r.word(wa + 4), w2 = r.word(wa + 8), w3 = r.word(wa + 12), w4 = r.word(wa + 16);
                    s0 = score4(__funnelshift_r(cur, w1, sh), s0);
                    s1 = score4(__funnelshift_r(w1, w2, sh), s1);
                    s0 = score4(__funnelshift_r(w2, w3, sh), s0);
                    s1 = score4(__funnelshift_r(w3, w4, sh), s1);
                    cur = w4;
                    wa += 16;
                }
                for (; j < full; j++) {
                    uint32_t nxt = r.word(wa + 4);
                    s0 = score4(__funnelshift_r(cur, nxt, sh), s0);
                    cur = nxt;
                    wa += 4;
                }
                if (rem) {
                    uint32_t nxt = r.word(wa + 4);
                    s1 = score4(__funnelshift_r(cur, nxt, sh) & ((1u << (8 * rem)) - 1), s1);
                }
            }
            uint32_t score = (s0 + s1) & 0xFFFFu;

            // ---- RG -> read-group code -> library
            uint32_t rg_len;
            int rg_at = find_rg(r, o_tags, rec_len - o_tags, &rg_len);
            uint32_t rg_o = rg_at >= 0 ? (uint32_t) rg_at : o_tags;
            if (rg_at < 0) rg_len = 0;
            rgc = rg_lookup(rgt.bytes, rgt.off, P.rg.n, r, rg_o, rg_len);
            uint32_t lib = rgc < RGC_UNKNOWN ? (uint32_t) rgt.lib[rgc] : (uint32_t) P.rg.unknown_lib;

            bool pe = (flag & 0x1) && !(flag & 0x8);                  // :157, :209
            int32_t mate_ref = (int32_t) rd_u32(r, 24);
            bool paired = pe && mate_ref != -1;                       // ReadEnds::isPaired, picard_structures.h:54

            // ---- pack the key
            const KeyLayout &L = P.kl;
            long long sc = (long long) coord + L.coord_bias;
            if (ref < 0 || (uint32_t) ref >= (1u << L.ref_bits) || sc < 0 || sc >= (1ll << L.coord_bits) || lib >= L.lib_invalid) {
                err |= DEV_ERR_KEY_RANGE;
            } else {
                uint64_t idx = P.idx_base + i;
                ent.lo = score;
                ent.hi = 0;
                bits_or(ent, L.f_idx, idx);
                bits_or(ent, L.f_paired, ((uint64_t) (paired ? 1 : 0)) | ((uint64_t) (rev ? 2 : 0)));      // f_orient = f_paired + 1
                bits_or(ent, L.f_coord, (uint64_t) sc);
                bits_or(ent, L.f_ref, (uint64_t) ref | ((uint64_t) lib << L.ref_bits));                  // f_lib = f_ref + ref_bits
                is_frag = true;
                is_unpaired = !paired;      // only these can be marked by the fragment pass (mark_duplicates.cpp:379, 517-538)
                if (pe) {
                    KeyHasher h;
                    h.init();
                    h.push_bytes(r, rg_o, rg_len);
                    h.push_tail(':', 1);
                    h.push_bytes(r, 36, l_name ? l_name - 1 : 0);
                    hk = h.finish();
                    is_pe = true;
                    if (KEYS) {
                    // compact key copy for the join
                    const uint32_t nlen = l_name ? l_name - 1 : 0;
                    uint32_t w[8];
                    w[0] = rgc | (l_name << 16) | (nlen ? (r.u8(36) << 24) : 0u);
#pragma unroll
                    for (int k = 1; k < 8; k++) {
                        const uint32_t first = 4 * k - 3;      // name bytes [first, first + 4)
                        uint32_t v = 0;
                        if (first < nlen) {
                            v = rd_u32(r, 36 + first);
                            const uint32_t have = nlen - first;
                            if (have < 4) v &= (1u << (8 * have)) - 1;
                        }
                        w[k] = v;
                    }
                    uint4 *t = reinterpret_cast<uint4 *>(P.tag + i);
                    t[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    t[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
            }
        }
    }
    reinterpret_cast<ulonglong2 *>(P.frag)[i] = make_ulonglong2(ent.lo, ent.hi);
    if (KEYS) P.hk[i] = hk;
    P.flag_in[i] = (uint16_t) flag;
    out.ent = ent;
    out.hk = hk;
    out.rgc = rgc;
    out.l_name = l_name;
}

// ---------------------------------------------------------------- the kernel
constexpr uint32_t EB_OFF_BYTES = ((EB_THREADS + 2) * 8 + 15) & ~15u;      // 129 offsets, rounded to 16 B
constexpr int EB_RG_SMEM_BYTES = 1024, EB_RG_SMEM_N = 32;

__global__ void __launch_bounds__(EB_THREADS) endbuild_kernel(EndbuildParams P, uint32_t stage_cap, uint32_t n_tiles) {
    extern __shared__ __align__(128) uint8_t smem[];      // [offsets: EB_OFF_BYTES][records: stage_cap]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint64_t s_a0;
    __shared__ uint32_t s_direct;
    __shared__ uint8_t s_rg_bytes[EB_RG_SMEM_BYTES];
    __shared__ uint32_t s_rg_off[EB_RG_SMEM_N + 1];
    __shared__ int16_t s_rg_lib[EB_RG_SMEM_N];

    const int tid = threadIdx.x;
    const uint64_t *s_off = reinterpret_cast<const uint64_t *>(smem);
    const uint32_t stage_addr = smem_u32(smem + EB_OFF_BYTES);

    // small @RG tables are served from shared memory
    RgSmem rgt{P.rg.bytes, P.rg.off, P.rg.lib};
    {
        uint32_t total = P.rg.n > 0 && P.rg.n <= EB_RG_SMEM_N ? P.rg.off[P.rg.n] : 0xFFFFFFFFu;
        if (total <= (uint32_t) EB_RG_SMEM_BYTES) {
            for (uint32_t j = tid; j < total; j += EB_THREADS) s_rg_bytes[j] = P.rg.bytes[j];
            for (int j = tid; j <= P.rg.n; j += EB_THREADS) s_rg_off[j] = P.rg.off[j];
            for (int j = tid; j < P.rg.n; j += EB_THREADS) s_rg_lib[j] = P.rg.lib[j];
            rgt = RgSmem{s_rg_bytes, s_rg_off, s_rg_lib};
        }
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_end = [&](uint32_t tile) {
        uint64_t r1 = (uint64_t) tile * EB_THREADS + EB_THREADS;
        return r1 < P.n ? r1 : P.n;
    };
    // thread 0: start the copies of a tile whose byte range is [b0, b1)
    auto issue = [&](uint32_t tile, uint64_t b0, uint64_t b1) {
        uint64_t a0 = b0 & ~15ull;
        uint64_t len = (b1 - a0 + 15) & ~15ull;
        uint64_t r0 = (uint64_t) tile * EB_THREADS;
        uint32_t off_bytes = (uint32_t) (((tile_end(tile) - r0 + 1) * 8 + 15) & ~15ull);
        s_a0 = a0;
        if (b1 >= b0 && len + 16 <= stage_cap) {
            s_direct = 0;
            mbar_expect_tx(&bar, (uint32_t) len + off_bytes);
            bulk_g2s(stage_addr, P.rec + a0, (uint32_t) len, &bar);
        } else {
            s_direct = 1;      // parsed straight from global memory
            mbar_expect_tx(&bar, off_bytes);
        }
        bulk_g2s(smem_u32(smem), P.off + r0, off_bytes, &bar);
    };

    if (tid == 0 && blockIdx.x < n_tiles) issue(blockIdx.x, P.off[(uint64_t) blockIdx.x * EB_THREADS], P.off[tile_end(blockIdx.x)]);

    uint32_t err = 0, n_frag = 0, n_pe = 0, n_unp = 0, k = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, k++) {
        // byte range of this CTA's next tile: requested now, needed after the parse
        const uint32_t nt = tile + gridDim.x;
        uint64_t nb0 = 0, nb1 = 0;
        if (tid == 0 && nt < n_tiles) {
            nb0 = __ldg(P.off + (uint64_t) nt * EB_THREADS);
            nb1 = __ldg(P.off + tile_end(nt));
        }
        mbar_wait(&bar, k & 1);
        const uint64_t r = (uint64_t) tile * EB_THREADS + tid;
        if (r < P.n) {
            const uint64_t o0 = s_off[tid], o1 = s_off[tid + 1];
            bool f = false, pe = false, unp = false;
            EndOut eo;
            if (o1 < o0 || o1 - o0 > 0xFFFFFFFFull) {
                err |= DEV_ERR_BAD_RECORD;
                reinterpret_cast<ulonglong2 *>(P.frag)[r] = make_ulonglong2(~0ull, ~0ull);
                P.hk[r] = 0;
                P.flag_in[r] = 0;
            } else if (s_direct) {
                GlobalRd rd{P.rec + o0};
                build_end<GlobalRd, true>(P, rgt, rd, (uint32_t) (o1 - o0), r, err, f, pe, unp, eo);
            } else {
                SharedRd rd{stage_addr + (uint32_t) (o0 - s_a0)};
                build_end<SharedRd, true>(P, rgt, rd, (uint32_t) (o1 - o0), r, err, f, pe, unp, eo);
            }
            n_frag += f;
            n_pe += pe;
            n_unp += unp;
        }
        __syncthreads();      // everyone is done with the stage
        if (tid == 0 && nt < n_tiles) issue(nt, nb0, nb1);
    }

    // counters: one atomic per warp
    for (int o = 16; o; o >>= 1) {
        n_frag += __shfl_xor_sync(0xFFFFFFFFu, n_frag, o);
        n_pe += __shfl_xor_sync(0xFFFFFFFFu, n_pe, o);
        n_unp += __shfl_xor_sync(0xFFFFFFFFu, n_unp, o);
        err |= __shfl_xor_sync(0xFFFFFFFFu, err, o);
    }
    if ((tid & 31) == 0) {
        if (n_frag) atomicAdd(&P.counters[CNT_FRAG], n_frag);
        if (n_unp) atomicAdd(&P.counters[CNT_UNPAIRED], n_unp);
        if (n_pe) atomicAdd(&P.counters[CNT_PAIR_ELIGIBLE], n_pe);
        if (err) atomicOr(&P.counters[CNT_ERR], err);
    }
}


// ================================================================ K1 fused with the in-CTA mate join
// Same parse as above, but every CTA walks a CONTIGUOUS range of tiles and keeps the records whose mate it
// has not seen yet in a shared-memory table (8-way buckets: a key word per way, 24 bytes of payload).
// Per tile, after the parse:
//   phase 1   every map-eligible record scans its bucket: its key word is there -> FOUND that way; else it
//             claims the first empty way with a shared-memory CAS (a failed CAS that returns its own key
//             word is a FOUND as well: the mate got there first) and writes its payload -> INSERTED
//   phase 2   a FOUND record takes the entry (first taker only) and confirms the match by comparing
//             read-group code, name length and the name bytes of the two records in global memory (L2:
//             both were just read); it builds the pair entry exactly as the global join does (flip rule
//             mark_duplicates.cpp:226-243, orientation :169-178, score :245), appends it and frees the way.
// Everything else -- bucket full, a second taker (name seen three times at once), equal hash but
// different key bytes, entries nobody came for within LJ_HORIZON records or by the end of the CTA's
// range -- goes on the `left` list with its hash and tag, for the global join.  Pairing here is by
// arrival, not by file order, which is only right for names seen exactly twice; that is what the check
// pass afterwards establishes (a name with any record on the `left` list has its pairs retracted and
// is replayed in file order by the exact path).  Within one CTA an entry is always the most recent
// unmatched sighting, so names without leftovers and without a double take were paired as the
// reference's toggle map (util/picard_structures.h:87-96) pairs them.
struct __align__(8) LjEntry {
    uint32_t h_hi, meta, ord, cnt;      // meta = rgcode | l_name << 16; cnt = takers so far
    uint64_t recoff;                    // byte offset of the record
};
static_assert(sizeof(LjEntry) + 4 == LJ_ENTRY_BYTES, "LJ_ENTRY_BYTES");

struct LjLeft {      // where a leftover goes (by value: a reference to the kernel parameters would put them on the stack)
    uint32_t *counter, *left;
    uint64_t *hk;
    NameTag *tag;
    const uint8_t *rec;
};

__device__ __noinline__ void lj_leftover(LjLeft L, uint32_t ord, uint64_t h, uint32_t meta, uint64_t recoff) {
    const uint32_t act = __activemask(), lane = threadIdx.x & 31;
    const int leader = __ffs(act) - 1;
    uint32_t base = 0;
    if ((int) lane == leader) base = atomicAdd(L.counter, (uint32_t) __popc(act));
    base = __shfl_sync(act, base, leader);
    const uint32_t pos = base + __popc(act & ((1u << lane) - 1));
    L.left[pos] = ord;
    L.hk[ord] = h;
    make_tag_global(L.rec + recoff, meta & 0xFFFFu, (meta >> 16) & 0xFFu, L.tag + ord);
}

__global__ void __launch_bounds__(EB_THREADS, 4) endbuild_join_kernel(EndbuildParams P, LocalJoinParams J, uint32_t stage_cap,
                                                                      uint32_t n_tiles, uint32_t dbg) {
    extern __shared__ __align__(128) uint8_t smem[];      // [offsets: EB_OFF_BYTES][records: stage_cap][key words][payloads]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint64_t s_a0;
    __shared__ uint32_t s_direct;
    __shared__ uint32_t s_cur_base, s_cur_left, s_next_base, s_used;      // pair-list positions reserved by this CTA
    __shared__ uint8_t s_rg_bytes[EB_RG_SMEM_BYTES];
    __shared__ uint32_t s_rg_off[EB_RG_SMEM_N + 1];
    __shared__ int16_t s_rg_lib[EB_RG_SMEM_N];

    const int tid = threadIdx.x;
    const uint64_t *s_off = reinterpret_cast<const uint64_t *>(smem);
    const uint32_t stage_addr = smem_u32(smem + EB_OFF_BYTES);
    const uint32_t n_entries = J.n_buckets * LJ_WAYS;
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem + EB_OFF_BYTES + stage_cap);
    LjEntry *pay = reinterpret_cast<LjEntry *>(keys + n_entries);

    RgSmem rgt{P.rg.bytes, P.rg.off, P.rg.lib};
    {
        uint32_t total = P.rg.n > 0 && P.rg.n <= EB_RG_SMEM_N ? P.rg.off[P.rg.n] : 0xFFFFFFFFu;
        if (total <= (uint32_t) EB_RG_SMEM_BYTES) {
            for (uint32_t j = tid; j < total; j += EB_THREADS) s_rg_bytes[j] = P.rg.bytes[j];
            for (int j = tid; j <= P.rg.n; j += EB_THREADS) s_rg_off[j] = P.rg.off[j];
            for (int j = tid; j < P.rg.n; j += EB_THREADS) s_rg_lib[j] = P.rg.lib[j];
            rgt = RgSmem{s_rg_bytes, s_rg_off, s_rg_lib};
        }
    }
    for (uint32_t j = tid; j < n_entries; j += EB_THREADS) keys[j] = 0;
    const uint32_t t0 = blockIdx.x * J.tiles_per_cta;
    const uint32_t t1 = min(t0 + J.tiles_per_cta, n_tiles);
    if (t0 >= t1) return;
    uint32_t pending = 0;
    bool has_pending = false;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t base = atomicAdd(&P.counters[CNT_PAIRS], 2 * LJ_PAIR_BLOCK);
        s_cur_base = base;
        s_cur_left = LJ_PAIR_BLOCK;
        s_next_base = base + LJ_PAIR_BLOCK;
        s_used = 0;
    }
    __syncthreads();

    auto tile_end = [&](uint32_t tile) {
        uint64_t r1 = (uint64_t) tile * EB_THREADS + EB_THREADS;
        return r1 < P.n ? r1 : P.n;
    };
    auto issue = [&](uint32_t tile, uint64_t b0, uint64_t b1) {
        uint64_t a0 = b0 & ~15ull;
        uint64_t len = (b1 - a0 + 15) & ~15ull;
        uint64_t r0 = (uint64_t) tile * EB_THREADS;
        uint32_t off_bytes = (uint32_t) (((tile_end(tile) - r0 + 1) * 8 + 15) & ~15ull);
        s_a0 = a0;
        if (b1 >= b0 && len + 16 <= stage_cap) {
            s_direct = 0;
            mbar_expect_tx(&bar, (uint32_t) len + off_bytes);
            bulk_g2s(stage_addr, P.rec + a0, (uint32_t) len, &bar);
        } else {
            s_direct = 1;
            mbar_expect_tx(&bar, off_bytes);
        }
        bulk_g2s(smem_u32(smem), P.off + r0, off_bytes, &bar);
    };
    if (tid == 0) issue(t0, P.off[(uint64_t) t0 * EB_THREADS], P.off[tile_end(t0)]);

    const uint32_t lane = tid & 31;
    const LjLeft LL{P.counters + CNT_LEFT, J.left, P.hk, P.tag, P.rec};
    uint32_t err = 0, n_frag = 0, n_pe = 0, n_unp = 0, k = 0;
    for (uint32_t tile = t0; tile < t1; tile++, k++) {
        const uint32_t nt = tile + 1;
        uint64_t nb0 = 0, nb1 = 0;
        if (tid == 0 && nt < t1) {
            nb0 = __ldg(P.off + (uint64_t) nt * EB_THREADS);
            nb1 = __ldg(P.off + tile_end(nt));
        }
        mbar_wait(&bar, k & 1);
        const uint64_t r = (uint64_t) tile * EB_THREADS + tid;
        bool pe = false;
        EndOut eo;
        eo.hk = 0;
        uint64_t o0 = 0;
        if (r < P.n) {
            o0 = s_off[tid];
            const uint64_t o1 = s_off[tid + 1];
            bool f = false, unp = false;
            if (o1 < o0 || o1 - o0 > 0xFFFFFFFFull) {
                err |= DEV_ERR_BAD_RECORD;
                reinterpret_cast<ulonglong2 *>(P.frag)[r] = make_ulonglong2(~0ull, ~0ull);
                P.flag_in[r] = 0;
            } else if (s_direct) {
                GlobalRd rd{P.rec + o0};
                build_end<GlobalRd, false>(P, rgt, rd, (uint32_t) (o1 - o0), r, err, f, pe, unp, eo);
            } else {
                SharedRd rd{stage_addr + (uint32_t) (o0 - s_a0)};
                build_end<SharedRd, false>(P, rgt, rd, (uint32_t) (o1 - o0), r, err, f, pe, unp, eo);
            }
            n_frag += f;
            n_pe += pe;
            n_unp += unp;
        }
        __syncthreads();      // A: the stage is free; the previous tile's phase 2 (and sweep) are complete
        if (tid == 0) {
            if (nt < t1) issue(nt, nb0, nb1);
            // pair-list positions: the block asked for one tile ago has arrived by now
            if (has_pending) { s_next_base = pending; has_pending = false; }
            const uint32_t u = s_used;
            if (u) {
                if (u >= s_cur_left) {
                    const uint32_t over = u - s_cur_left;
                    s_cur_base = s_next_base + over;
                    s_cur_left = LJ_PAIR_BLOCK - over;      // >= EB_THREADS: enough for this tile whatever happens
                    pending = atomicAdd(&P.counters[CNT_PAIRS], LJ_PAIR_BLOCK);
                    has_pending = true;
                } else {
                    s_cur_base += u;
                    s_cur_left -= u;
                }
                s_used = 0;
            }
        }

#ifdef OGE_TESTING
        if (dbg & 1) continue;      // measurement only (wrong results): parse alone
#endif
        // ---- phase 1: find the mate's entry, or leave one
        const uint32_t ord = (uint32_t) r;
        const uint32_t k32 = (uint32_t) eo.hk, h_hi = (uint32_t) (eo.hk >> 32);
        const uint32_t meta = eo.rgc | (eo.l_name << 16);
        int way = -1;             // FOUND: the way; INSERTED: -2; overflow: -3
        uint32_t slot0 = 0;
        if (pe) {
            slot0 = __umulhi(h_hi, J.n_buckets) * LJ_WAYS;
            const uint4 ka = *reinterpret_cast<const uint4 *>(keys + slot0), kb = *reinterpret_cast<const uint4 *>(keys + slot0 + 4);
            const uint32_t kk[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
            int first_empty = -1;
#pragma unroll
            for (int w = 7; w >= 0; w--) {
                if (kk[w] == k32) way = w;
                if (kk[w] == 0) first_empty = w;
            }
            if (way < 0) {
                way = -3;
                for (int w = first_empty; w >= 0 && w < LJ_WAYS; w++) {
                    const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(keys + slot0 + w);
                    if (cur == k32) { way = w; break; }
                    if (cur != 0) continue;
                    const uint32_t old = atomicCAS(keys + slot0 + w, 0u, k32);
                    if (old == 0) {
                        LjEntry &e = pay[slot0 + w];
                        e.h_hi = h_hi; e.meta = meta; e.ord = ord; e.cnt = 0; e.recoff = o0;
                        way = -2;
                        break;
                    }
                    if (old == k32) { way = w; break; }
                }
            }
        }
        __syncthreads();      // B: payloads of this tile's insertions are visible

#ifdef OGE_TESTING
        if (dbg & 2) continue;      // measurement only (wrong results): no phase 2
#endif
        // ---- phase 2: take the entry, confirm the names, emit the pair
        if (pe && way != -2) {
            bool left_self = true;
            if (way >= 0) {
                LjEntry &e = pay[slot0 + way];
                if (e.h_hi == h_hi && atomicAdd(&e.cnt, 1u) == 0) {
                    const uint32_t other = e.ord, ometa = e.meta;
                    const uint64_t ooff = e.recoff;
#ifdef OGE_TESTING
                    const E128 oent = (dbg & 8) ? eo.ent : ld_frag(P.frag + other);
                    const uint32_t nlen = (dbg & 4) ? 0 : (eo.l_name ? eo.l_name - 1 : 0);
#else
                    const E128 oent = ld_frag(P.frag + other);
                    const uint32_t nlen = eo.l_name ? eo.l_name - 1 : 0;
#endif
                    const bool same = ometa == meta && eo.rgc != RGC_UNKNOWN && bytes_equal_global(P.rec + o0 + 36, P.rec + ooff + 36, nlen);
                    keys[slot0 + way] = 0;      // the way is free again (nobody reads key words before the next barrier)
                    if (same) {
                        left_self = false;
                        uint32_t i1, i2;
                        bool far;
                        const bool self_first = ord < other;      // file order
                        const E128 ent = make_pair_entry(P.kl, self_first ? eo.ent : oent, self_first ? oent : eo.ent, &i1, &i2, P.idx_base, &far);
                        if (far) {
                            const uint32_t at = atomicAdd(&P.counters[CNT_PAIRS_FAR], 1u);
                            if (at < J.far_cap) {
                                reinterpret_cast<ulonglong2 *>(J.pair_far)[at] = make_ulonglong2(ent.lo, ent.hi);
                                J.pair_far_hk[at] = eo.hk;
                            } else err |= DEV_ERR_CAPACITY;
                        } else {
                            const uint32_t j = atomicAdd(&s_used, 1u);
                            const uint32_t at = j < s_cur_left ? s_cur_base + j : s_next_base + (j - s_cur_left);
                            if (at < J.pair_cap) {
                                reinterpret_cast<ulonglong2 *>(J.pair)[at] = make_ulonglong2(ent.lo, ent.hi);
                                J.pair_hk[at] = eo.hk;
                            } else err |= DEV_ERR_CAPACITY;
                        }
                        J.mate_of[i1] = (uint32_t) (i2 + P.idx_base);
                    } else {
                        // one hash, two keys (or read groups the header does not list): both to the global join
                        lj_leftover(LL, other, ((uint64_t) e.h_hi << 32) | k32, ometa, ooff);
                    }
                }
            }
            if (left_self) lj_leftover(LL, ord, eo.hk, meta, o0);
        }

        // ---- entries nobody came for leave for the global join
        if ((k % LJ_SWEEP_TILES) == LJ_SWEEP_TILES - 1) {
            __syncthreads();
            for (uint32_t j = tid; j < n_entries; j += EB_THREADS) {
                const uint32_t kw = keys[j];
                if (kw && (uint32_t) (tile * EB_THREADS) - pay[j].ord > LJ_HORIZON) {
                    const LjEntry e = pay[j];
                    keys[j] = 0;
                    lj_leftover(LL, e.ord, ((uint64_t) e.h_hi << 32) | kw, e.meta, e.recoff);
                }
            }
        }
    }

    // ---- end of the CTA's range: whatever is still waiting leaves; unused reserved pair positions become dead entries
    __syncthreads();
    for (uint32_t j = tid; j < n_entries; j += EB_THREADS) {
        const uint32_t kw = keys[j];
        if (kw) {
            const LjEntry e = pay[j];
            lj_leftover(LL, e.ord, ((uint64_t) e.h_hi << 32) | kw, e.meta, e.recoff);
        }
    }
    if (tid == 0) {
        if (has_pending) s_next_base = pending;
        const uint32_t u = s_used;
        if (u >= s_cur_left) {      // the block after `next` was asked for only if this branch ran above: `next` is all there is
            const uint32_t over = u - s_cur_left;
            s_cur_base = s_next_base + over;
            s_cur_left = LJ_PAIR_BLOCK - over;
            s_next_base = 0xFFFFFFFFu;
        } else {
            s_cur_base += u;
            s_cur_left -= u;
        }
    }
    __syncthreads();
    {
        const uint32_t cb = s_cur_base, cl = s_cur_left, nb = s_next_base;
        const uint32_t dead = cl + (nb != 0xFFFFFFFFu ? LJ_PAIR_BLOCK : 0u);
        for (uint32_t j = tid; j < dead; j += EB_THREADS) {
            const uint32_t at = j < cl ? cb + j : nb + (j - cl);
            if (at < J.pair_cap) {
                reinterpret_cast<ulonglong2 *>(J.pair)[at] = make_ulonglong2(~0ull, ~0ull);
                J.pair_hk[at] = 0;
            }
        }
        if (tid == 0 && dead) atomicAdd(&P.counters[CNT_PAIRS_RETRACTED], dead);
    }

    for (int o = 16; o; o >>= 1) {
        n_frag += __shfl_xor_sync(0xFFFFFFFFu, n_frag, o);
        n_pe += __shfl_xor_sync(0xFFFFFFFFu, n_pe, o);
        n_unp += __shfl_xor_sync(0xFFFFFFFFu, n_unp, o);
        err |= __shfl_xor_sync(0xFFFFFFFFu, err, o);
    }
    if (lane == 0) {
        if (n_frag) atomicAdd(&P.counters[CNT_FRAG], n_frag);
        if (n_unp) atomicAdd(&P.counters[CNT_UNPAIRED], n_unp);
        if (n_pe) atomicAdd(&P.counters[CNT_PAIR_ELIGIBLE], n_pe);
        if (err) atomicOr(&P.counters[CNT_ERR], err);
    }
}


// ---------------------------------------------------------------- warp-specialised form
// The same algorithm with the two halves in different warps of a 256-thread CTA, so that the latency chains of the
// join (shared-memory atomics, the L2 round trip for the mate's name and end entry, the bit packing of the pair
// entry) overlap with the parse of the next tile instead of adding to it:
//   warps 0-3  PARSE   wait for the tile's bulk copy, parse one record per thread, hand (end entry, key hash,
//                      rgcode/l_name, record offset) to the join warps through a 4.5 KB mailbox, start the next copy
//   warps 4-7  JOIN    phase 1 / phase 2 / sweep exactly as above, one tile behind the parse warps
// Hand-over through two mbarriers (FULL: 128 parse arrivals, release; EMPTY: 128 join arrivals once the mailbox
// is in registers), group-internal synchronisation through named barriers of 128 threads.
constexpr int WS_THREADS = 2 * EB_THREADS;
constexpr uint32_t WS_MAILBOX_BYTES = EB_THREADS * (4 * 8 + 4);

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// n name bytes at a and at b equal?  All loads of a 32-byte chunk are issued before the first comparison (one L2 round
// trip per chunk instead of one per word).
__device__ __forceinline__ bool names_equal_global(const uint8_t *a, const uint8_t *b, uint32_t n) {
    for (uint32_t j = 0; j < n; j += 32) {
        const uintptr_t pa = (uintptr_t) (a + j), pb = (uintptr_t) (b + j);
        const uint32_t *wa = reinterpret_cast<const uint32_t *>(pa & ~(uintptr_t) 3), *wb = reinterpret_cast<const uint32_t *>(pb & ~(uintptr_t) 3);
        const uint32_t sa = (uint32_t) (pa & 3) * 8, sb = (uint32_t) (pb & 3) * 8;
        const uint32_t have = n - j < 32 ? n - j : 32, words = (have + 3) >> 2;
        uint32_t xa[9], xb[9];
#pragma unroll
        for (int k = 0; k < 9; k++) {
            xa[k] = (uint32_t) k <= words ? wa[k] : 0u;
            xb[k] = (uint32_t) k <= words ? wb[k] : 0u;
        }
        uint32_t diff = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t d = __funnelshift_r(xa[k], xa[k + 1], sa) ^ __funnelshift_r(xb[k], xb[k + 1], sb);
            const uint32_t left = have > 4u * k ? have - 4u * k : 0u;      // bytes of this word that count
            const uint32_t m = left >= 4 ? 0xFFFFFFFFu : ((1u << (8 * left)) - 1);
            diff |= d & m;
        }
        if (diff) return false;
    }
    return true;
}

__global__ void __launch_bounds__(WS_THREADS, 4) endbuild_join_ws_kernel(EndbuildParams P, LocalJoinParams J, uint32_t stage_cap,
                                                                         uint32_t n_tiles, uint32_t dbg) {
    extern __shared__ __align__(128) uint8_t smem[];      // [offsets][records: stage_cap][mailbox][key words][payloads]
    __shared__ __align__(8) uint64_t bar, bar_full, bar_empty;
    __shared__ uint64_t s_a0;
    __shared__ uint32_t s_direct;
    __shared__ uint32_t s_cur_base, s_cur_left, s_next_base, s_used;
    __shared__ uint8_t s_rg_bytes[EB_RG_SMEM_BYTES];
    __shared__ uint32_t s_rg_off[EB_RG_SMEM_N + 1];
    __shared__ int16_t s_rg_lib[EB_RG_SMEM_N];

    const int tid = threadIdx.x;
    const bool parser = tid < EB_THREADS;
    const int gt = tid & (EB_THREADS - 1);      // thread within its group
    const uint64_t *s_off = reinterpret_cast<const uint64_t *>(smem);
    const uint32_t stage_addr = smem_u32(smem + EB_OFF_BYTES);
    uint64_t *mb_lo = reinterpret_cast<uint64_t *>(smem + EB_OFF_BYTES + stage_cap);
    uint64_t *mb_hi = mb_lo + EB_THREADS, *mb_hk = mb_hi + EB_THREADS, *mb_off = mb_hk + EB_THREADS;
    uint32_t *mb_meta = reinterpret_cast<uint32_t *>(mb_off + EB_THREADS);
    const uint32_t n_entries = J.n_buckets * LJ_WAYS;
    uint32_t *keys = mb_meta + EB_THREADS;
    LjEntry *pay = reinterpret_cast<LjEntry *>(keys + n_entries);

    RgSmem rgt{P.rg.bytes, P.rg.off, P.rg.lib};
    {
        uint32_t total = P.rg.n > 0 && P.rg.n <= EB_RG_SMEM_N ? P.rg.off[P.rg.n] : 0xFFFFFFFFu;
        if (total <= (uint32_t) EB_RG_SMEM_BYTES) {
            for (uint32_t j = tid; j < total; j += WS_THREADS) s_rg_bytes[j] = P.rg.bytes[j];
            for (int j = tid; j <= P.rg.n; j += WS_THREADS) s_rg_off[j] = P.rg.off[j];
            for (int j = tid; j < P.rg.n; j += WS_THREADS) s_rg_lib[j] = P.rg.lib[j];
            rgt = RgSmem{s_rg_bytes, s_rg_off, s_rg_lib};
        }
    }
    for (uint32_t j = tid; j < n_entries; j += WS_THREADS) keys[j] = 0;
    const uint32_t t0 = blockIdx.x * J.tiles_per_cta;
    const uint32_t t1 = min(t0 + J.tiles_per_cta, n_tiles);
    if (t0 >= t1) return;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar_full, EB_THREADS);
        mbar_init(&bar_empty, EB_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == EB_THREADS) {
        const uint32_t base = atomicAdd(&P.counters[CNT_PAIRS], 2 * LJ_PAIR_BLOCK);
        s_cur_base = base;
        s_cur_left = LJ_PAIR_BLOCK;
        s_next_base = base + LJ_PAIR_BLOCK;
        s_used = 0;
    }
    __syncthreads();
    const uint32_t lane = tid & 31;
    uint32_t err = 0;

    if (parser) {
        // ================================================================ PARSE warps
        auto tile_end = [&](uint32_t tile) {
            uint64_t r1 = (uint64_t) tile * EB_THREADS + EB_THREADS;
            return r1 < P.n ? r1 : P.n;
        };
        auto issue = [&](uint32_t tile, uint64_t b0, uint64_t b1) {
            uint64_t a0 = b0 & ~15ull;
            uint64_t len = (b1 - a0 + 15) & ~15ull;
            uint64_t r0 = (uint64_t) tile * EB_THREADS;
            uint32_t off_bytes = (uint32_t) (((tile_end(tile) - r0 + 1) * 8 + 15) & ~15ull);
            s_a0 = a0;
            if (b1 >= b0 && len + 16 <= stage_cap) {
                s_direct = 0;
                mbar_expect_tx(&bar, (uint32_t) len + off_bytes);
                bulk_g2s(stage_addr, P.rec + a0, (uint32_t) len, &bar);
            } else {
                s_direct = 1;
                mbar_expect_tx(&bar, off_bytes);
            }
            bulk_g2s(smem_u32(smem), P.off + r0, off_bytes, &bar);
        };
        if (tid == 0) issue(t0, P.off[(uint64_t) t0 * EB_THREADS], P.off[tile_end(t0)]);
        uint32_t n_frag = 0, n_pe = 0, n_unp = 0, k = 0;
        for (uint32_t tile = t0; tile < t1; tile++, k++) {
            const uint32_t nt = tile + 1;
            uint64_t nb0 = 0, nb1 = 0;
            if (tid == 0 && nt < t1) {
                nb0 = __ldg(P.off + (uint64_t) nt * EB_THREADS);
                nb1 = __ldg(P.off + tile_end(nt));
            }
            mbar_wait(&bar, k & 1);
            const uint64_t r = (uint64_t) tile * EB_THREADS + tid;
            bool pe = false;
            EndOut eo;
            eo.hk = 0; eo.rgc = 0; eo.l_name = 0; eo.ent.lo = eo.ent.hi = ~0ull;
            uint64_t o0 = 0;
            if (r < P.n) {
                o0 = s_off[tid];
                const uint64_t o1 = s_off[tid + 1];
                bool f = false, unp = false;
                if (o1 < o0 || o1 - o0 > 0xFFFFFFFFull) {
                    err |= DEV_ERR_BAD_RECORD;
                    reinterpret_cast<ulonglong2 *>(P.frag)[r] = make_ulonglong2(~0ull, ~0ull);
                    P.flag_in[r] = 0;
                } else if (s_direct) {
                    GlobalRd rd{P.rec + o0};
                    build_end<GlobalRd, false>(P, rgt, rd, (uint32_t) (o1 - o0), r, err, f, pe, unp, eo);
                } else {
                    SharedRd rd{stage_addr + (uint32_t) (o0 - s_a0)};
                    build_end<SharedRd, false>(P, rgt, rd, (uint32_t) (o1 - o0), r, err, f, pe, unp, eo);
                }
                n_frag += f;
                n_pe += pe;
                n_unp += unp;
            }
            named_bar_sync(1, EB_THREADS);      // every parse thread is done with the stage
            if (tid == 0 && nt < t1) issue(nt, nb0, nb1);
            if (k > 0) mbar_wait(&bar_empty, (k - 1) & 1);      // the join warps have taken the previous tile out of the mailbox
            mb_lo[tid] = eo.ent.lo;
            mb_hi[tid] = eo.ent.hi;
            mb_hk[tid] = pe ? eo.hk : 0ull;
            mb_off[tid] = o0;
            mb_meta[tid] = eo.rgc | (eo.l_name << 16);
            mbar_arrive(&bar_full);             // release: the mailbox and this thread's frag[] / flag_in[] stores
        }
        for (int o = 16; o; o >>= 1) {
            n_frag += __shfl_xor_sync(0xFFFFFFFFu, n_frag, o);
            n_pe += __shfl_xor_sync(0xFFFFFFFFu, n_pe, o);
            n_unp += __shfl_xor_sync(0xFFFFFFFFu, n_unp, o);
            err |= __shfl_xor_sync(0xFFFFFFFFu, err, o);
        }
        if (lane == 0) {
            if (n_frag) atomicAdd(&P.counters[CNT_FRAG], n_frag);
            if (n_unp) atomicAdd(&P.counters[CNT_UNPAIRED], n_unp);
            if (n_pe) atomicAdd(&P.counters[CNT_PAIR_ELIGIBLE], n_pe);
            if (err) atomicOr(&P.counters[CNT_ERR], err);
        }
        return;
    }

    // ==================================================================== JOIN warps
    const LjLeft LL{P.counters + CNT_LEFT, J.left, P.hk, P.tag, P.rec};
    uint32_t pending = 0, k = 0;
    bool has_pending = false;
    for (uint32_t tile = t0; tile < t1; tile++, k++) {
        mbar_wait(&bar_full, k & 1);
        E128 ent;
        ent.lo = mb_lo[gt];
        ent.hi = mb_hi[gt];
        const uint64_t hk = mb_hk[gt], o0 = mb_off[gt];
        const uint32_t meta = mb_meta[gt];
        mbar_arrive(&bar_empty);
        if (gt == 0) {
            // pair-list positions: the block asked for one tile ago has arrived by now
            if (has_pending) { s_next_base = pending; has_pending = false; }
            const uint32_t u = s_used;
            if (u) {
                if (u >= s_cur_left) {
                    const uint32_t over = u - s_cur_left;
                    s_cur_base = s_next_base + over;
                    s_cur_left = LJ_PAIR_BLOCK - over;      // >= EB_THREADS: enough for this tile whatever happens
                    pending = atomicAdd(&P.counters[CNT_PAIRS], LJ_PAIR_BLOCK);
                    has_pending = true;
                } else {
                    s_cur_base += u;
                    s_cur_left -= u;
                }
                s_used = 0;
            }
        }
        const bool pe = hk != 0;
        const uint32_t ord = tile * EB_THREADS + gt;
        const uint32_t k32 = (uint32_t) hk, h_hi = (uint32_t) (hk >> 32);
#ifdef OGE_TESTING
        if (dbg & 1) continue;      // measurement only (wrong results): no join at all
#endif
        // ---- phase 1: find the mate's entry, or leave one
        int way = -1;             // FOUND: the way; INSERTED: -2; overflow: -3
        uint32_t slot0 = 0;
        if (pe) {
            slot0 = __umulhi(h_hi, J.n_buckets) * LJ_WAYS;
            const uint4 ka = *reinterpret_cast<const uint4 *>(keys + slot0), kb = *reinterpret_cast<const uint4 *>(keys + slot0 + 4);
            const uint32_t kk[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
            int first_empty = -1;
#pragma unroll
            for (int w = 7; w >= 0; w--) {
                if (kk[w] == k32) way = w;
                if (kk[w] == 0) first_empty = w;
            }
            if (way < 0) {
                way = -3;
                for (int w = first_empty; w >= 0 && w < LJ_WAYS; w++) {
                    const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(keys + slot0 + w);
                    if (cur == k32) { way = w; break; }
                    if (cur != 0) continue;
                    const uint32_t old = atomicCAS(keys + slot0 + w, 0u, k32);
                    if (old == 0) {
                        LjEntry &e = pay[slot0 + w];
                        e.h_hi = h_hi; e.meta = meta; e.ord = ord; e.cnt = 0; e.recoff = o0;
                        way = -2;
                        break;
                    }
                    if (old == k32) { way = w; break; }
                }
            }
        }
        named_bar_sync(2, EB_THREADS);      // payloads of this tile's insertions are visible; the block bookkeeping is done

        // ---- phase 2: take the entry, confirm the names, emit the pair
#ifdef OGE_TESTING
        if (!(dbg & 2))
#endif
        if (pe && way != -2) {
            bool left_self = true;
            if (way >= 0) {
                LjEntry &e = pay[slot0 + way];
                if (e.h_hi == h_hi && atomicAdd(&e.cnt, 1u) == 0) {
                    const uint32_t other = e.ord, ometa = e.meta;
                    const uint64_t ooff = e.recoff;
                    const E128 oent = ld_frag(P.frag + other);
                    const uint32_t l_name = (meta >> 16) & 0xFFu, nlen = l_name ? l_name - 1 : 0;
                    const bool same = ometa == meta && (meta & 0xFFFFu) != RGC_UNKNOWN && names_equal_global(P.rec + o0 + 36, P.rec + ooff + 36, nlen);
                    keys[slot0 + way] = 0;      // the way is free again (nobody reads key words before the next barrier)
                    if (same) {
                        left_self = false;
                        uint32_t i1, i2;
                        bool far;
                        const bool self_first = ord < other;      // file order
                        const E128 pent = make_pair_entry(P.kl, self_first ? ent : oent, self_first ? oent : ent, &i1, &i2, P.idx_base, &far);
                        if (far) {
                            const uint32_t at = atomicAdd(&P.counters[CNT_PAIRS_FAR], 1u);
                            if (at < J.far_cap) {
                                reinterpret_cast<ulonglong2 *>(J.pair_far)[at] = make_ulonglong2(pent.lo, pent.hi);
                                J.pair_far_hk[at] = hk;
                            } else err |= DEV_ERR_CAPACITY;
                        } else {
                            const uint32_t j = atomicAdd(&s_used, 1u);
                            const uint32_t at = j < s_cur_left ? s_cur_base + j : s_next_base + (j - s_cur_left);
                            if (at < J.pair_cap) {
                                reinterpret_cast<ulonglong2 *>(J.pair)[at] = make_ulonglong2(pent.lo, pent.hi);
                                J.pair_hk[at] = hk;
                            } else err |= DEV_ERR_CAPACITY;
                        }
                        J.mate_of[i1] = (uint32_t) (i2 + P.idx_base);
                    } else {
                        // one hash, two keys (or read groups the header does not list): both to the global join
                        lj_leftover(LL, other, ((uint64_t) e.h_hi << 32) | k32, ometa, ooff);
                    }
                }
            }
            if (left_self) lj_leftover(LL, ord, hk, meta, o0);
        }

        // ---- entries nobody came for leave for the global join
        if ((k % LJ_SWEEP_TILES) == LJ_SWEEP_TILES - 1) {
            named_bar_sync(2, EB_THREADS);
            for (uint32_t j = gt; j < n_entries; j += EB_THREADS) {
                const uint32_t kw = keys[j];
                if (kw && (uint32_t) (tile * EB_THREADS) - pay[j].ord > LJ_HORIZON) {
                    const LjEntry e = pay[j];
                    keys[j] = 0;
                    lj_leftover(LL, e.ord, ((uint64_t) e.h_hi << 32) | kw, e.meta, e.recoff);
                }
            }
        }
        named_bar_sync(2, EB_THREADS);      // phase 2 (and the sweep) are complete before the next tile's phase 1
    }

    // ---- end of the CTA's range: whatever is still waiting leaves; unused reserved pair positions become dead entries
    for (uint32_t j = gt; j < n_entries; j += EB_THREADS) {
        const uint32_t kw = keys[j];
        if (kw) {
            const LjEntry e = pay[j];
            lj_leftover(LL, e.ord, ((uint64_t) e.h_hi << 32) | kw, e.meta, e.recoff);
        }
    }
    if (gt == 0) {
        if (has_pending) s_next_base = pending;
        const uint32_t u = s_used;
        if (u >= s_cur_left) {
            const uint32_t over = u - s_cur_left;
            s_cur_base = s_next_base + over;
            s_cur_left = LJ_PAIR_BLOCK - over;
            s_next_base = 0xFFFFFFFFu;
        } else {
            s_cur_base += u;
            s_cur_left -= u;
        }
    }
    named_bar_sync(2, EB_THREADS);
    {
        const uint32_t cb = s_cur_base, cl = s_cur_left, nb = s_next_base;
        const uint32_t dead = cl + (nb != 0xFFFFFFFFu ? LJ_PAIR_BLOCK : 0u);
        for (uint32_t j = gt; j < dead; j += EB_THREADS) {
            const uint32_t at = j < cl ? cb + j : nb + (j - cl);
            if (at < J.pair_cap) {
                reinterpret_cast<ulonglong2 *>(J.pair)[at] = make_ulonglong2(~0ull, ~0ull);
                J.pair_hk[at] = 0;
            }
        }
        if (gt == 0 && dead) atomicAdd(&P.counters[CNT_PAIRS_RETRACTED], dead);
    }
    for (int o = 16; o; o >>= 1) err |= __shfl_xor_sync(0xFFFFFFFFu, err, o);
    if (lane == 0 && err) atomicOr(&P.counters[CNT_ERR], err);
}

int launch_endbuild(const EndbuildParams &P, uint32_t avg_rec_bytes, int sms, cudaStream_t stream, uint64_t *launches) {
    if (P.n == 0) return 0;
    if (P.kl.f_orient != P.kl.f_paired + 1 || P.kl.f_lib != P.kl.f_ref + P.kl.ref_bits)
        return fail_msg(-1, "endbuild: key layout must keep paired|orient and ref|lib adjacent");
    // stage sized for a typical tile + 6 % (a tile that does not fit is parsed from global memory)
    uint64_t want = ((uint64_t) avg_rec_bytes * EB_THREADS * 17 / 16 + 512 + 127) & ~127ull;
    uint32_t stage_cap = (uint32_t) (want < 8192 ? 8192 : (want > 160 * 1024 ? 160 * 1024 : want));
    size_t smem = (size_t) stage_cap + EB_OFF_BYTES;
    // per device (the attribute belongs to the current device's copy of the function), and cheap: no process-wide flag
    OGE_CUDA_TRY(cudaFuncSetAttribute(endbuild_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (200 * 1024)));
    int per_sm = (int) ((224 * 1024) / (smem + 3 * 1024));      // + static shared memory and the per-CTA reserve
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    uint64_t n_tiles = (P.n + EB_THREADS - 1) / EB_THREADS;
    uint64_t grid = (uint64_t) sms * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    endbuild_kernel<<<(uint32_t) grid, EB_THREADS, smem, stream>>>(P, stage_cap, (uint32_t) n_tiles);
    *launches += 1;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

// CTAs per SM of the fused kernel: the shared memory of an SM (228 KB, 1 KB reserved per CTA) split evenly
constexpr uint32_t LJ_SMEM_PER_SM = 228 * 1024, LJ_STATIC_SMEM = 2048;
constexpr int LJ_MAX_PER_SM = 4;

uint32_t endbuild_join_max_grid(int sms) { return (uint32_t) sms * LJ_MAX_PER_SM; }

int launch_endbuild_join(const EndbuildParams &P, LocalJoinParams J, uint32_t avg_rec_bytes, int sms, cudaStream_t stream,
                         uint64_t *launches, uint32_t *grid_out) {
    *grid_out = 0;
    if (P.n == 0) return 0;
    if (P.kl.f_orient != P.kl.f_paired + 1 || P.kl.f_lib != P.kl.f_ref + P.kl.ref_bits)
        return fail_msg(-1, "endbuild: key layout must keep paired|orient and ref|lib adjacent");
    uint64_t want = ((uint64_t) avg_rec_bytes * EB_THREADS * 17 / 16 + 512 + 127) & ~127ull;
    uint32_t stage_cap = (uint32_t) (want < 8192 ? 8192 : (want > 160 * 1024 ? 160 * 1024 : want));
    uint32_t dbg = 0;
#ifdef OGE_TESTING
    if (const char *e = getenv("OGE_LJ_DBG")) dbg = (uint32_t) atoi(e);      // measurement knobs (most give wrong results)
#endif
    const bool ws = !(dbg & 16);      // warp-specialised form (default)
    // as many CTAs per SM as leave every one of them a table of at least 256 entries next to its stage
    int per_sm = LJ_MAX_PER_SM;
    uint32_t table_bytes = 0;
    for (; per_sm >= 1; per_sm--) {
        const uint32_t budget = LJ_SMEM_PER_SM / per_sm - 1024 - LJ_STATIC_SMEM;
        const uint32_t fixed = stage_cap + EB_OFF_BYTES + (ws ? WS_MAILBOX_BYTES : 0u);
        if (budget >= fixed + 256 * LJ_ENTRY_BYTES || per_sm == 1) {
            table_bytes = budget > fixed ? budget - fixed : 0;
            break;
        }
    }
    if (per_sm < 1) per_sm = 1;
    uint32_t n_buckets = table_bytes / (LJ_ENTRY_BYTES * LJ_WAYS);
    if (n_buckets > 256) n_buckets = 256;
    if (n_buckets < 1) {      // records too large for a stage plus a table: the stage shrinks (oversized tiles are parsed from global memory)
        stage_cap = 64 * 1024;
        n_buckets = 64;
    }
    J.n_buckets = n_buckets;
    const size_t smem = (size_t) stage_cap + EB_OFF_BYTES + (ws ? WS_MAILBOX_BYTES : 0u) + (size_t) n_buckets * LJ_WAYS * LJ_ENTRY_BYTES;
    OGE_CUDA_TRY(cudaFuncSetAttribute(endbuild_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (220 * 1024)));
    OGE_CUDA_TRY(cudaFuncSetAttribute(endbuild_join_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (220 * 1024)));
    const uint64_t n_tiles = (P.n + EB_THREADS - 1) / EB_THREADS;
    // contiguous tile ranges, at least 16 tiles each (pairs across a range boundary go through the global join)
    uint64_t grid = (uint64_t) sms * per_sm;
    if (grid > (n_tiles + 15) / 16) grid = (n_tiles + 15) / 16;
    if (grid < 1) grid = 1;
    const uint64_t tpc = (n_tiles + grid - 1) / grid;
    grid = (n_tiles + tpc - 1) / tpc;
    J.tiles_per_cta = (uint32_t) tpc;
    if (ws) endbuild_join_ws_kernel<<<(uint32_t) grid, WS_THREADS, smem, stream>>>(P, J, stage_cap, (uint32_t) n_tiles, dbg);
    else endbuild_join_kernel<<<(uint32_t) grid, EB_THREADS, smem, stream>>>(P, J, stage_cap, (uint32_t) n_tiles, dbg);
    *launches += 1;
    *grid_out = (uint32_t) grid;
    OGE_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace oge
