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
#include "kernels.cuh"

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
        if (h2 == 0) h2 = 1;      // the low word is the key word of the windowed join's table (0 = empty way); h == 0 = "not in the mate map"
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
    return __dp4a(w, msb >> 7, acc);                                           // unsigned bytes of w times 0/1 selectors
}

struct RgSmem {
    const uint8_t *bytes;
    const uint32_t *off;
    const int16_t *lib;
};

// ---------------------------------------------------------------- one record
template <class Rd>
__device__ __forceinline__ void build_end(const EndbuildParams &P, const RgSmem &rgt, const Rd &r, uint32_t rec_len, uint64_t i,
                                          uint32_t &err, bool &is_frag, bool &is_pe, bool &is_unpaired) {
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
                if (P.route_out) {
                    const uint64_t packed = ((uint64_t) ref << L.coord_bits) | (uint64_t) sc;
                    if (packed < P.own_lo || packed >= P.own_hi) {
                        const uint32_t at = atomicAdd(&P.counters[CNT_ROUTE], 1u);
                        if (at < P.route_cap) {
                            RouteEntry re;
                            re.e = ent; re.idx2 = 0; re.kind = 0; re.rsv = 0;
                            reinterpret_cast<RouteEntry *>(P.route_out)[at] = re;
                        }
                    }
                }
                is_unpaired = !paired;      // only these can be marked by the fragment pass (mark_duplicates.cpp:379, 517-538)
                if (pe) {
                    KeyHasher h;
                    h.init();
                    h.push_bytes(r, rg_o, rg_len);
                    h.push_tail(':', 1);
                    h.push_bytes(r, 36, l_name ? l_name - 1 : 0);
                    hk = h.finish();
                    is_pe = true;
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
    reinterpret_cast<ulonglong2 *>(P.frag)[i] = make_ulonglong2(ent.lo, ent.hi);
    P.hk[i] = hk;
    P.flag_in[i] = (uint16_t) flag;
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
            if (o1 < o0 || o1 - o0 > 0xFFFFFFFFull) {
                err |= DEV_ERR_BAD_RECORD;
                reinterpret_cast<ulonglong2 *>(P.frag)[r] = make_ulonglong2(~0ull, ~0ull);
                P.hk[r] = 0;
                P.flag_in[r] = 0;
            } else if (s_direct) {
                GlobalRd rd{P.rec + o0};
                build_end(P, rgt, rd, (uint32_t) (o1 - o0), r, err, f, pe, unp);
            } else {
                SharedRd rd{stage_addr + (uint32_t) (o0 - s_a0)};
                build_end(P, rgt, rd, (uint32_t) (o1 - o0), r, err, f, pe, unp);
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

}  // namespace oge
