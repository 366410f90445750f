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
// Shape: persistent CTAs; each tile of 128 consecutive records is one contiguous byte range,
// fetched into shared memory with a 1-D bulk async copy (cp.async.bulk, completion on an
// mbarrier), double buffered; one thread then parses one record out of shared memory with
// 32-bit word reads.  HBM-bound: the record bytes are read once.
// Outputs per record (coalesced): 16 B end entry, 8 B name hash, 2 B read-group code, 2 B flag.
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
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

// ---------------------------------------------------------------- unaligned reads
// 32-bit little-endian read at any byte address: two aligned words + funnel shift.
// May touch up to 3 bytes past p+4; every buffer it is used on carries >= 16 B of slack.
__device__ __forceinline__ uint32_t ldu32(const uint8_t *p) {
    uintptr_t a = (uintptr_t) p;
    const uint32_t *w = (const uint32_t *) (a & ~(uintptr_t) 3);
    uint32_t sh = (uint32_t) (a & 3) * 8;
    uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
__device__ __forceinline__ uint32_t ldu16(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8); }

// ---------------------------------------------------------------- pairing-key hash
// A function of the byte string RG + ":" + name only (not of how it splits into RG and
// name: the reference's map key is the concatenation, mark_duplicates.cpp:214).
struct KeyHasher {
    uint32_t h1, h2, buf, nb, len;
    __device__ __forceinline__ void init() {
        h1 = 0x9E3779B9u; h2 = 0x85EBCA6Bu; buf = 0; nb = 0; len = 0;
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
    // n (1..4) bytes in the low end of w; bytes above n must be zero
    __device__ __forceinline__ void push(uint32_t w, uint32_t n) {
        buf |= w << (8 * nb);
        if (nb + n >= 4) {
            mix(buf);
            buf = nb ? (w >> (8 * (4 - nb))) : 0;
            nb = nb + n - 4;
        } else {
            nb += n;
        }
        len += n;
    }
    __device__ __forceinline__ void push_bytes(const uint8_t *p, uint32_t n) {
        for (uint32_t o = 0; o < n; o += 4) {
            uint32_t w = ldu32(p + o), m = n - o;
            if (m < 4) w &= (1u << (8 * m)) - 1;
            push(w, m < 4 ? m : 4);
        }
    }
    __device__ __forceinline__ uint64_t finish() {
        if (nb) mix(buf);
        h1 ^= len; h2 ^= len;
        h1 += h2; h2 += h1;
        h1 ^= h1 >> 16; h1 *= 0x85EBCA6Bu; h1 ^= h1 >> 13; h1 *= 0xC2B2AE35u; h1 ^= h1 >> 16;
        h2 ^= h2 >> 16; h2 *= 0x85EBCA6Bu; h2 ^= h2 >> 13; h2 *= 0xC2B2AE35u; h2 ^= h2 >> 16;
        h1 += h2; h2 += h1;
        uint64_t h = ((uint64_t) h1 << 32) | h2;
        return h ? h : 1;      // 0 = "not in the mate map"
    }
};

// ---------------------------------------------------------------- RG tag walk
// FindTag + SkipToNextTag for "RG" (BamAlignment.cpp:270-294, 699-786); the value is taken as
// a NUL-terminated string whatever its type code (GetTag<string>, BamAlignment.h:575-606).
// Returns the value's offset inside `tags` (or -1) and its length bounded by the record end.
__device__ int find_rg(const uint8_t *tags, uint32_t n, uint32_t *len) {
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
            default: return -1;      // includes type == 0
        }
        if (parsed >= n) return -1;
        if (tags[parsed] == 0) return -1;
    }
    return -1;
}

// read-group code: index into the host-resolved @RG table, RGC_ABSENT for no tag / empty
// value, RGC_UNKNOWN for a value the header does not list
__device__ __forceinline__ uint32_t rg_lookup(const RgTable &t, const uint8_t *rg, uint32_t len) {
    if (len == 0) return RGC_ABSENT;
    for (int i = 0; i < t.n; i++) {
        uint32_t a = t.off[i], b = t.off[i + 1];
        if (b - a != len) continue;
        uint32_t j = 0;
        while (j < len && t.bytes[a + j] == rg[j]) j++;
        if (j == len) return (uint32_t) i;
    }
    return RGC_UNKNOWN;
}

// ---------------------------------------------------------------- one record
// p points at block_size; rec_len = bytes up to the next record.
__device__ __forceinline__ void build_end(const EndbuildParams &P, const uint8_t *p, uint32_t rec_len, uint64_t i,
                                          uint32_t &err, bool &is_frag, bool &is_pe) {
    E128 ent;
    ent.lo = ent.hi = ~0ull;
    uint64_t hk = 0;
    uint32_t rgc = RGC_ABSENT, flag = 0;
    is_frag = is_pe = false;

    bool ok = rec_len >= 36;
    uint32_t block_size = 0, l_name = 0, n_cig = 0, l_seq = 0;
    uint32_t o_cig = 0, o_qual = 0, o_tags = 0;
    if (ok) {
        block_size = ldu32(p);
        l_name = p[12];
        n_cig = ldu16(p + 16);
        flag = ldu16(p + 18);
        l_seq = ldu32(p + 20);
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
        int32_t ref = (int32_t) ldu32(p + 4);
        if (!(flag & 0x4) && ref != -1 && !(flag & 0x100)) {          // mark_duplicates.cpp:202-205
            int32_t pos = (int32_t) ldu32(p + 8);
            bool rev = (flag & 0x10) != 0;
            // ---- CIGAR: reference length + leading / trailing clip runs in one walk
            uint32_t lead = 0, trail = 0, reflen = 0;
            bool in_lead = true;
            const uint8_t *cg = p + o_cig;
            for (uint32_t c = 0; c < n_cig; c++) {
                uint32_t w = ldu32(cg + 4 * c), op = w & 0xF, len = w >> 4;
                if (op == 4 || op == 5) {
                    if (in_lead) lead += len;
                    trail += len;
                } else {
                    in_lead = false;
                    trail = 0;
                    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += len;
                }
            }
            int32_t coord = rev ? (int32_t) ((uint32_t) pos + reflen - 1u + trail) : (int32_t) ((uint32_t) pos - lead);

            // ---- score: sum of quality bytes >= 15, mod 2^16 (short accumulator in the reference)
            uint32_t sum = 0;
            {
                const uint8_t *q = p + o_qual;
                uintptr_t a = (uintptr_t) q;
                const uint32_t *wp = (const uint32_t *) (a & ~(uintptr_t) 3);
                uint32_t sh = (uint32_t) (a & 3) * 8;
                uint32_t nw = (l_seq + 3) >> 2;
                uint32_t cur = wp[0];
                for (uint32_t j = 0; j < nw; j++) {
                    uint32_t nxt = wp[j + 1];
                    uint32_t w = __funnelshift_r(cur, nxt, sh);
                    cur = nxt;
                    uint32_t left = l_seq - 4 * j;
                    if (left < 4) w &= (1u << (8 * left)) - 1;
                    uint32_t m = __vcmpgeu4(w, 0x0F0F0F0Fu);
                    sum = __dp4a(w & m, 0x01010101u, sum);
                }
            }
            uint32_t score = sum & 0xFFFFu;

            // ---- RG -> read-group code -> library
            uint32_t rg_len;
            const uint8_t *tags = p + o_tags;
            int rg_at = find_rg(tags, rec_len - o_tags, &rg_len);
            const uint8_t *rg = rg_at >= 0 ? tags + rg_at : tags;
            if (rg_at < 0) rg_len = 0;
            rgc = rg_lookup(P.rg, rg, rg_len);
            uint32_t lib = rgc < RGC_UNKNOWN ? (uint32_t) P.rg.lib[rgc] : (uint32_t) P.rg.unknown_lib;

            bool pe = (flag & 0x1) && !(flag & 0x8);                  // :157, :209
            int32_t mate_ref = (int32_t) ldu32(p + 24);
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
                bits_or(ent, L.f_paired, paired ? 1 : 0);
                bits_or(ent, L.f_orient, rev ? 1 : 0);
                bits_or(ent, L.f_coord, (uint64_t) sc);
                bits_or(ent, L.f_ref, (uint64_t) ref);
                bits_or(ent, L.f_lib, lib);
                is_frag = true;
                if (pe) {
                    KeyHasher h;
                    h.init();
                    h.push_bytes(rg, rg_len);
                    h.push(':', 1);
                    h.push_bytes(p + 36, l_name ? l_name - 1 : 0);
                    hk = h.finish();
                    is_pe = true;
                }
            }
        }
    }
    reinterpret_cast<ulonglong2 *>(P.frag)[i] = make_ulonglong2(ent.lo, ent.hi);
    P.hk[i] = hk;
    P.rgcode[i] = (uint16_t) rgc;
    P.flag_in[i] = (uint16_t) flag;
}

// ---------------------------------------------------------------- the kernel
struct StageMeta {
    uint64_t a0;      // byte offset (in the record buffer) of shared-memory byte 0
    uint32_t direct;  // tile did not fit the stage: parse straight from global memory
    uint32_t pad;
};

__global__ void __launch_bounds__(EB_THREADS) endbuild_kernel(EndbuildParams P, uint32_t stage_cap, uint32_t n_tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[EB_STAGES];
    __shared__ StageMeta meta[EB_STAGES];

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < EB_STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_end = [&](uint32_t tile) {
        uint64_t r1 = (uint64_t) tile * EB_THREADS + EB_THREADS;
        return r1 < P.n ? r1 : P.n;
    };
    auto issue = [&](int s, uint64_t b0, uint64_t b1) {      // thread 0 only; [b0, b1) = the tile's bytes
        uint64_t a0 = b0 & ~15ull;
        uint64_t len = (b1 - a0 + 15) & ~15ull;
        meta[s].a0 = a0;
        if (len + 16 <= stage_cap && b1 >= b0) {
            meta[s].direct = 0;
            mbar_expect_tx(&bars[s], (uint32_t) len);
            bulk_g2s(smem + (size_t) s * stage_cap, P.rec + a0, (uint32_t) len, &bars[s]);
        } else {
            meta[s].direct = 1;
            mbar_expect_tx(&bars[s], 0);
        }
    };

    if (tid == 0) {
        for (int s = 0; s < EB_STAGES; s++) {
            uint32_t t = blockIdx.x + s * gridDim.x;
            if (t < n_tiles) issue(s, P.off[(uint64_t) t * EB_THREADS], P.off[tile_end(t)]);
        }
    }
    __syncthreads();      // meta[] visible

    uint32_t err = 0, n_frag = 0, n_pe = 0;
    uint32_t k = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, k++) {
        const int s = k % EB_STAGES;
        const uint32_t parity = (k / EB_STAGES) & 1;
        // byte range of the tile that will refill this stage: loaded now, used after the parse
        const uint32_t nt = tile + EB_STAGES * gridDim.x;
        uint64_t nb0 = 0, nb1 = 0;
        if (tid == 0 && nt < n_tiles) {
            nb0 = P.off[(uint64_t) nt * EB_THREADS];
            nb1 = P.off[tile_end(nt)];
        }
        uint64_t r = (uint64_t) tile * EB_THREADS + tid;
        uint64_t o0 = 0, o1 = 0;
        if (r < P.n) {
            o0 = P.off[r];
            o1 = P.off[r + 1];
        }
        mbar_wait(&bars[s], parity);
        const StageMeta m = meta[s];
        if (r < P.n) {
            bool f = false, pe = false;
            if (o1 < o0 || o1 - o0 > 0xFFFFFFFFull) {
                err |= DEV_ERR_BAD_RECORD;
                reinterpret_cast<ulonglong2 *>(P.frag)[r] = make_ulonglong2(~0ull, ~0ull);
                P.hk[r] = 0;
                P.rgcode[r] = (uint16_t) RGC_ABSENT;
                P.flag_in[r] = 0;
            } else {
                const uint8_t *p = m.direct ? P.rec + o0 : smem + (size_t) s * stage_cap + (o0 - m.a0);
                build_end(P, p, (uint32_t) (o1 - o0), r, err, f, pe);
            }
            n_frag += f;
            n_pe += pe;
        }
        __syncthreads();      // everyone is done with stage s (and has read meta[s])
        // refill; meta[s] is next read EB_STAGES (>= 2) iterations on, behind a later barrier
        if (tid == 0 && nt < n_tiles) issue(s, nb0, nb1);
    }

    // counters: one atomic per warp
    for (int o = 16; o; o >>= 1) {
        n_frag += __shfl_xor_sync(0xFFFFFFFFu, n_frag, o);
        n_pe += __shfl_xor_sync(0xFFFFFFFFu, n_pe, o);
        err |= __shfl_xor_sync(0xFFFFFFFFu, err, o);
    }
    if ((tid & 31) == 0) {
        if (n_frag) atomicAdd(&P.counters[CNT_FRAG], n_frag);
        if (n_pe) atomicAdd(&P.counters[CNT_PAIR_ELIGIBLE], n_pe);
        if (err) atomicOr(&P.counters[CNT_ERR], err);
    }
}

int launch_endbuild(const EndbuildParams &P, uint32_t avg_rec_bytes, int sms, cudaStream_t stream, uint64_t *launches) {
    if (P.n == 0) return 0;
    // stage sized for a typical tile + 25 % (tiles that do not fit are parsed from global memory)
    uint64_t want = ((uint64_t) avg_rec_bytes * EB_THREADS * 5 / 4 + 1024 + 127) & ~127ull;
    uint32_t stage_cap = (uint32_t) (want < 16384 ? 16384 : (want > 100 * 1024 ? 100 * 1024 : want));
    size_t smem = (size_t) stage_cap * EB_STAGES;
    static size_t configured = 0;
    if (smem > configured) {
        OGE_CUDA_TRY(cudaFuncSetAttribute(endbuild_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (200 * 1024)));
        configured = 200 * 1024;
    }
    int per_sm = (int) ((220 * 1024) / (smem + 2048));
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
