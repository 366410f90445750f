// Internal: the context object shared by the single-GPU path (dedup_api.cu) and the range-sharded
// path (shard_api.cu).  Not part of the C ABI.
#pragma once
#include <algorithm>
#include <vector>

#include "kernels.cuh"
#include "oge_gpu_dedup.h"
#include "radix_sort.cuh"

namespace oge {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;      // elements
    int reserve(size_t n, bool keep, cudaStream_t s) {
        if (n <= cap) return 0;
        // grow with slack: sizes that follow data-dependent (and race-dependent) counts must not reallocate on every run.  The
        // FIRST allocation of a small buffer gets the slack as well: a count that came out a little higher on the second or
        // third run used to put a cudaFree + cudaMalloc (3 ms and more) into that run (bench: one slow step out of five).  The
        // big buffers (records, per-record arrays) are sized from the input and stay exact.
        const bool small_buf = n * sizeof(T) < ((size_t) 256 << 20);
        size_t want = keep && cap ? std::max(n, cap + cap / 2) : (cap || small_buf ? n + n / 8 + 4096 : n);
        T *q = nullptr;
        cudaError_t e = cudaMalloc((void **) &q, want * sizeof(T) + 256);
        if (e != cudaSuccess && want > n) {
            cudaGetLastError();
            want = n;
            e = cudaMalloc((void **) &q, want * sizeof(T) + 256);
        }
        if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc", __FILE__, __LINE__);
        if (keep && p && cap) {
            e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { cudaFree(q); return fail_cuda(e, "grow copy", __FILE__, __LINE__); }
        }
        if (p) cudaFree(p);
        p = q;
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// ---- multi-GPU range sharding
struct ShardState {
    bool on = false;
    uint64_t global_n = 0;
    std::vector<uint64_t> bases;          // world + 1 record ordinals
    std::vector<uint64_t> split_keys;     // world - 1 packed (ref << coord_bits | biased coord): first key of ranks 1..
    DevBuf<uint64_t> d_split, d_bases, pub_hash;
    DevBuf<unsigned long long> hset;      // hashes other ranks published
    DevBuf<uint8_t> pub_raw, pub_send, froute_send, proute_send, oroute_send, marks_send;      // lists ordered by destination rank
    DevBuf<uint32_t> bk;                  // per-destination counters of the bucketing
    // the step driven from C++ over NCCL (shard_nccl.cu): communicator, count staging, receive buffers
    void *nccl_comm = nullptr;
    DevBuf<uint64_t> x_cnt;
    DevBuf<uint8_t> r_pub, r_pub2, r_froute, r_hash, r_proute, r_oroute, r_marks;
    uint64_t x_bytes = 0, x_calls = 0;
    uint32_t k1_route_cap = 0;            // room K1 has for boundary fragment ends (0: the sweep does it)
    uint64_t own_lo = 0, own_hi = ~0ull;  // this rank's key range, packed
    uint32_t entry_bytes = 0;             // size of a published entry, agreed by all ranks
    uint32_t n_loc = 0, n_loc_far = 0;    // pairs the windowed join settled (their hashes are kept by list position)
    DevBuf<PubEntry> pub, pub2;           // published entries, rounds 1 and 2
    DevBuf<RouteEntry> route, froute;
    DevBuf<uint32_t> marks, marks_frag, pub_list;
    DevBuf<uint8_t> scratch2;             // sort scratch of the side stream
    cudaEvent_t ev_main = nullptr, ev_far = nullptr, ev_side[3] = {nullptr, nullptr, nullptr};
    bool frag_busy = false, frag_ran = false;
    int frag_mode = 2;                    // 0 nothing, 1 reduced fragment pass, 2 every fragment end
    uint64_t n_unpaired = 0, ucap = 0, uset_slots = 0;
    int side_pass_used = 0;
    uint64_t side_pass_bytes = 0, n_froute_all = 0;
    DevBuf<uint64_t> fm;                  // foreign mates: (idx1 << 32 | idx2), sorted by idx1
    DevBuf<E128> fm_sort, w_sort, w_sort2;
    uint64_t n_frag = 0, n_pe = 0, n_pairs = 0, n_retracted = 0, n_slots = 0, n_fm = 0, n_frag_total = 0, n_w = 0, n_far = 0, n_far_dead = 0;
    int phase = 0;
};

}  // namespace oge

using namespace oge;      // internal header: only this library's .cu files include it

struct oge_gpu_dedup_ctx {
    oge_gpu_dedup_config cfg;
    int sms = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr, side_stream = nullptr;
    cudaStream_t inflate_stream = nullptr;      // push_bgzf: engine submissions / inflate kernels only, never a host-to-device copy

    cudaEvent_t copy_done = nullptr;
    cudaEvent_t ev[10];
    cudaEvent_t ev_piece[3] = {nullptr, nullptr, nullptr};      // push_bgzf: stream hand-overs per piece (no timing)
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};               // push_bgzf from pageable memory: a staging buffer's upload has left it
    uint8_t *h_stage[2] = {nullptr, nullptr};                    // ... the two pinned staging buffers (allocated on first use)
    uint64_t h_stage_bytes = 0;
    // sharded path: phase clocks are resolved lazily (no host sync per phase)
    static constexpr int N_CLK = 48;
    cudaEvent_t clk_ev[2 * N_CLK];
    float *clk_slot[N_CLK];
    int clk_used = 0;
    cudaEvent_t pass_ev[2 * 48];      // profile_events: one pair per radix-sort pass launch
    // profile_events: pairs around the other kernels of a run; k_slot[i] = which OGE_K_* figure pair i adds to
    static constexpr int N_KEV = 24;
    cudaEvent_t k_ev[2 * N_KEV];
    int k_slot[N_KEV];
    int k_used = 0;
    void k_begin(int slot, cudaStream_t s) {
        if (cfg.profile_events && k_used < N_KEV) { k_slot[k_used] = slot; cudaEventRecord(k_ev[2 * k_used], s); }
    }
    void k_end(cudaStream_t s) {
        if (cfg.profile_events && k_used < N_KEV) { cudaEventRecord(k_ev[2 * k_used + 1], s); k_used++; }
    }

    // resident input
    DevBuf<uint8_t> rec;
    DevBuf<uint64_t> off;
    uint64_t n = 0, rec_bytes = 0;
    uint64_t rec_lead = 0;      // records start at rec.p + rec_lead (non-zero after push_bgzf: the BAM header sits in front)
    uint8_t *recs() const { return rec.p + rec_lead; }

    // push_bgzf: the compressed file and its block table on the device
    DevBuf<uint8_t> zcomp;
    DevBuf<uint64_t> zoff;
    DevBuf<uint32_t> zcs;
    DevBuf<uint64_t> frame_w;      // oge_gpu_dedup_frame: per-chunk entry / exit / count / base
    DevBuf<uint32_t> frame_wb;
    // oge_gpu_dedup_deflate: the finished BGZF members of the output
    DevBuf<uint8_t> zfile;
    uint64_t zfile_bytes = 0;

    // read-group table
    DevBuf<uint8_t> rg_bytes;
    DevBuf<uint32_t> rg_off;
    DevBuf<int16_t> rg_lib;
    int n_rg = 0, n_libs = 1;
    int16_t unknown_lib = 1;

    // work arrays
    DevBuf<E128> frag, sortbuf, pair, pair2, pairf, pairf2;      // pair = near pairs, pairf = far pairs
    DevBuf<E128> ufrag, ufrag2;                                  // reduced fragment pass: the entries that can matter
    DevBuf<unsigned long long> uset;                             // keys of the unpaired ends
    DevBuf<uint64_t> hk, pair_hk, pairf_hk;                      // pair_hk: key hash of the pairs formed inside the CTAs of the fused end-build
    DevBuf<uint32_t> left, couple_count;                         // windowed join: records handed to the global join; couples per CTA
    DevBuf<uint4> couples;                                       // windowed join: (taker, entry, hash) of every couple matched inside a CTA
    DevBuf<E128> cplx_sort;                                      // ping-pong buffer of the exact path's sort
    DevBuf<uint16_t> flag_in, flag_out;
    DevBuf<NameTag> tag;
    DevBuf<uint8_t> dup, scratch, cplx_state;
    DevBuf<uint32_t> mate_of, counters, cplx_slots;
    DevBuf<MateSlot> table;
    uint32_t *h_counters = nullptr;      // pinned
    uint64_t h_counters_k1_unpaired = 0;

    // coordinate sort in front of the run (coordsort.cu)
    DevBuf<uint32_t> cs_u32, perm;
    DevBuf<uint64_t> cs_bsum, off2;
    DevBuf<uint8_t> rec2;
    uint64_t sort_stats[3] = {0, 0, 0};      // tied records, refinement rounds, kernel launches of the last sort
    float ms_sort_records = 0;

    KeyLayout kl;
    bool ran = false;
    bool sorted = false;      // oge_gpu_dedup_sort has run on the resident records (pull may follow without a run)
    oge::ShardState sh;
    oge_gpu_dedup_stats stats;
};


namespace oge {

struct JoinStage {      // what join_stage leaves behind (dedup_api.cu)
    bool fused;
    uint64_t n_frag, n_pe, n_unpaired, n_left, n_pairs, n_far, n_retracted, n_far_retracted, n_cplx, n_slots;
    uint32_t n_loc, n_loc_far;
};
int join_stage(oge_gpu_dedup_ctx *c, bool replay_locally, JoinStage *out, uint64_t *launches);
int compute_layout(oge_gpu_dedup_ctx *c, KeyLayout *L);
RgTable rg_table(oge_gpu_dedup_ctx *c);
int ensure_work(oge_gpu_dedup_ctx *c);
float ms_between(cudaEvent_t a, cudaEvent_t b);
int check_endbuild_errors(oge_gpu_dedup_ctx *c);

}  // namespace oge
