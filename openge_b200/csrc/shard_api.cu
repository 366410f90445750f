// C ABI of the range-sharded path (include/oge_gpu_dedup.h, oge_gpu_shard_*): the device phases one
// rank runs between the exchanges of DESIGN.md section 6.  The host (openge_b200/sharded.py, or any
// MPI/NCCL program) moves the small lists; records, end entries, sorts and selects stay on the rank.
//
//   setup   ranges: global record count, first ordinal of every rank, first (refID, pos) of ranks 1..
//   begin   K1 + local mate join; publishes the records whose name was not seen exactly twice here
//   probe   (all ranks' round-1 entries) retracts local couples of names published elsewhere -> round 2
//   replay  (all entries of both rounds) the sequential toggle over the published set; keeps owned pairs
//   route   end entries whose key range belongs to another rank leave
//   finish  (all ranks' routed entries) K3 + K4 on what this rank owns; marks for other ranks' records
//   apply   (all ranks' marks) K5
#include <stdlib.h>
#include <string.h>

#include "ctx.cuh"

namespace oge {
int launch_sh_singletons(const MateSlot *table, uint64_t n_slots, uint32_t *list, uint32_t *counters, cudaStream_t s, uint64_t *launches);
int launch_sh_complex(const E128 *cplx, uint32_t n_cplx, uint32_t *list, uint32_t *counters, cudaStream_t s, uint64_t *launches);
int launch_sh_gather(const uint32_t *list, uint32_t n_list, const E128 *frag, const uint64_t *hk, const NameTag *tag, PubEntry *out,
                     cudaStream_t s, uint64_t *launches);
int launch_sh_probe(const PubEntry *pub, uint64_t n_pub, const ShardParams &S, MateSlot *table, uint64_t n_slots, E128 *pair, E128 *pair_far,
                    uint32_t *list2, cudaStream_t s, uint64_t *launches);
int launch_sh_wbuild(const PubEntry *w, uint32_t n_w, const KeyLayout &L, E128 *out, cudaStream_t s, uint64_t *launches);
int launch_sh_replay(const E128 *sorted, uint32_t n_w, const PubEntry *w, uint8_t *state, const ShardParams &S, E128 *pair, uint32_t pair_cap,
                     E128 *pair_far, uint32_t far_cap, uint32_t *mate_of, uint64_t *fm, uint32_t fm_cap, const RgTable &rg, cudaStream_t s,
                     uint64_t *launches);
int launch_sh_route(E128 *ents, uint64_t n_ents, int kind, const ShardParams &S, const uint32_t *mate_of, const uint64_t *fm, uint32_t n_fm,
                    RouteEntry *out, uint32_t out_cap, int dry, cudaStream_t s, uint64_t *launches);
int launch_sh_receive(const RouteEntry *in, uint64_t n_in, const ShardParams &S, uint32_t kinds, E128 *frag_extra, uint32_t frag_cap, E128 *pair,
                      uint32_t pair_cap, E128 *pair_far, uint32_t far_cap, uint32_t *mate_of, uint64_t *fm, uint32_t fm_cap, cudaStream_t s,
                      uint64_t *launches);
int launch_sh_fm_pack(const uint64_t *fm, uint32_t n, E128 *out, cudaStream_t s, uint64_t *launches);
int launch_sh_fm_unpack(const E128 *in, uint32_t n, uint64_t *fm, cudaStream_t s, uint64_t *launches);
int launch_sh_apply_marks(const uint32_t *marks, uint64_t n_marks, uint64_t idx_base, uint64_t n, uint8_t *dup, cudaStream_t s,
                          uint64_t *launches);
int launch_sh_keylen(const uint8_t *rec, const uint64_t *off, uint64_t n, const uint64_t *hk, uint32_t *out_max, cudaStream_t s, uint64_t *launches);
int launch_sh_gather2(const uint32_t *list, uint32_t n_list, const E128 *frag, const uint64_t *hk, const uint8_t *rec, const uint64_t *off, uint8_t *out,
                      uint32_t stride, uint64_t *hashes, uint32_t *err, cudaStream_t s, uint64_t *launches);
int launch_sh_set_build(const uint64_t *h, uint64_t n, unsigned long long *set, uint64_t n_slots, cudaStream_t s, uint64_t *launches);
int launch_sh_probe_table(MateSlot *table, uint64_t n_slots, const unsigned long long *set, uint64_t set_slots, E128 *pair, E128 *pair_far, uint32_t *list2,
                          uint32_t *counters, cudaStream_t s, uint64_t *launches);
int launch_sh_probe_pairs(const uint64_t *pair_hk, uint32_t n_pairs, const unsigned long long *set, uint64_t set_slots, E128 *list, int far,
                          const uint32_t *mate_of, const ShardParams &S, uint32_t *list2, cudaStream_t s, uint64_t *launches);
int launch_sh_wbuild2(const uint8_t *w, uint32_t stride, uint32_t n_w, const KeyLayout &L, E128 *out, cudaStream_t s, uint64_t *launches);
int launch_sh_replay2(const E128 *sorted, uint32_t n_w, const uint8_t *w, uint32_t stride, uint8_t *state, const ShardParams &S, E128 *pair, uint32_t pair_cap,
                      E128 *pair_far, uint32_t far_cap, uint32_t *mate_of, uint64_t *fm, uint32_t fm_cap, RouteEntry *out, uint32_t out_cap, cudaStream_t s,
                      uint64_t *launches);
int launch_sh_bucket_count(const void *items, uint64_t n, uint32_t item_bytes, int kind, const ShardParams &S, const uint64_t *bases, uint32_t *count,
                           cudaStream_t s, uint64_t *launches);
int launch_sh_bucket_scatter(const void *items, uint64_t n, uint32_t item_bytes, int kind, const ShardParams &S, const uint64_t *bases, const uint32_t *start,
                             uint32_t *fill, void *out, cudaStream_t s, uint64_t *launches);
}  // namespace oge

namespace {

ShardParams shard_params(oge_gpu_dedup_ctx *c) {
    ShardParams S;
    S.split = c->sh.d_split.p;
    S.world = c->cfg.world;
    S.rank = c->cfg.rank;
    S.idx_base = c->cfg.index_base;
    S.n = c->n;
    S.kl = c->kl;
    S.counters = c->counters.p;
    return S;
}

int read_counters(oge_gpu_dedup_ctx *c) {
    OGE_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, c->stream));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int zero_counter(oge_gpu_dedup_ctx *c, int which) {
    OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p + which, 0, 4, c->stream));
    return 0;
}

// every phase is bracketed by events on the main stream; they are resolved (and added to the stage times and to
// stats.ms_total) at the next point where the host has synchronised anyway -- no sync of their own
struct PhaseClock {
    oge_gpu_dedup_ctx *c;
    int k;
    PhaseClock(oge_gpu_dedup_ctx *ctx, float *stage) : c(ctx), k(-1) {
        if (c->clk_used < oge_gpu_dedup_ctx::N_CLK) {
            k = c->clk_used++;
            c->clk_slot[k] = stage;
            cudaEventRecord(c->clk_ev[2 * k], c->stream);
        }
    }
    void stop() {
        if (k >= 0) cudaEventRecord(c->clk_ev[2 * k + 1], c->stream);
    }
};

// call after a stream synchronisation
void resolve_clocks(oge_gpu_dedup_ctx *c) {
    for (int k = 0; k < c->clk_used; k++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->clk_ev[2 * k], c->clk_ev[2 * k + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
        c->stats.ms_total += ms;
        if (c->clk_slot[k]) *c->clk_slot[k] += ms;
    }
    c->clk_used = 0;
}

int need_phase(oge_gpu_dedup_ctx *c, int phase, const char *name) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "%s: null context", name);
    if (!c->sh.on) return fail_msg(OGE_ERR_STATE, "%s: call oge_gpu_shard_setup first", name);
    if (c->sh.phase != phase) return fail_msg(OGE_ERR_STATE, "%s: phases run in the order begin, probe, replay, route, finish, apply", name);
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    return 0;
}

}  // namespace

extern "C" {

int oge_gpu_shard_setup(oge_gpu_dedup_ctx *c, uint64_t global_n, const uint64_t *bases, const int32_t *split_ref, const int32_t *split_pos) {
    if (!c || !bases) return fail_msg(OGE_ERR_INVALID_ARG, "shard_setup: null argument");
    const int world = c->cfg.world, rank = c->cfg.rank;
    if (world < 1 || rank < 0 || rank >= world) return fail_msg(OGE_ERR_INVALID_ARG, "shard_setup: rank %d of %d", rank, world);
    if (world > 1 && (!split_ref || !split_pos)) return fail_msg(OGE_ERR_INVALID_ARG, "shard_setup: split points missing");
    if (global_n > (1ull << 32)) return fail_msg(OGE_ERR_TOO_LARGE, "shard_setup: more than 2^32 records over all ranks");
    if (bases[rank] != c->cfg.index_base) return fail_msg(OGE_ERR_INVALID_ARG, "shard_setup: bases[rank] differs from the context's index_base");
    for (int r = 0; r < world; r++)
        if (bases[r] > bases[r + 1]) return fail_msg(OGE_ERR_INVALID_ARG, "shard_setup: bases must not decrease");
    if (bases[world] != global_n) return fail_msg(OGE_ERR_INVALID_ARG, "shard_setup: bases[world] must equal the global record count");
    ShardState &sh = c->sh;
    sh.on = true;
    sh.global_n = global_n;
    sh.bases.assign(bases, bases + world + 1);
    sh.split_keys.assign(2 * (size_t) (world - 1), 0);      // raw (ref, pos) until the layout is known
    for (int r = 0; r + 1 < world; r++) {
        sh.split_keys[2 * r] = (uint64_t) (uint32_t) split_ref[r];
        sh.split_keys[2 * r + 1] = (uint64_t) (uint32_t) split_pos[r];
    }
    sh.phase = 0;
    return OGE_OK;
}

// sweep `ents` for entries owned by other ranks into sh.route (appending behind `have` stored entries)
static int route_sweep(oge_gpu_dedup_ctx *c, int n_lists, E128 *const *lists, const uint64_t *counts, const int *kinds, int mode,
                       uint64_t *n_out, uint64_t *launches) {
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    const ShardParams S = shard_params(c);
    int rc;
    uint64_t total = 0;
    for (int i = 0; i < n_lists; i++) total += counts[i];
    {
#ifdef OGE_TESTING
        const char *e = getenv("OGE_ROUTE_CAP");      // test hook: force the second sweep
#else
        const char *e = nullptr;
#endif
        const uint64_t want = e && *e ? (uint64_t) atoll(e) : std::max<uint64_t>(1u << 16, total / 64);
        if (e && *e) sh.route.release();
        if ((rc = sh.route.reserve(std::max<uint64_t>(want, 1), false, s))) return rc;
    }
    if ((rc = zero_counter(c, CNT_ROUTE)) || (rc = zero_counter(c, CNT_SCRATCH0)) || (rc = zero_counter(c, CNT_SCRATCH1))) return rc;
    *n_out = 0;
    for (int sweep = 0; sweep < 2; sweep++) {      // entries that found no room stay in place: a second sweep collects them
        const uint32_t cap = (uint32_t) sh.route.cap;
        for (int i = 0; i < n_lists; i++)
            if ((rc = launch_sh_route(lists[i], counts[i], kinds[i], S, c->mate_of.p, sh.fm.p, 0, sh.route.p, cap, mode, s, launches))) return rc;
        if ((rc = read_counters(c))) return rc;
        *n_out = c->h_counters[CNT_ROUTE];
        if (*n_out <= cap) break;
        if (sweep == 1) return fail_msg(OGE_ERR_STATE, "shard route: entry count changed between sweeps");
        if ((rc = sh.route.reserve(*n_out, mode == 0, s))) return rc;
        // moving sweep: continue behind what is stored; copying sweep: nothing left the lists, start over
        const uint32_t restart = mode == 0 ? cap : 0u;
        OGE_CUDA_TRY(cudaMemcpyAsync(c->counters.p + CNT_ROUTE, &restart, 4, cudaMemcpyHostToDevice, s));
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
        if (mode != 0 && ((rc = zero_counter(c, CNT_SCRATCH0)) || (rc = zero_counter(c, CNT_SCRATCH1)))) return rc;
    }
    return 0;
}

__global__ void sh_fit_counter_kernel(uint32_t *dst, const uint32_t *counter, uint32_t cap) { *dst = *counter <= cap ? *counter : 0u; }

__global__ void sh_add_counter_kernel(uint32_t *dst, uint32_t base, const uint32_t *counter, uint32_t cap) {
    uint32_t v = base + min(*counter, cap);
    *dst = v;
}


// What a rank sends, ordered by destination rank (kind 0: published entries by name owner, 1: routed end entries by key
// owner, 2: marks by record owner): counts[d] items for rank d, back to back in `out`.  One host round trip (the counts
// are what the exchange needs as its split sizes anyway).
static int bucket_by_destination(oge_gpu_dedup_ctx *c, const void *items, uint64_t n, uint32_t item_bytes, int kind, DevBuf<uint8_t> &out,
                                 uint64_t *counts, uint64_t *launches) {
    ShardState &sh = c->sh;
    cudaStream_t s = c->stream;
    const int W = c->cfg.world;
    for (int d = 0; d < W; d++) counts[d] = 0;
    if (n == 0) return 0;
    int rc;
    if ((rc = sh.bk.reserve(3 * (size_t) W, false, s)) || (rc = out.reserve(n * item_bytes, false, s))) return rc;
    OGE_CUDA_TRY(cudaMemsetAsync(sh.bk.p, 0, 3 * (size_t) W * 4, s));
    const ShardParams S = shard_params(c);
    if ((rc = launch_sh_bucket_count(items, n, item_bytes, kind, S, sh.d_bases.p, sh.bk.p, s, launches))) return rc;
    std::vector<uint32_t> h(W), start(W);
    OGE_CUDA_TRY(cudaMemcpyAsync(h.data(), sh.bk.p, (size_t) W * 4, cudaMemcpyDeviceToHost, s));
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    uint32_t acc = 0;
    for (int d = 0; d < W; d++) { start[d] = acc; acc += h[d]; counts[d] = h[d]; }
    if (acc != n) return fail_msg(OGE_ERR_STATE, "shard: bucket counts do not add up");
    OGE_CUDA_TRY(cudaMemcpyAsync(sh.bk.p + W, start.data(), (size_t) W * 4, cudaMemcpyHostToDevice, s));
    if ((rc = launch_sh_bucket_scatter(items, n, item_bytes, kind, S, sh.d_bases.p, sh.bk.p + W, sh.bk.p + 2 * W, out.p, s, launches))) return rc;
    OGE_CUDA_TRY(cudaStreamSynchronize(s));      // `start` lives on this stack frame
    return 0;
}

int oge_gpu_shard_key_bytes(oge_gpu_dedup_ctx *c, uint32_t *max_key_bytes) {
    if (!c || !max_key_bytes) return fail_msg(OGE_ERR_INVALID_ARG, "shard_key_bytes: null argument");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    *max_key_bytes = 0;
    if (c->n == 0) return OGE_OK;
    cudaStream_t s = c->stream;
    OGE_CUDA_TRY(cudaEventRecord(c->copy_done, c->copy_stream));
    OGE_CUDA_TRY(cudaStreamWaitEvent(s, c->copy_done, 0));
    uint64_t launches = 0;
    OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p + CNT_SCRATCH0, 0, 4, s));
    int rc = launch_sh_keylen(c->recs(), c->off.p, c->n, nullptr, c->counters.p + CNT_SCRATCH0, s, &launches);
    if (rc) return rc;
    OGE_CUDA_TRY(cudaMemcpyAsync(max_key_bytes, c->counters.p + CNT_SCRATCH0, 4, cudaMemcpyDeviceToHost, s));
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    return OGE_OK;
}

int oge_gpu_shard_set_entry_bytes(oge_gpu_dedup_ctx *c, uint32_t entry_bytes) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "shard_set_entry_bytes: null context");
    if (entry_bytes < 64 || entry_bytes % 32 || entry_bytes > 32 + 544) return fail_msg(OGE_ERR_INVALID_ARG, "shard_set_entry_bytes: %u", entry_bytes);
    c->sh.entry_bytes = entry_bytes;
    return OGE_OK;
}

int oge_gpu_shard_begin(oge_gpu_dedup_ctx *c, void **pub_dev, uint64_t *pub_counts, void **hash_dev, uint64_t *n_hash, void **froute_dev,
                        uint64_t *froute_counts) {
    int rc = need_phase(c, c ? c->sh.phase : 0, "shard_begin");
    if (rc) return rc;
    if (!pub_dev || !pub_counts || !hash_dev || !n_hash || !froute_dev || !froute_counts) return fail_msg(OGE_ERR_INVALID_ARG, "shard_begin: null argument");
    if (c->sh.bases[c->cfg.rank + 1] - c->sh.bases[c->cfg.rank] != c->n)
        return fail_msg(OGE_ERR_INVALID_ARG, "shard_begin: the context holds %llu records, the ranges say %llu", (unsigned long long) c->n,
                        (unsigned long long) (c->sh.bases[c->cfg.rank + 1] - c->sh.bases[c->cfg.rank]));
    if (c->sh.entry_bytes == 0) return fail_msg(OGE_ERR_STATE, "shard_begin: call oge_gpu_shard_set_entry_bytes first (the size all ranks agreed on)");
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    const int W = c->cfg.world;
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.n_records = c->n;
    c->ran = false;
    c->clk_used = 0;
    *pub_dev = *froute_dev = *hash_dev = nullptr;
    *n_hash = 0;
    for (int d = 0; d < W; d++) pub_counts[d] = froute_counts[d] = 0;
    if ((rc = compute_layout(c, &c->kl))) return rc;
    const uint64_t n = c->n;
    uint64_t launches = 0;
    // split keys in the entries' own packing: (ref << coord_bits) | (pos + bias); the record ranges
    {
        std::vector<uint64_t> packed((size_t) std::max(1, W - 1), 0);
        for (int r = 0; r + 1 < W; r++) {
            const int64_t ref = (int32_t) sh.split_keys[2 * r], pos = (int32_t) sh.split_keys[2 * r + 1];
            const int64_t biased = std::min<int64_t>(std::max<int64_t>(pos + c->kl.coord_bias, 0), (1ll << c->kl.coord_bits) - 1);
            packed[r] = ((uint64_t) std::max<int64_t>(ref, 0) << c->kl.coord_bits) | (uint64_t) biased;
            if (ref < 0) packed[r] = ~0ull;      // a shard that starts in the unmapped tail owns no key
        }
        if ((rc = sh.d_split.reserve(packed.size(), false, s)) || (rc = sh.d_bases.reserve(sh.bases.size(), false, s))) return rc;
        OGE_CUDA_TRY(cudaMemcpyAsync(sh.d_split.p, packed.data(), packed.size() * 8, cudaMemcpyHostToDevice, s));
        OGE_CUDA_TRY(cudaMemcpyAsync(sh.d_bases.p, sh.bases.data(), sh.bases.size() * 8, cudaMemcpyHostToDevice, s));
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
        sh.own_lo = c->cfg.rank > 0 ? packed[c->cfg.rank - 1] : 0;
        sh.own_hi = c->cfg.rank + 1 < W ? packed[c->cfg.rank] : ~0ull;
        sh.k1_route_cap = 0;
        if (W > 1 && c->n) {      // K1 lists the boundary fragment ends itself; the sweep below is the fall-back when they do not fit
#ifdef OGE_TESTING
            const bool forced_sweep = getenv("OGE_ROUTE_CAP") != nullptr;
#else
            const bool forced_sweep = false;
#endif
            const uint64_t cap = std::max<uint64_t>(1u << 16, c->n / 64);
            if (!forced_sweep) {
                if ((rc = sh.route.reserve(cap, false, s))) return rc;
                sh.k1_route_cap = (uint32_t) cap;
            }
        }
    }
    OGE_CUDA_TRY(cudaEventRecord(c->copy_done, c->copy_stream));
    OGE_CUDA_TRY(cudaStreamWaitEvent(s, c->copy_done, 0));
    sh.n_frag = sh.n_pe = sh.n_pairs = sh.n_retracted = sh.n_far = sh.n_far_dead = sh.n_slots = sh.n_fm = sh.n_froute_all = sh.n_unpaired = 0;
    sh.n_loc = sh.n_loc_far = 0;
    sh.frag_mode = 0;
    sh.side_pass_used = 0;
    sh.side_pass_bytes = 0;
    OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, CNT_N * 4, s));
    uint64_t n_list = 0;
    if (n) {
        if ((rc = ensure_work(c))) return rc;
        OGE_CUDA_TRY(cudaEventRecord(c->ev[0], s));
        OGE_CUDA_TRY(cudaMemsetAsync(c->dup.p, 0, n, s));
        // K1 + the mate join of this rank's records, as on one GPU but without the replay of the exact path: a name that is
        // not a plain couple here is published instead
        JoinStage js;
        if ((rc = join_stage(c, false, &js, &launches))) return rc;
        OGE_CUDA_TRY(cudaEventRecord(c->ev[2], s));
        sh.n_frag = js.n_frag; sh.n_pe = js.n_pe; sh.n_unpaired = js.n_unpaired;
        sh.n_slots = js.n_slots; sh.n_loc = js.n_loc; sh.n_loc_far = js.n_loc_far;
        sh.n_pairs = js.n_pairs; sh.n_retracted = js.n_retracted; sh.n_far = js.n_far; sh.n_far_dead = js.n_far_retracted;
        c->stats.n_complex_names = js.n_cplx;
        c->stats.n_local_pairs = js.n_loc + js.n_loc_far;
        c->stats.n_join_leftovers = js.n_left;
        // published: names seen once among the leftovers (their slot holds one arrival) and everything on the exact-path list
        if (js.n_slots || js.n_cplx) {
            if ((rc = sh.pub_list.reserve(js.n_left + js.n_cplx + 2 * (uint64_t) (js.n_loc + js.n_loc_far) / 1 + 16, false, s))) return rc;
            if (js.n_slots && (rc = launch_sh_singletons(c->table.p, js.n_slots, sh.pub_list.p, c->counters.p, s, &launches))) return rc;
            if ((rc = launch_sh_complex(c->sortbuf.p, (uint32_t) js.n_cplx, sh.pub_list.p, c->counters.p, s, &launches))) return rc;
            if ((rc = read_counters(c))) return rc;
            n_list = c->h_counters[CNT_PUB];
        }
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
        c->stats.ms_endbuild += ms_between(c->ev[0], c->ev[1]);
        c->stats.ms_join += ms_between(c->ev[1], c->ev[2]);
        c->stats.ms_total += ms_between(c->ev[0], c->ev[2]);
        for (int i = 0; i < c->k_used; i++) c->stats.ms_kernel[c->k_slot[i]] += ms_between(c->k_ev[2 * i], c->k_ev[2 * i + 1]);
    }
    {
        PhaseClock clk(c, &c->stats.ms_join);
        if (n_list) {
            if ((rc = sh.pub_raw.reserve(n_list * sh.entry_bytes, false, s)) || (rc = sh.pub_hash.reserve(n_list, false, s))) return rc;
            if ((rc = launch_sh_gather2(sh.pub_list.p, (uint32_t) n_list, c->frag.p, c->hk.p, c->recs(), c->off.p, sh.pub_raw.p, sh.entry_bytes,
                                        sh.pub_hash.p, c->counters.p + CNT_ERR, s, &launches)))
                return rc;
        }
        if ((rc = bucket_by_destination(c, sh.pub_raw.p, n_list, sh.entry_bytes, 0, sh.pub_send, pub_counts, &launches))) return rc;
        clk.stop();
    }
    // copies of the fragment ends whose key lies in another rank's range leave now: they travel with
    // the first exchange, so that the fragment sort can overlap the pair exchanges (the originals stay:
    // K4 leaves runs alone whose key another rank owns)
    uint64_t n_fr = 0;
    if (n && W > 1) {
        PhaseClock clk(c, &c->stats.ms_select);
        n_fr = sh.k1_route_cap ? c->h_counters[CNT_ROUTE] : 0;
        if (!sh.k1_route_cap || n_fr > sh.k1_route_cap) {      // not listed by K1 (or more of them than it had room for): sweep the end entries
            E128 *lists[1] = {c->frag.p};
            const uint64_t counts[1] = {n};
            const int kinds[1] = {0};
            if ((rc = route_sweep(c, 1, lists, counts, kinds, 2, &n_fr, &launches))) return rc;
        }
        if ((rc = bucket_by_destination(c, sh.route.p, n_fr, sizeof(RouteEntry), 1, sh.froute_send, froute_counts, &launches))) return rc;
        clk.stop();
    }
    if ((rc = read_counters(c))) return rc;
    if (c->h_counters[CNT_ERR] & DEV_ERR_CAPACITY) return fail_msg(OGE_ERR_STATE, "shard_begin: a key is longer than the agreed entry size holds");
    resolve_clocks(c);
    c->stats.launches += launches;
    *pub_dev = sh.pub_send.p;
    *hash_dev = sh.pub_hash.p;
    *n_hash = n_list;
    *froute_dev = sh.froute_send.p;
    sh.phase = 1;
    return OGE_OK;
}

int oge_gpu_shard_probe(oge_gpu_dedup_ctx *c, const void *hash_in_dev, uint64_t n_hash_in, const void *froute_all_dev, uint64_t n_fr_all,
                        void **pub2_dev, uint64_t *pub2_counts, void **proute_dev, uint64_t *proute_counts) {
    int rc = need_phase(c, 1, "shard_probe");
    if (rc) return rc;
    if (!pub2_dev || !pub2_counts || !proute_dev || !proute_counts || (n_hash_in && !hash_in_dev) || (n_fr_all && !froute_all_dev))
        return fail_msg(OGE_ERR_INVALID_ARG, "shard_probe: null argument");
    cudaStream_t s = c->stream, s2 = c->side_stream;
    ShardState &sh = c->sh;
    const int W = c->cfg.world;
    uint64_t launches = 0, n2 = 0, n_pr = 0;
    const uint64_t n = c->n;
    for (int d = 0; d < W; d++) pub2_counts[d] = proute_counts[d] = 0;
    *pub2_dev = *proute_dev = nullptr;

    // ---- local pairs of names some other rank published are retracted and published (this reads the fragment
    //      array, so it comes before the fragment sort starts moving it)
    if (sh.n_pe && n_hash_in) {
        PhaseClock clk(c, &c->stats.ms_join);
        uint64_t set_slots = 1024;
        while (set_slots < 2 * n_hash_in) set_slots <<= 1;
        if ((rc = sh.hset.reserve(set_slots, false, s))) return rc;
        OGE_CUDA_TRY(cudaMemsetAsync(sh.hset.p, 0, set_slots * 8, s));
        if ((rc = launch_sh_set_build((const uint64_t *) hash_in_dev, n_hash_in, sh.hset.p, set_slots, s, &launches))) return rc;
        if ((rc = zero_counter(c, CNT_PUB))) return rc;
        if ((rc = sh.pub_list.reserve(2 * (sh.n_pairs + sh.n_far) + 16, false, s))) return rc;
        if (sh.n_slots && (rc = launch_sh_probe_table(c->table.p, sh.n_slots, sh.hset.p, set_slots, c->pair.p, c->pairf.p, sh.pub_list.p, c->counters.p, s, &launches)))
            return rc;
        const ShardParams S = shard_params(c);
        if ((rc = launch_sh_probe_pairs(c->pair_hk.p, sh.n_loc, sh.hset.p, set_slots, c->pair.p, 0, c->mate_of.p, S, sh.pub_list.p, s, &launches))) return rc;
        if ((rc = launch_sh_probe_pairs(c->pairf_hk.p, sh.n_loc_far, sh.hset.p, set_slots, c->pairf.p, 1, c->mate_of.p, S, sh.pub_list.p, s, &launches))) return rc;
        if ((rc = read_counters(c))) return rc;
        n2 = c->h_counters[CNT_PUB];
        sh.n_retracted = c->h_counters[CNT_PAIRS_RETRACTED];
        sh.n_far_dead = c->h_counters[CNT_FAR_RETRACTED];
        if (n2) {
            if ((rc = sh.pub_raw.reserve(n2 * sh.entry_bytes, false, s))) return rc;
            if ((rc = launch_sh_gather2(sh.pub_list.p, (uint32_t) n2, c->frag.p, c->hk.p, c->recs(), c->off.p, sh.pub_raw.p, sh.entry_bytes, nullptr,
                                        c->counters.p + CNT_ERR, s, &launches)))
                return rc;
        }
        clk.stop();
    }
    if ((rc = bucket_by_destination(c, sh.pub_raw.p, n2, sh.entry_bytes, 0, sh.pub_send, pub2_counts, &launches))) return rc;
    // ---- side stream: the fragment ends are complete once the routed copies are in -> K3 + K4 on them,
    //      concurrently with the pair routing, the second exchange and the replay on the main stream
    sh.n_froute_all = n_fr_all;
    sh.frag_busy = false;
    // how much of the fragment work is needed at all (fragfilter.cu): nothing when neither this shard nor
    // the routed copies hold an unpaired end; the reduced pass when they are rare; else everything
    sh.frag_mode = 2;
    if (sh.n_unpaired == 0 && n_fr_all == 0) sh.frag_mode = 0;
    else if (!c->cfg.debug_full_frag_sort && sh.n_unpaired <= sh.n_frag / 16 && c->kl.f_end - c->kl.f_orient <= 63) sh.frag_mode = 1;
    if (sh.frag_mode && n + n_fr_all) {
        const uint64_t n_all_frag = n + n_fr_all;
        // room for the unpaired ends and the paired ends sharing their keys; with no local unpaired end only the
        // few routed copies can matter, and small capacities keep the (mostly empty) launches of the sort small
        sh.ucap = (sh.n_unpaired ? std::max<uint64_t>(sh.n_frag / 4, 4 * sh.n_unpaired) : 0) + 64 * n_fr_all + 1024;
#ifdef OGE_TESTING
        if (const char *e = getenv("OGE_UFRAG_CAP")) sh.ucap = std::max<uint64_t>(1, (uint64_t) atoll(e));      // test hook: force the fallback
#endif
        if ((rc = c->frag.reserve(n_all_frag, true, s))) return rc;
        if (sh.frag_mode == 2) {
            if ((rc = c->sortbuf.reserve(n_all_frag, false, s))) return rc;
            if ((rc = sh.scratch2.reserve(sort_scratch_bytes(std::max<uint64_t>(n_all_frag, sh.n_pe / 2 + n_hash_in + 1024)), false, s))) return rc;
        } else {
            uint64_t n_slots = 1024;
            while (n_slots < 4 * (sh.n_unpaired + n_fr_all)) n_slots <<= 1;
            sh.uset_slots = n_slots;
            if ((rc = c->ufrag.reserve(sh.ucap, false, s))) return rc;
            if ((rc = c->ufrag2.reserve(sh.ucap, false, s))) return rc;
            if ((rc = c->uset.reserve(n_slots, false, s))) return rc;
            if ((rc = sh.scratch2.reserve(sort_scratch_bytes(std::max<uint64_t>(sh.ucap, sh.n_pe / 2 + n_hash_in + 1024)), false, s))) return rc;
        }
        if ((rc = sh.marks_frag.reserve(n_fr_all + 16, false, s))) return rc;
        if (n == 0 && (rc = c->mate_of.reserve(1, false, s))) return rc;
        if (n == 0 && (rc = c->dup.reserve(1, false, s))) return rc;
        OGE_CUDA_TRY(cudaEventRecord(sh.ev_main, s));
        OGE_CUDA_TRY(cudaStreamWaitEvent(s2, sh.ev_main, 0));
        OGE_CUDA_TRY(cudaEventRecord(sh.ev_side[0], s2));
        if (n_fr_all) {
            OGE_CUDA_TRY(cudaMemsetAsync(c->frag.p + n, 0xFF, n_fr_all * sizeof(E128), s2));      // unused tail slots are dead entries
            if ((rc = launch_sh_receive((const RouteEntry *) froute_all_dev, n_fr_all, shard_params(c), 1u, c->frag.p + n, (uint32_t) n_fr_all,
                                        nullptr, 0, nullptr, 0, nullptr, nullptr, 0, s2, &launches)))
                return rc;
        }
        PassTimer timer2{c->pass_ev + 48, 24, 0, 0};
        E128 *sorted_frags = c->frag.p;
        SelectParams sp;
        sp.dup = c->dup.p; sp.mate_of = c->mate_of.p; sp.idx_base = c->cfg.index_base; sp.n_records = n;
        sp.counters = c->counters.p; sp.kl = c->kl;
        sp.fm = nullptr; sp.n_fm = 0; sp.foreign_marks = sh.marks_frag.p; sp.foreign_cap = (uint32_t) sh.marks_frag.cap;
        sp.foreign_counter = c->counters.p + CNT_FOREIGN_MARKS_FRAG;
        sp.split = sh.d_split.p; sp.world = c->cfg.world; sp.rank = c->cfg.rank;
        if (sh.frag_mode == 1) {
            // sizes stay on the device (the side stream never waits for the host): CNT_UFRAG counts what was
            // collected, CNT_FRAG_VALID is that count if it fits the list and 0 otherwise (overflow: finish()
            // notices and falls back to the full sort)
            OGE_CUDA_TRY(cudaMemsetAsync(c->uset.p, 0, sh.uset_slots * 8, s2));
            OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p + CNT_UFRAG, 0, 4, s2));
            // unpaired ends: the local ones (none on clean paired-end data: the host knows from K1) and the routed copies
            if (sh.n_unpaired && (rc = launch_ff_collect(c->frag.p, n, c->kl, c->ufrag.p, (uint32_t) sh.ucap, c->counters.p, s2, &launches))) return rc;
            if ((rc = launch_ff_collect(c->frag.p + n, n_fr_all, c->kl, c->ufrag.p, (uint32_t) sh.ucap, c->counters.p, s2, &launches))) return rc;
            OGE_CUDA_TRY(cudaMemcpyAsync(c->counters.p + CNT_SCRATCH2, c->counters.p + CNT_UFRAG, 4, cudaMemcpyDeviceToDevice, s2));      // size of the set
            if ((rc = launch_ff_set_build(c->ufrag.p, c->counters.p + CNT_UFRAG, (uint32_t) std::min<uint64_t>(sh.n_unpaired + n_fr_all, sh.ucap),
                                          c->kl, c->uset.p, sh.uset_slots, s2, &launches)))
                return rc;
            if ((rc = launch_ff_filter(c->frag.p, n_all_frag, c->kl, c->uset.p, sh.uset_slots, c->ufrag.p, (uint32_t) sh.ucap, c->counters.p,
                                       c->counters.p + CNT_SCRATCH2, s2, &launches)))
                return rc;
            sh_fit_counter_kernel<<<1, 1, 0, s2>>>(c->counters.p + CNT_FRAG_VALID, c->counters.p + CNT_UFRAG, (uint32_t) sh.ucap);
            if ((rc = radix_sort_128(c->ufrag.p, c->ufrag2.p, sh.ucap, c->counters.p + CNT_FRAG_VALID, c->kl.f_orient, c->kl.f_end, sh.scratch2.p,
                                     s2, &sorted_frags, &launches, c->cfg.profile_events ? &timer2 : nullptr)))
                return rc;
            sp.n_max = (uint32_t) sh.ucap;
        } else {
            sh_add_counter_kernel<<<1, 1, 0, s2>>>(c->counters.p + CNT_FRAG_VALID, (uint32_t) sh.n_frag, c->counters.p + CNT_FRAG_EXTRA, (uint32_t) n_fr_all);
            if ((rc = radix_sort_128(c->frag.p, c->sortbuf.p, n_all_frag, nullptr, c->kl.f_orient, c->kl.f_end, sh.scratch2.p, s2, &sorted_frags,
                                     &launches, c->cfg.profile_events ? &timer2 : nullptr)))
                return rc;
            sp.n_max = (uint32_t) n_all_frag;
        }
        sh.side_pass_used = timer2.used;
        sh.side_pass_bytes = timer2.bytes;
        OGE_CUDA_TRY(cudaEventRecord(sh.ev_side[1], s2));
        sp.sorted = sorted_frags; sp.n_dev = c->counters.p + CNT_FRAG_VALID;
        if ((rc = launch_select_frags(sp, s2, &launches))) return rc;
        OGE_CUDA_TRY(cudaEventRecord(sh.ev_side[2], s2));
        sh.frag_busy = true;
    }

    // ---- the remaining local couples whose key lies in another rank's range leave
    if (c->cfg.world > 1 && (sh.n_pairs || sh.n_far)) {
        PhaseClock clk(c, &c->stats.ms_select);
        E128 *lists[2] = {c->pair.p, c->pairf.p};
        const uint64_t counts[2] = {sh.n_pairs, sh.n_far};
        const int kinds[2] = {1, 2};
        if ((rc = route_sweep(c, 2, lists, counts, kinds, 0, &n_pr, &launches))) return rc;
        sh.n_retracted += n_pr ? c->h_counters[CNT_SCRATCH0] : 0;      // dead pair entries, whatever the reason
        sh.n_far_dead += n_pr ? c->h_counters[CNT_SCRATCH1] : 0;
        clk.stop();
    }
    if ((rc = bucket_by_destination(c, sh.route.p, n_pr, sizeof(RouteEntry), 1, sh.proute_send, proute_counts, &launches))) return rc;
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    resolve_clocks(c);
    c->stats.launches += launches;
    *pub2_dev = sh.pub_send.p;
    *proute_dev = sh.proute_send.p;
    sh.phase = 2;
    return OGE_OK;
}

int oge_gpu_shard_replay(oge_gpu_dedup_ctx *c, const void *w_dev, uint64_t n_w, void **oroute_dev, uint64_t *oroute_counts) {
    int rc = need_phase(c, 2, "shard_replay");
    if (rc) return rc;
    if (!oroute_dev || !oroute_counts || (n_w && !w_dev)) return fail_msg(OGE_ERR_INVALID_ARG, "shard_replay: null argument");
    if (n_w >= (1ull << 30)) return fail_msg(OGE_ERR_TOO_LARGE, "shard_replay: published set too large");
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    const int W = c->cfg.world;
    uint64_t launches = 0, n_or = 0;
    const uint64_t n = c->n;
    sh.n_w = n_w;
    *oroute_dev = nullptr;
    for (int d = 0; d < W; d++) oroute_counts[d] = 0;
    // ---- the names this rank owns, replayed over all ranks' sightings: the pairs whose key range it owns too join its
    //      lists, the others leave for their owners
    if (n_w) {
        PhaseClock clk(c, &c->stats.ms_join);
        const uint64_t pair_cap = sh.n_pairs + n_w / 2 + 16, far_cap = sh.n_far + n_w / 2 + 16;
        if ((rc = c->pair.reserve(pair_cap, true, s))) return rc;
        if ((rc = c->pair2.reserve(pair_cap, false, s))) return rc;
        if ((rc = c->pairf.reserve(far_cap, true, s))) return rc;
        if ((rc = c->pairf2.reserve(far_cap, false, s))) return rc;
        if ((rc = sh.fm.reserve(n_w / 2 + 16, false, s))) return rc;
        if ((rc = sh.route.reserve(n_w / 2 + 16, false, s))) return rc;
        if ((rc = c->scratch.reserve(std::max(sort_scratch_bytes(std::max(std::max(n_w, pair_cap), far_cap)), c->scratch.cap), true, s))) return rc;
        if (n == 0 && (rc = c->mate_of.reserve(1, false, s))) return rc;
        if ((rc = sh.w_sort.reserve(n_w, false, s))) return rc;
        if ((rc = sh.w_sort2.reserve(n_w, false, s))) return rc;
        if ((rc = c->cplx_state.reserve(n_w, false, s))) return rc;
        if ((rc = zero_counter(c, CNT_ROUTE))) return rc;
        if ((rc = launch_sh_wbuild2((const uint8_t *) w_dev, sh.entry_bytes, (uint32_t) n_w, c->kl, sh.w_sort.p, s, &launches))) return rc;
        E128 *sorted = nullptr;
        if ((rc = radix_sort_128(sh.w_sort.p, sh.w_sort2.p, n_w, nullptr, 32, 96, c->scratch.p, s, &sorted, &launches))) return rc;
        if ((rc = launch_sh_replay2(sorted, (uint32_t) n_w, (const uint8_t *) w_dev, sh.entry_bytes, c->cplx_state.p, shard_params(c), c->pair.p,
                                    (uint32_t) c->pair.cap, c->pairf.p, (uint32_t) c->pairf.cap, c->mate_of.p, sh.fm.p, (uint32_t) sh.fm.cap,
                                    sh.route.p, (uint32_t) sh.route.cap, s, &launches)))
            return rc;
        if ((rc = read_counters(c))) return rc;
        n_or = c->h_counters[CNT_ROUTE];
        if (n_or > sh.route.cap) return fail_msg(OGE_ERR_STATE, "shard_replay: route list overran its buffer");
        clk.stop();
    }
    if ((rc = bucket_by_destination(c, sh.route.p, n_or, sizeof(RouteEntry), 1, sh.oroute_send, oroute_counts, &launches))) return rc;
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    resolve_clocks(c);
    c->stats.launches += launches;
    *oroute_dev = sh.oroute_send.p;
    sh.phase = 3;
    return OGE_OK;
}

int oge_gpu_shard_finish(oge_gpu_dedup_ctx *c, const void *proute_all_dev, uint64_t n_all, void **marks_dev, uint64_t *marks_counts) {
    int rc = need_phase(c, 3, "shard_finish");
    if (rc) return rc;
    if (!marks_dev || !marks_counts || (n_all && !proute_all_dev)) return fail_msg(OGE_ERR_INVALID_ARG, "shard_finish: null argument");
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    const int W = c->cfg.world;
    uint64_t launches = 0;
    const uint64_t n = c->n, n_w = sh.n_w;
    PassTimer timer{c->pass_ev, 48, 0, 0};
    PassTimer *tp = c->cfg.profile_events ? &timer : nullptr;
    *marks_dev = nullptr;
    for (int d = 0; d < W; d++) marks_counts[d] = 0;

    // ---- pair ends other ranks handed over (their local pairs, and pairs their replay formed): the ones whose key this rank owns
    if (n_all) {
        PhaseClock clk(c, &c->stats.ms_select);
        const uint64_t pair_cap = sh.n_pairs + n_w / 2 + n_all + 16, far_cap = sh.n_far + n_w / 2 + n_all + 16;
        if ((rc = c->pair.reserve(pair_cap, true, s))) return rc;
        if ((rc = c->pair2.reserve(pair_cap, false, s))) return rc;
        if ((rc = c->pairf.reserve(far_cap, true, s))) return rc;
        if ((rc = c->pairf2.reserve(far_cap, false, s))) return rc;
        if ((rc = sh.fm.reserve(n_w / 2 + n_all + 16, true, s))) return rc;
        if ((rc = c->scratch.reserve(std::max(sort_scratch_bytes(std::max(pair_cap, far_cap)), c->scratch.cap), true, s))) return rc;
        if (n == 0 && (rc = c->mate_of.reserve(1, false, s))) return rc;
        if ((rc = launch_sh_receive((const RouteEntry *) proute_all_dev, n_all, shard_params(c), 6u, nullptr, 0, c->pair.p, (uint32_t) c->pair.cap,
                                    c->pairf.p, (uint32_t) c->pairf.cap, c->mate_of.p, sh.fm.p, (uint32_t) sh.fm.cap, s, &launches)))
            return rc;
        clk.stop();
    }
    if ((rc = read_counters(c))) return rc;
    sh.n_pairs = c->h_counters[CNT_PAIRS];
    sh.n_far = c->h_counters[CNT_PAIRS_FAR];
    sh.n_fm = c->h_counters[CNT_FM];
    if (sh.n_pairs > c->pair.cap || sh.n_far > c->pairf.cap || sh.n_fm > sh.fm.cap)
        return fail_msg(OGE_ERR_STATE, "shard_finish: pair lists overran their buffers");
    if (sh.n_fm > 1) {      // foreign mates sorted by idx1 for the binary search in K4
        PhaseClock clk(c, &c->stats.ms_select);
        if ((rc = sh.fm_sort.reserve(2 * sh.n_fm, false, s))) return rc;
        if ((rc = launch_sh_fm_pack(sh.fm.p, (uint32_t) sh.n_fm, sh.fm_sort.p, s, &launches))) return rc;
        E128 *sorted = nullptr;
        if ((rc = radix_sort_128(sh.fm_sort.p, sh.fm_sort.p + sh.n_fm, sh.n_fm, nullptr, 32, 64, c->scratch.p, s, &sorted, &launches))) return rc;
        if ((rc = launch_sh_fm_unpack(sorted, (uint32_t) sh.n_fm, sh.fm.p, s, &launches))) return rc;
        clk.stop();
    }
    // marks on other ranks' records can only come from entries that crossed ranks: pairs formed by the
    // replay (at most one per two published entries) and routed entries
    if ((rc = sh.marks.reserve(n_w + 2 * n_all + 1024, false, s))) return rc;

    SelectParams sp;
    sp.dup = c->dup.p; sp.mate_of = c->mate_of.p; sp.idx_base = c->cfg.index_base; sp.n_records = n;
    sp.counters = c->counters.p; sp.kl = c->kl; sp.n_dev = nullptr;
    sp.fm = sh.fm.p; sp.n_fm = (uint32_t) sh.n_fm; sp.foreign_marks = sh.marks.p; sp.foreign_cap = (uint32_t) sh.marks.cap;
    sp.foreign_counter = c->counters.p + CNT_FOREIGN_MARKS;
    sp.split = sh.d_split.p; sp.world = c->cfg.world; sp.rank = c->cfg.rank;
    // far pairs (cross-contig / huge inserts: a short list, 9 launch-bound passes) go to the side stream, behind
    // whatever fragment work is still queued there, and run concurrently with the near-pair sort
    bool far_on_side = false;
    if (sh.n_far) {
        const size_t need = sort_scratch_bytes(sh.n_far);
        if (sh.scratch2.cap >= need || !sh.frag_busy) {
            if (sh.scratch2.cap < need && (rc = sh.scratch2.reserve(need, false, s))) return rc;
            far_on_side = true;
        }
    }
    for (int far = 1; far >= 0; far--) {      // far pairs first (queued on the side stream), then near pairs (short key)
        const uint64_t cnt = far ? sh.n_far : sh.n_pairs, dead = far ? sh.n_far_dead : sh.n_retracted;
        if (!cnt) continue;
        const bool side = far && far_on_side;
        cudaStream_t st = side ? c->side_stream : s;
        E128 *a = far ? c->pairf.p : c->pair.p, *b = far ? c->pairf2.p : c->pair2.p, *sorted = a;
        if (side) {
            OGE_CUDA_TRY(cudaEventRecord(sh.ev_main, s));      // the lists, the mate table and the foreign mates are final
            OGE_CUDA_TRY(cudaStreamWaitEvent(st, sh.ev_main, 0));
            PassTimer timer3{c->pass_ev + 48 + 2 * sh.side_pass_used, 24 - sh.side_pass_used, 0, 0};
            if ((rc = radix_sort_128(a, b, cnt, nullptr, c->kl.p_coord2, c->kl.p_end, sh.scratch2.p, st, &sorted, &launches,
                                     c->cfg.profile_events ? &timer3 : nullptr)))
                return rc;
            sh.side_pass_used += timer3.used;
            sh.side_pass_bytes += timer3.bytes;
            if (cnt > dead) {
                sp.sorted = sorted; sp.n_max = (uint32_t) (cnt - dead);
                if ((rc = launch_select_pairs(sp, true, st, &launches))) return rc;
            }
            OGE_CUDA_TRY(cudaEventRecord(sh.ev_far, st));
            continue;
        }
        {
            PhaseClock clk(c, &c->stats.ms_sort_pair);
            if ((rc = radix_sort_128(a, b, cnt, nullptr, far ? c->kl.p_coord2 : c->kl.n_delta, c->kl.p_end, c->scratch.p, s, &sorted, &launches, tp)))
                return rc;
            clk.stop();
        }
        if (cnt > dead) {
            PhaseClock clk(c, &c->stats.ms_select);
            sp.sorted = sorted; sp.n_max = (uint32_t) (cnt - dead);
            if ((rc = launch_select_pairs(sp, far != 0, s, &launches))) return rc;
            clk.stop();
        }
    }
    if (far_on_side) {
        PhaseClock clk(c, &c->stats.ms_sort_pair);      // whatever of the far-pair work the near-pair work did not hide
        OGE_CUDA_TRY(cudaStreamWaitEvent(s, sh.ev_far, 0));
        clk.stop();
    }
    // ---- join the side stream: the fragment verdicts
    uint64_t extra = 0;
    sh.frag_ran = sh.frag_busy;
    if (sh.frag_busy) {
        PhaseClock clk(c, nullptr);      // whatever of the fragment work the pair work did not hide
        OGE_CUDA_TRY(cudaStreamWaitEvent(s, sh.ev_side[2], 0));
        clk.stop();
        sh.frag_busy = false;
    }
    if ((rc = read_counters(c))) return rc;
    if (sh.frag_mode == 1 && c->h_counters[CNT_UFRAG] > sh.ucap) {
        // more paired ends share a key with an unpaired one than the list holds: nothing was selected
        // (the device-side size was forced to 0); sort every fragment end now, on this stream
        PhaseClock clk(c, &c->stats.ms_sort_frag);
        const uint64_t n_all_frag = n + sh.n_froute_all;
        if ((rc = c->sortbuf.reserve(n_all_frag, false, s))) return rc;
        if ((rc = c->scratch.reserve(std::max(sort_scratch_bytes(n_all_frag), c->scratch.cap), true, s))) return rc;
        sh_add_counter_kernel<<<1, 1, 0, s>>>(c->counters.p + CNT_FRAG_VALID, (uint32_t) sh.n_frag, c->counters.p + CNT_FRAG_EXTRA, (uint32_t) sh.n_froute_all);
        E128 *sorted_frags = c->frag.p;
        if ((rc = radix_sort_128(c->frag.p, c->sortbuf.p, n_all_frag, nullptr, c->kl.f_orient, c->kl.f_end, c->scratch.p, s, &sorted_frags, &launches, tp)))
            return rc;
        sp.fm = nullptr; sp.n_fm = 0; sp.foreign_marks = sh.marks_frag.p; sp.foreign_cap = (uint32_t) sh.marks_frag.cap;
        sp.foreign_counter = c->counters.p + CNT_FOREIGN_MARKS_FRAG;
        sp.sorted = sorted_frags; sp.n_max = (uint32_t) n_all_frag; sp.n_dev = c->counters.p + CNT_FRAG_VALID;
        if ((rc = launch_select_frags(sp, s, &launches))) return rc;
        clk.stop();
        if ((rc = read_counters(c))) return rc;
        sh.frag_mode = 2;
    }
    if (sh.n_froute_all || n) {
        extra = std::min<uint64_t>(c->h_counters[CNT_FRAG_EXTRA], sh.n_froute_all);
        if (sh.frag_ran) {
            c->stats.ms_sort_frag += ms_between(sh.ev_side[0], sh.ev_side[1]);
            c->stats.ms_select += ms_between(sh.ev_side[1], sh.ev_side[2]);
        }
    }
    resolve_clocks(c);
    const uint64_t n_foreign = c->h_counters[CNT_FOREIGN_MARKS], n_foreign_frag = c->h_counters[CNT_FOREIGN_MARKS_FRAG];
    if (n_foreign > sh.marks.cap || n_foreign_frag > sh.marks_frag.cap)
        return fail_msg(OGE_ERR_STATE, "shard_finish: %llu + %llu marks for other ranks, room for %llu + %llu", (unsigned long long) n_foreign,
                        (unsigned long long) n_foreign_frag, (unsigned long long) sh.marks.cap, (unsigned long long) sh.marks_frag.cap);
    c->stats.launches += launches;
    c->stats.n_frag_entries = sh.n_frag + extra;
    c->stats.n_pair_entries = sh.n_pairs - sh.n_retracted + sh.n_far - sh.n_far_dead;
    for (int i = 0; i < timer.used; i++) c->stats.ms_sort_pass_kernels += ms_between(c->pass_ev[2 * i], c->pass_ev[2 * i + 1]);
    for (int i = 0; i < sh.side_pass_used; i++) c->stats.ms_sort_pass_kernels += ms_between(c->pass_ev[48 + 2 * i], c->pass_ev[48 + 2 * i + 1]);
    c->stats.sort_pass_launches = timer.used + sh.side_pass_used;
    c->stats.sort_pass_bytes = timer.bytes + sh.side_pass_bytes;
    // both mark lists (from pairs, from fragments) leave as one, ordered by the rank that holds the record
    if (n_foreign_frag) {
        if ((rc = sh.marks.reserve(n_foreign + n_foreign_frag, true, s))) return rc;
        OGE_CUDA_TRY(cudaMemcpyAsync(sh.marks.p + n_foreign, sh.marks_frag.p, n_foreign_frag * 4, cudaMemcpyDeviceToDevice, s));
    }
    if ((rc = bucket_by_destination(c, sh.marks.p, n_foreign + n_foreign_frag, 4, 2, sh.marks_send, marks_counts, &launches))) return rc;
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    *marks_dev = sh.marks_send.p;
    sh.phase = 5;
    return OGE_OK;
}

int oge_gpu_shard_apply(oge_gpu_dedup_ctx *c, const void *marks_all_dev, uint64_t n_all) {
    int rc = need_phase(c, 5, "shard_apply");
    if (rc) return rc;
    if (n_all && !marks_all_dev) return fail_msg(OGE_ERR_INVALID_ARG, "shard_apply: null argument");
    cudaStream_t s = c->stream;
    uint64_t launches = 0;
    if (c->n) {
        PhaseClock clk(c, &c->stats.ms_flags);
        if ((rc = launch_sh_apply_marks((const uint32_t *) marks_all_dev, n_all, c->cfg.index_base, c->n, c->dup.p, s, &launches))) return rc;
        FlagParams fp;
        fp.rec = c->recs(); fp.off = c->off.p; fp.n = c->n; fp.flag_in = c->flag_in.p; fp.flag_out = c->flag_out.p;
        fp.dup = c->dup.p; fp.counters = c->counters.p; fp.quiet_index_bug = 0;
        if ((rc = launch_flags(fp, s, &launches))) return rc;
        clk.stop();
        if ((rc = read_counters(c))) return rc;
        resolve_clocks(c);
        c->stats.n_duplicates = c->h_counters[CNT_DUPS];
    }
    c->stats.launches += launches;
    c->stats.n_hash_mismatch = c->h_counters[CNT_HASH_MISMATCH];
    c->stats.frag_key_bits = c->kl.f_end - c->kl.f_orient;
    c->stats.pair_key_bits = c->kl.p_end - c->kl.n_delta;
    c->stats.frag_sort_passes = make_sort_plan(c->kl.f_orient, c->kl.f_end).n_pass;
    c->stats.pair_sort_passes = make_sort_plan(c->kl.n_delta, c->kl.p_end).n_pass;
    c->ran = true;
    c->sh.phase = 0;
    return OGE_OK;
}

}  // extern "C"
