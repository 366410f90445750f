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
int launch_sh_receive(const RouteEntry *in, uint64_t n_in, const ShardParams &S, E128 *frag_extra, uint32_t frag_cap, E128 *pair,
                      uint32_t pair_cap, E128 *pair_far, uint32_t far_cap, uint32_t *mate_of, uint64_t *fm, uint32_t fm_cap, cudaStream_t s,
                      uint64_t *launches);
int launch_sh_fm_pack(const uint64_t *fm, uint32_t n, E128 *out, cudaStream_t s, uint64_t *launches);
int launch_sh_fm_unpack(const E128 *in, uint32_t n, uint64_t *fm, cudaStream_t s, uint64_t *launches);
int launch_sh_apply_marks(const uint32_t *marks, uint64_t n_marks, uint64_t idx_base, uint64_t n, uint8_t *dup, cudaStream_t s,
                          uint64_t *launches);
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

// every phase is bracketed by events; the device time of the phases adds up to stats.ms_total
struct PhaseClock {
    oge_gpu_dedup_ctx *c;
    float *slot;
    PhaseClock(oge_gpu_dedup_ctx *ctx, float *stage) : c(ctx), slot(stage) { cudaEventRecord(c->ev[8], c->stream); }
    void stop() {
        cudaEventRecord(c->ev[9], c->stream);
        cudaEventSynchronize(c->ev[9]);
        float ms = ms_between(c->ev[8], c->ev[9]);
        c->stats.ms_total += ms;
        if (slot) *slot += ms;
    }
};

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

int oge_gpu_shard_begin(oge_gpu_dedup_ctx *c, void **pub_dev, uint64_t *n_pub) {
    int rc = need_phase(c, c ? c->sh.phase : 0, "shard_begin");
    if (rc) return rc;
    if (!pub_dev || !n_pub) return fail_msg(OGE_ERR_INVALID_ARG, "shard_begin: null argument");
    if (c->sh.bases[c->cfg.rank + 1] - c->sh.bases[c->cfg.rank] != c->n)
        return fail_msg(OGE_ERR_INVALID_ARG, "shard_begin: the context holds %llu records, the ranges say %llu", (unsigned long long) c->n,
                        (unsigned long long) (c->sh.bases[c->cfg.rank + 1] - c->sh.bases[c->cfg.rank]));
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.n_records = c->n;
    c->ran = false;
    *pub_dev = nullptr;
    *n_pub = 0;
    if ((rc = compute_layout(c, &c->kl))) return rc;
    const uint64_t n = c->n;
    uint64_t launches = 0;
    // split keys in the entries' own packing: (ref << coord_bits) | (pos + bias)
    {
        std::vector<uint64_t> packed((size_t) std::max(1, c->cfg.world - 1), 0);
        for (int r = 0; r + 1 < c->cfg.world; r++) {
            const int64_t ref = (int32_t) sh.split_keys[2 * r], pos = (int32_t) sh.split_keys[2 * r + 1];
            const int64_t biased = std::min<int64_t>(std::max<int64_t>(pos + c->kl.coord_bias, 0), (1ll << c->kl.coord_bits) - 1);
            packed[r] = ((uint64_t) std::max<int64_t>(ref, 0) << c->kl.coord_bits) | (uint64_t) biased;
            if (ref < 0) packed[r] = ~0ull;      // a shard that starts in the unmapped tail owns no key
        }
        if ((rc = sh.d_split.reserve(packed.size(), false, s))) return rc;
        OGE_CUDA_TRY(cudaMemcpyAsync(sh.d_split.p, packed.data(), packed.size() * 8, cudaMemcpyHostToDevice, s));
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
    }
    OGE_CUDA_TRY(cudaEventRecord(c->copy_done, c->copy_stream));
    OGE_CUDA_TRY(cudaStreamWaitEvent(s, c->copy_done, 0));
    sh.n_frag = sh.n_pe = sh.n_pairs = sh.n_retracted = sh.n_far = sh.n_far_dead = sh.n_slots = sh.n_fm = 0;
    if (n) {
        if ((rc = ensure_work(c))) return rc;
        PhaseClock clk(c, &c->stats.ms_endbuild);
        OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, CNT_N * 4, s));
        OGE_CUDA_TRY(cudaMemsetAsync(c->dup.p, 0, n, s));
        EndbuildParams eb;
        eb.rec = c->rec.p; eb.off = c->off.p; eb.n = n; eb.idx_base = c->cfg.index_base;
        eb.frag = c->frag.p; eb.hk = c->hk.p; eb.tag = c->tag.p; eb.flag_in = c->flag_in.p;
        eb.counters = c->counters.p; eb.rg = rg_table(c); eb.kl = c->kl;
        if ((rc = launch_endbuild(eb, (uint32_t) (c->rec_bytes / n), c->sms, s, &launches))) return rc;
        clk.stop();
        if ((rc = read_counters(c))) return rc;
        if ((rc = check_endbuild_errors(c))) return rc;
        sh.n_frag = c->h_counters[CNT_FRAG];
        sh.n_pe = c->h_counters[CNT_PAIR_ELIGIBLE];
    } else {
        OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, CNT_N * 4, s));
    }
    uint64_t n_list = 0;
    if (sh.n_pe) {
        PhaseClock clk(c, &c->stats.ms_join);
        const uint64_t n_pe = sh.n_pe;
        sh.n_slots = n_pe + 1024;
        if ((rc = c->table.reserve(sh.n_slots, false, s))) return rc;
        if ((rc = c->pair.reserve(n_pe / 2 + 1024, false, s))) return rc;
        if ((rc = c->pair2.reserve(n_pe / 2 + 1024, false, s))) return rc;
        if ((rc = c->pairf.reserve(n_pe / 2 + 1024, false, s))) return rc;
        if ((rc = c->pairf2.reserve(n_pe / 2 + 1024, false, s))) return rc;
        if ((rc = c->cplx_slots.reserve(n_pe / 3 + 1024, false, s))) return rc;
        if ((rc = sh.pub_list.reserve(n_pe + 16, false, s))) return rc;
        OGE_CUDA_TRY(cudaMemsetAsync(c->table.p, 0, sh.n_slots * sizeof(MateSlot), s));
        JoinParams jp;
        jp.rec = c->rec.p; jp.off = c->off.p; jp.n = n; jp.idx_base = c->cfg.index_base;
        jp.frag = c->frag.p; jp.hk = c->hk.p; jp.tag = c->tag.p;
        jp.table = c->table.p; jp.n_slots = sh.n_slots;
        jp.pair = c->pair.p; jp.pair_far = c->pairf.p; jp.mate_of = c->mate_of.p; jp.cplx = c->sortbuf.p; jp.cplx_slots = c->cplx_slots.p;
        jp.counters = c->counters.p; jp.rg = rg_table(c); jp.kl = c->kl; jp.verify_names = c->cfg.verify_names;
        if ((rc = launch_mate_join(jp, s, &launches))) return rc;
        if ((rc = read_counters(c))) return rc;
        if (c->h_counters[CNT_COMPLEX_SLOTS]) {
            if ((rc = launch_mate_fixup(jp, c->h_counters[CNT_COMPLEX_SLOTS], s, &launches))) return rc;
            if ((rc = read_counters(c))) return rc;
        }
        // published: names seen once (their slot holds one arrival) and everything on the exact-path list
        if ((rc = launch_sh_singletons(c->table.p, sh.n_slots, sh.pub_list.p, c->counters.p, s, &launches))) return rc;
        if ((rc = launch_sh_complex(c->sortbuf.p, c->h_counters[CNT_COMPLEX], sh.pub_list.p, c->counters.p, s, &launches))) return rc;
        if ((rc = read_counters(c))) return rc;
        n_list = c->h_counters[CNT_PUB];
        sh.n_pairs = c->h_counters[CNT_PAIRS];
        sh.n_retracted = c->h_counters[CNT_PAIRS_RETRACTED];
        sh.n_far = c->h_counters[CNT_PAIRS_FAR];
        sh.n_far_dead = c->h_counters[CNT_FAR_RETRACTED];
        c->stats.n_complex_names = c->h_counters[CNT_COMPLEX];
        if ((rc = sh.pub.reserve(n_list + 1, false, s))) return rc;
        if ((rc = launch_sh_gather(sh.pub_list.p, (uint32_t) n_list, c->frag.p, c->hk.p, c->tag.p, sh.pub.p, s, &launches))) return rc;
        clk.stop();
    }
    c->stats.launches += launches;
    *pub_dev = sh.pub.p;
    *n_pub = n_list;
    sh.phase = 1;
    return OGE_OK;
}

int oge_gpu_shard_probe(oge_gpu_dedup_ctx *c, const void *pub_all_dev, uint64_t n_all, void **pub2_dev, uint64_t *n_pub2) {
    int rc = need_phase(c, 1, "shard_probe");
    if (rc) return rc;
    if (!pub2_dev || !n_pub2 || (n_all && !pub_all_dev)) return fail_msg(OGE_ERR_INVALID_ARG, "shard_probe: null argument");
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    uint64_t launches = 0, n2 = 0;
    if (sh.n_pe && n_all) {
        PhaseClock clk(c, &c->stats.ms_join);
        if ((rc = zero_counter(c, CNT_PUB))) return rc;
        if ((rc = launch_sh_probe((const PubEntry *) pub_all_dev, n_all, shard_params(c), c->table.p, sh.n_slots, c->pair.p, c->pairf.p, sh.pub_list.p, s,
                                  &launches)))
            return rc;
        if ((rc = read_counters(c))) return rc;
        n2 = c->h_counters[CNT_PUB];
        sh.n_retracted = c->h_counters[CNT_PAIRS_RETRACTED];
        sh.n_far_dead = c->h_counters[CNT_FAR_RETRACTED];
        if ((rc = sh.pub2.reserve(n2 + 1, false, s))) return rc;
        if ((rc = launch_sh_gather(sh.pub_list.p, (uint32_t) n2, c->frag.p, c->hk.p, c->tag.p, sh.pub2.p, s, &launches))) return rc;
        clk.stop();
    }
    c->stats.launches += launches;
    *pub2_dev = sh.pub2.p;
    *n_pub2 = n2;
    sh.phase = 2;
    return OGE_OK;
}

int oge_gpu_shard_replay(oge_gpu_dedup_ctx *c, const void *w_dev, uint64_t n_w) {
    int rc = need_phase(c, 2, "shard_replay");
    if (rc) return rc;
    if (n_w && !w_dev) return fail_msg(OGE_ERR_INVALID_ARG, "shard_replay: null argument");
    if (n_w >= (1ull << 30)) return fail_msg(OGE_ERR_TOO_LARGE, "shard_replay: published set too large");
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    uint64_t launches = 0;
    sh.n_w = n_w;
    if (n_w) {
        PhaseClock clk(c, &c->stats.ms_join);
        const uint64_t pair_cap = sh.n_pairs + n_w / 2 + 16, far_cap = sh.n_far + n_w / 2 + 16;
        if ((rc = c->pair.reserve(pair_cap, true, s))) return rc;
        if ((rc = c->pair2.reserve(pair_cap, false, s))) return rc;
        if ((rc = c->pairf.reserve(far_cap, true, s))) return rc;
        if ((rc = c->pairf2.reserve(far_cap, false, s))) return rc;
        if ((rc = sh.fm.reserve(n_w / 2 + 16, false, s))) return rc;
        if ((rc = sh.w_sort.reserve(n_w, false, s))) return rc;
        if ((rc = sh.w_sort2.reserve(n_w, false, s))) return rc;
        if ((rc = c->cplx_state.reserve(n_w, false, s))) return rc;
        if ((rc = c->scratch.reserve(std::max(sort_scratch_bytes(n_w), c->scratch.cap), true, s))) return rc;
        if (c->n == 0 && (rc = c->mate_of.reserve(1, false, s))) return rc;
        if ((rc = launch_sh_wbuild((const PubEntry *) w_dev, (uint32_t) n_w, c->kl, sh.w_sort.p, s, &launches))) return rc;
        E128 *sorted = nullptr;
        if ((rc = radix_sort_128(sh.w_sort.p, sh.w_sort2.p, n_w, nullptr, 32, 96, c->scratch.p, s, &sorted, &launches))) return rc;
        if ((rc = launch_sh_replay(sorted, (uint32_t) n_w, (const PubEntry *) w_dev, c->cplx_state.p, shard_params(c), c->pair.p,
                                   (uint32_t) c->pair.cap, c->pairf.p, (uint32_t) c->pairf.cap, c->mate_of.p, sh.fm.p, (uint32_t) sh.fm.cap, rg_table(c), s, &launches)))
            return rc;
        if ((rc = read_counters(c))) return rc;
        sh.n_pairs = c->h_counters[CNT_PAIRS];
        sh.n_far = c->h_counters[CNT_PAIRS_FAR];
        sh.n_fm = c->h_counters[CNT_FM];
        clk.stop();
    }
    c->stats.launches += launches;
    sh.phase = 3;
    return OGE_OK;
}

int oge_gpu_shard_route(oge_gpu_dedup_ctx *c, void **route_dev, uint64_t *n_route) {
    int rc = need_phase(c, 3, "shard_route");
    if (rc) return rc;
    if (!route_dev || !n_route) return fail_msg(OGE_ERR_INVALID_ARG, "shard_route: null argument");
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    uint64_t launches = 0, n_out = 0;
    sh.n_frag_total = c->n;
    if (c->cfg.world > 1 && (c->n || sh.n_pairs || sh.n_far)) {
        PhaseClock clk(c, &c->stats.ms_select);
        const ShardParams S = shard_params(c);
        // one sweep over the entries into a buffer sized for the usual case (boundary entries are a tiny
        // fraction); entries that found no room stay in place and a second sweep collects them
        {
            const char *e = getenv("OGE_ROUTE_CAP");      // test hook: force the second sweep
            const uint64_t want = e && *e ? (uint64_t) atoll(e) : std::max<uint64_t>(1u << 16, (c->n + sh.n_pairs + sh.n_far) / 64);
            if (e && *e) sh.route.release();
            if ((rc = sh.route.reserve(std::max<uint64_t>(want, 1), false, s))) return rc;
        }
        if ((rc = zero_counter(c, CNT_ROUTE)) || (rc = zero_counter(c, CNT_SCRATCH0)) || (rc = zero_counter(c, CNT_SCRATCH1))) return rc;
        for (int sweep = 0; sweep < 2; sweep++) {
            const uint32_t cap = (uint32_t) sh.route.cap;
            if ((rc = launch_sh_route(c->frag.p, c->n, 0, S, c->mate_of.p, sh.fm.p, 0, sh.route.p, cap, 0, s, &launches))) return rc;
            if ((rc = launch_sh_route(c->pair.p, sh.n_pairs, 1, S, c->mate_of.p, sh.fm.p, 0, sh.route.p, cap, 0, s, &launches))) return rc;
            if ((rc = launch_sh_route(c->pairf.p, sh.n_far, 2, S, c->mate_of.p, sh.fm.p, 0, sh.route.p, cap, 0, s, &launches))) return rc;
            if ((rc = read_counters(c))) return rc;
            n_out = c->h_counters[CNT_ROUTE];
            if (n_out <= cap) break;
            if (sweep == 1) return fail_msg(OGE_ERR_STATE, "shard_route: entry count changed between sweeps");
            if ((rc = sh.route.reserve(n_out, true, s))) return rc;
            OGE_CUDA_TRY(cudaMemcpyAsync(c->counters.p + CNT_ROUTE, &cap, 4, cudaMemcpyHostToDevice, s));      // continue behind what is stored
            OGE_CUDA_TRY(cudaStreamSynchronize(s));
        }
        const uint64_t routed_near = n_out ? c->h_counters[CNT_SCRATCH0] : 0, routed_far = n_out ? c->h_counters[CNT_SCRATCH1] : 0;
        sh.n_retracted += routed_near;                  // dead pair entries, whatever the reason
        sh.n_far_dead += routed_far;
        sh.n_frag -= n_out - routed_near - routed_far;
        clk.stop();
    }
    c->stats.launches += launches;
    *route_dev = sh.route.p;
    *n_route = n_out;
    sh.phase = 4;
    return OGE_OK;
}

int oge_gpu_shard_finish(oge_gpu_dedup_ctx *c, const void *route_all_dev, uint64_t n_all, void **marks_dev, uint64_t *n_marks) {
    int rc = need_phase(c, 4, "shard_finish");
    if (rc) return rc;
    if (!marks_dev || !n_marks || (n_all && !route_all_dev)) return fail_msg(OGE_ERR_INVALID_ARG, "shard_finish: null argument");
    cudaStream_t s = c->stream;
    ShardState &sh = c->sh;
    uint64_t launches = 0;
    const uint64_t n = c->n;
    PassTimer timer{c->pass_ev, 48, 0, 0};
    PassTimer *tp = c->cfg.profile_events ? &timer : nullptr;
    uint64_t extra = 0;
    if (n_all) {
        PhaseClock clk(c, &c->stats.ms_select);
        if ((rc = c->frag.reserve(n + n_all, true, s))) return rc;
        if ((rc = c->sortbuf.reserve(n + n_all, false, s))) return rc;
        if ((rc = c->pair.reserve(sh.n_pairs + n_all, true, s))) return rc;
        if ((rc = c->pair2.reserve(sh.n_pairs + n_all, false, s))) return rc;
        if ((rc = c->pairf.reserve(sh.n_far + n_all, true, s))) return rc;
        if ((rc = c->pairf2.reserve(sh.n_far + n_all, false, s))) return rc;
        if ((rc = sh.fm.reserve(sh.n_fm + n_all, true, s))) return rc;
        if ((rc = c->scratch.reserve(std::max(sort_scratch_bytes(std::max(std::max(n + n_all, sh.n_pairs + n_all), sh.n_far + n_all)), c->scratch.cap), true, s))) return rc;
        if (n == 0 && (rc = c->mate_of.reserve(1, false, s))) return rc;
        if ((rc = zero_counter(c, CNT_FRAG_EXTRA))) return rc;
        if ((rc = launch_sh_receive((const RouteEntry *) route_all_dev, n_all, shard_params(c), c->frag.p + n, (uint32_t) n_all, c->pair.p,
                                    (uint32_t) c->pair.cap, c->pairf.p, (uint32_t) c->pairf.cap, c->mate_of.p, sh.fm.p, (uint32_t) sh.fm.cap, s, &launches)))
            return rc;
        if ((rc = read_counters(c))) return rc;
        extra = c->h_counters[CNT_FRAG_EXTRA];
        sh.n_pairs = c->h_counters[CNT_PAIRS];
        sh.n_far = c->h_counters[CNT_PAIRS_FAR];
        sh.n_fm = c->h_counters[CNT_FM];
        clk.stop();
    }
    if (sh.n_fm > 1) {      // foreign mates sorted by idx1 for the binary search in K4
        PhaseClock clk(c, &c->stats.ms_select);
        if ((rc = sh.fm_sort.reserve(2 * sh.n_fm, false, s))) return rc;
        if ((rc = c->scratch.reserve(std::max(sort_scratch_bytes(sh.n_fm), c->scratch.cap), true, s))) return rc;
        if ((rc = launch_sh_fm_pack(sh.fm.p, (uint32_t) sh.n_fm, sh.fm_sort.p, s, &launches))) return rc;
        E128 *sorted = nullptr;
        if ((rc = radix_sort_128(sh.fm_sort.p, sh.fm_sort.p + sh.n_fm, sh.n_fm, nullptr, 32, 64, c->scratch.p, s, &sorted, &launches))) return rc;
        if ((rc = launch_sh_fm_unpack(sorted, (uint32_t) sh.n_fm, sh.fm.p, s, &launches))) return rc;
        clk.stop();
    }
    // marks on other ranks' records can only come from entries that crossed ranks: pairs formed by the
    // replay (at most one per two published entries) and routed entries
    if ((rc = sh.marks.reserve(sh.n_w + 2 * n_all + 1024, false, s))) return rc;
    if ((rc = zero_counter(c, CNT_FOREIGN_MARKS))) return rc;

    SelectParams sp;
    sp.dup = c->dup.p; sp.mate_of = c->mate_of.p; sp.idx_base = c->cfg.index_base; sp.n_records = n;
    sp.counters = c->counters.p; sp.kl = c->kl; sp.n_dev = nullptr;
    sp.fm = sh.fm.p; sp.n_fm = (uint32_t) sh.n_fm; sp.foreign_marks = sh.marks.p; sp.foreign_cap = (uint32_t) sh.marks.cap;
    const uint64_t n_pairs = sh.n_pairs, n_dead = sh.n_retracted;
    for (int far = 0; far < 2; far++) {      // near pairs (short key), then far pairs
        const uint64_t cnt = far ? sh.n_far : n_pairs, dead = far ? sh.n_far_dead : n_dead;
        if (!cnt) continue;
        E128 *a = far ? c->pairf.p : c->pair.p, *b = far ? c->pairf2.p : c->pair2.p, *sorted = a;
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
    const uint64_t n_frag_valid = sh.n_frag + extra;
    if (n_frag_valid) {
        E128 *sorted_frags = c->frag.p;
        {
            PhaseClock clk(c, &c->stats.ms_sort_frag);
            if ((rc = radix_sort_128(c->frag.p, c->sortbuf.p, n + extra, nullptr, c->kl.f_orient, c->kl.f_end, c->scratch.p, s, &sorted_frags,
                                     &launches, tp)))
                return rc;
            clk.stop();
        }
        PhaseClock clk(c, &c->stats.ms_select);
        sp.sorted = sorted_frags; sp.n_max = (uint32_t) n_frag_valid;
        if ((rc = launch_select_frags(sp, s, &launches))) return rc;
        clk.stop();
    }
    if ((rc = read_counters(c))) return rc;
    const uint64_t n_foreign = c->h_counters[CNT_FOREIGN_MARKS];
    if (n_foreign > sh.marks.cap) return fail_msg(OGE_ERR_STATE, "shard_finish: %llu marks for other ranks, room for %llu",
                                                  (unsigned long long) n_foreign, (unsigned long long) sh.marks.cap);
    c->stats.launches += launches;
    c->stats.n_frag_entries = n_frag_valid;
    c->stats.n_pair_entries = n_pairs - n_dead + sh.n_far - sh.n_far_dead;
    for (int i = 0; i < timer.used; i++) c->stats.ms_sort_pass_kernels += ms_between(c->pass_ev[2 * i], c->pass_ev[2 * i + 1]);
    c->stats.sort_pass_launches = timer.used;
    c->stats.sort_pass_bytes = timer.bytes;
    *marks_dev = sh.marks.p;
    *n_marks = n_foreign;
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
        fp.rec = c->rec.p; fp.off = c->off.p; fp.n = c->n; fp.flag_in = c->flag_in.p; fp.flag_out = c->flag_out.p;
        fp.dup = c->dup.p; fp.counters = c->counters.p; fp.quiet_index_bug = 0;
        if ((rc = launch_flags(fp, s, &launches))) return rc;
        clk.stop();
        if ((rc = read_counters(c))) return rc;
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
