// C ABI of the B200 duplicate-marking path (include/oge_gpu_dedup.h).
//
// A context keeps the pushed BAM records resident in HBM and runs, on oge_gpu_dedup_run():
//   K1 end-build -> K2 mate join (insert / resolve / exact slow path) -> K3 onesweep radix sort
//   of pair and fragment entries -> K4 group-and-select -> K5 flag write
// i.e. MarkDuplicates::runInternal (reference algorithms/mark_duplicates.cpp:422-475) without
// the temp-file spill.  There is no CPU fallback: every entry point fails with OGE_ERR_CUDA when
// no device is usable.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "ctx.cuh"

namespace oge {

static thread_local char g_err[512] = "";

int fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d: %s", (int) e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? OGE_ERR_NOMEM : OGE_ERR_CUDA;
}

int fail_msg(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int bit_length(uint64_t v) {
    int b = 0;
    while (v) { b++; v >>= 1; }
    return b;
}

}  // namespace oge

using namespace oge;

namespace oge {

int compute_layout(oge_gpu_dedup_ctx *c, KeyLayout *L) {
    memset(L, 0, sizeof(*L));
    uint64_t top = std::max<uint64_t>(c->cfg.index_base + c->n, c->sh.global_n);
    L->idx_bits = std::max(1, bit_length(top ? top - 1 : 0));
    L->ref_bits = std::max(1, bit_length(c->cfg.n_ref > 1 ? (uint64_t) c->cfg.n_ref - 1 : 1));
    if (c->cfg.max_ref_len > 0) {
        int64_t margin = c->cfg.clip_margin > 0 ? c->cfg.clip_margin : (1 << 20);
        uint64_t span = (uint64_t) c->cfg.max_ref_len + 2 * (uint64_t) margin;
        L->coord_bits = bit_length(span - 1);
        L->coord_bias = margin;
        if (L->coord_bits > 32) { L->coord_bits = 32; L->coord_bias = 1ll << 31; }
    } else {
        L->coord_bits = 32;      // the whole int32 range
        L->coord_bias = 1ll << 31;
    }
    L->lib_bits = bit_length((uint64_t) c->n_libs + 1);
    L->lib_invalid = (1u << L->lib_bits) - 1;
    int b = 16;
    L->f_idx = b; b += L->idx_bits;
    L->f_paired = b; b += 1;
    L->f_orient = b; b += 1;
    L->f_coord = b; b += L->coord_bits;
    L->f_ref = b; b += L->ref_bits;
    L->f_lib = b; b += L->lib_bits;
    L->f_end = b;
    L->p_idx = 16;
    L->p_end = 128;
    L->p_lib = 128 - L->lib_bits;
    L->p_ref1 = L->p_lib - L->ref_bits;
    L->p_coord1 = L->p_ref1 - L->coord_bits;
    L->p_orient = L->p_coord1 - 2;
    L->p_ref2 = L->p_orient - L->ref_bits;
    L->p_coord2 = L->p_ref2 - L->coord_bits;
    // near pairs: the distance field is sized so that the key is a whole number of 8-bit digits (>= 11 bits)
    {
        const int fixed = L->lib_bits + L->ref_bits + L->coord_bits + 2;
        int d = 11;
        while ((fixed + d) % 8) d++;
        if (d > L->coord_bits) d = L->coord_bits;
        L->delta_bits = d;
        L->n_delta = L->p_orient - d;
    }
    {
        // the 64-bit word form of the pair-entry builder (pairing.cuh: make_pair_entry) needs: the [coord][ref][lib] block
        // of a fragment entry reachable with one two-word shift and at most 62 bits wide, the ordinal inside the low word, the
        // near key inside the high word, far fields that do not overlap the payload
        const int block = L->coord_bits + L->ref_bits + L->lib_bits;
        L->fast = L->f_coord > 0 && L->f_coord < 64 && L->f_orient < 64 && block <= 62 && 16 + L->idx_bits <= 64 &&
                  block + 2 + L->delta_bits <= 64 && L->n_delta >= 64 && L->coord_bits + L->ref_bits <= 62;
#ifdef OGE_TESTING
        if (getenv("OGE_GENERIC_PAIR_ENTRY")) L->fast = 0;      // test hook: the field-by-field form
#endif
    }
    if (L->f_end > 128 || L->p_coord2 < 16 + L->idx_bits)
        return fail_msg(OGE_ERR_KEY_RANGE,
                        "key layout needs %d (frag) / %d (pair) bits, more than the 128 of a 16-byte entry: "
                        "idx %d, coord %d, ref %d, lib %d bits (set max_ref_len / n_ref in the config)",
                        L->f_end, 16 + L->idx_bits + 128 - L->p_coord2, L->idx_bits, L->coord_bits, L->ref_bits, L->lib_bits);
    return 0;
}

RgTable rg_table(oge_gpu_dedup_ctx *c) {
    RgTable t;
    t.bytes = c->rg_bytes.p;
    t.off = c->rg_off.p;
    t.lib = c->rg_lib.p;
    t.n = c->n_rg;
    t.unknown_lib = c->unknown_lib;
    return t;
}

__global__ void offsets_rebase(uint64_t *off, uint64_t n, uint64_t base) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] += base;
}

int ensure_work(oge_gpu_dedup_ctx *c) {
    const uint64_t n = c->n;
    cudaStream_t s = c->stream;
    int rc;
    if ((rc = c->frag.reserve(n, false, s))) return rc;
    if ((rc = c->sortbuf.reserve(n, false, s))) return rc;
    if ((rc = c->hk.reserve(n, false, s))) return rc;
    if ((rc = c->tag.reserve(n, false, s))) return rc;
    if ((rc = c->flag_in.reserve(n, false, s))) return rc;
    if ((rc = c->flag_out.reserve(n, false, s))) return rc;
    if ((rc = c->dup.reserve(n, false, s))) return rc;
    if ((rc = c->mate_of.reserve(n, false, s))) return rc;
    if ((rc = c->scratch.reserve(std::max(sort_scratch_bytes(n), compact_scratch_bytes(n)), false, s))) return rc;
    return 0;
}

int check_endbuild_errors(oge_gpu_dedup_ctx *c) {
    if (c->h_counters[CNT_ERR] & DEV_ERR_BAD_RECORD)
        return fail_msg(OGE_ERR_BAD_RECORD, "run: malformed record (block_size disagrees with offsets, or sections overrun the record)");
    if (c->h_counters[CNT_ERR] & DEV_ERR_KEY_RANGE)
        return fail_msg(OGE_ERR_KEY_RANGE, "run: a record's refID / unclipped coordinate / library does not fit the key layout "
                                           "(n_ref=%d max_ref_len=%d clip_margin=%d)", c->cfg.n_ref, c->cfg.max_ref_len, c->cfg.clip_margin);
    if (c->h_counters[CNT_ERR] & DEV_ERR_CAPACITY)
        return fail_msg(OGE_ERR_STATE, "run: internal error, a pair list overran its capacity");
    if ((c->cfg.index_base + c->n > (1ull << 32)) || c->sh.global_n > (1ull << 32))
        return fail_msg(OGE_ERR_TOO_LARGE, "run: global record ordinals must stay below 2^32");
    return 0;
}

float ms_between(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

}  // namespace oge

extern "C" {

int oge_gpu_abi_version(void) { return OGE_GPU_DEDUP_ABI_VERSION; }

int oge_gpu_sizeof(int which) {
    switch (which) {
        case 0: return (int) sizeof(oge_gpu_dedup_config);
        case 1: return (int) sizeof(oge_gpu_dedup_stats);
        case 2: return (int) sizeof(oge_gpu_end);
        case 3: return (int) sizeof(oge_gpu_flagstats);
        default: return -1;
    }
}

const char *oge_gpu_last_error(void) { return g_err; }

int oge_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void *oge_gpu_host_alloc(size_t nbytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void oge_gpu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int oge_gpu_dedup_create(const oge_gpu_dedup_config *cfg, oge_gpu_dedup_ctx **out) {
    if (!cfg || !out) return fail_msg(OGE_ERR_INVALID_ARG, "create: null argument");
    if (cfg->abi_version != OGE_GPU_DEDUP_ABI_VERSION)
        return fail_msg(OGE_ERR_INVALID_ARG, "create: ABI version %d, library is %d", cfg->abi_version, OGE_GPU_DEDUP_ABI_VERSION);
    int ndev = oge_gpu_device_count();
    if (ndev <= 0) return fail_msg(OGE_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail_msg(OGE_ERR_INVALID_ARG, "create: device %d of %d", cfg->device, ndev);
    OGE_CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    OGE_CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10)
        return fail_msg(OGE_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    oge_gpu_dedup_ctx *c = new oge_gpu_dedup_ctx();
    c->cfg = *cfg;
#ifdef OGE_TESTING
    if (c->cfg.verify_names < 0) c->cfg.verify_names = 1;
#else
    c->cfg.verify_names = 1;      // the product library always pairs by the key's bytes: hash-only pairing is a measurement knob
#endif
    c->sms = prop.multiProcessorCount;
    memset(&c->stats, 0, sizeof(c->stats));
    int rc = 0;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->inflate_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->sh.ev_main, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->sh.ev_far, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreate(&c->sh.ev_side[0]) != cudaSuccess || cudaEventCreate(&c->sh.ev_side[1]) != cudaSuccess ||
            cudaEventCreate(&c->sh.ev_side[2]) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming) != cudaSuccess) {
            rc = fail_cuda(cudaGetLastError(), "stream/event create", __FILE__, __LINE__);
            break;
        }
        for (auto &e : c->ev) cudaEventCreate(&e);
        for (auto &e : c->ev_piece) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        for (auto &e : c->ev_stage) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        for (auto &e : c->clk_ev) cudaEventCreate(&e);
        for (auto &e : c->pass_ev) { e = nullptr; if (cfg->profile_events) cudaEventCreate(&e); }
        for (auto &e : c->k_ev) { e = nullptr; if (cfg->profile_events) cudaEventCreate(&e); }
        if (cudaHostAlloc((void **) &c->h_counters, CNT_N * 4, cudaHostAllocDefault) != cudaSuccess) {
            rc = fail_cuda(cudaGetLastError(), "cudaHostAlloc", __FILE__, __LINE__);
            break;
        }
        if ((rc = c->counters.reserve(CNT_N, false, c->stream))) break;
        if ((rc = radix_sort_init())) break;
        if (cfg->capacity_bytes && (rc = c->rec.reserve(cfg->capacity_bytes, false, c->stream))) break;
        if (cfg->capacity_records && (rc = c->off.reserve(cfg->capacity_records + 1, false, c->stream))) break;
    } while (0);
    if (rc) {
        oge_gpu_dedup_destroy(c);
        return rc;
    }
    *out = c;
    return OGE_OK;
}

void oge_gpu_dedup_destroy(oge_gpu_dedup_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->side_stream) cudaStreamSynchronize(c->side_stream);
    c->zcomp.release(); c->zoff.release(); c->zcs.release(); c->zfile.release(); c->frame_w.release(); c->frame_wb.release();
    c->rec.release(); c->off.release(); c->rg_bytes.release(); c->rg_off.release(); c->rg_lib.release();
    c->frag.release(); c->sortbuf.release(); c->pair.release(); c->pair2.release(); c->pairf.release(); c->pairf2.release(); c->ufrag.release(); c->ufrag2.release(); c->uset.release(); c->hk.release();
    c->tag.release(); c->flag_in.release(); c->flag_out.release(); c->dup.release(); c->scratch.release();
    c->cplx_state.release(); c->cplx_slots.release(); c->cplx_sort.release(); c->pair_hk.release(); c->pairf_hk.release(); c->left.release(); c->couples.release(); c->couple_count.release();
    c->cs_u32.release(); c->perm.release(); c->cs_bsum.release(); c->off2.release(); c->rec2.release(); c->mate_of.release(); c->counters.release(); c->table.release();
    if (c->h_counters) cudaFreeHost(c->h_counters);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->ev_piece) if (e) cudaEventDestroy(e);
    for (auto &e : c->ev_stage) if (e) cudaEventDestroy(e);
    for (auto &h : c->h_stage) if (h) cudaFreeHost(h);
    for (auto &e : c->clk_ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->pass_ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->k_ev) if (e) cudaEventDestroy(e);
    if (c->copy_done) cudaEventDestroy(c->copy_done);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->inflate_stream) cudaStreamDestroy(c->inflate_stream);
    if (c->sh.ev_main) cudaEventDestroy(c->sh.ev_main);
    if (c->sh.ev_far) cudaEventDestroy(c->sh.ev_far);
    for (auto &e : c->sh.ev_side) if (e) cudaEventDestroy(e);
    oge_gpu_shard_comm_destroy(c);
    c->sh.x_cnt.release(); c->sh.r_pub.release(); c->sh.r_pub2.release(); c->sh.r_froute.release(); c->sh.r_hash.release();
    c->sh.r_proute.release(); c->sh.r_oroute.release(); c->sh.r_marks.release();
    c->sh.d_bases.release(); c->sh.pub_hash.release(); c->sh.hset.release(); c->sh.pub_raw.release(); c->sh.pub_send.release();
    c->sh.froute_send.release(); c->sh.proute_send.release(); c->sh.oroute_send.release(); c->sh.marks_send.release(); c->sh.bk.release();
    c->sh.d_split.release(); c->sh.pub.release(); c->sh.pub2.release(); c->sh.route.release(); c->sh.froute.release();
    c->sh.marks.release(); c->sh.marks_frag.release(); c->sh.pub_list.release(); c->sh.fm.release(); c->sh.fm_sort.release();
    c->sh.w_sort.release(); c->sh.w_sort2.release(); c->sh.scratch2.release();
    delete c;
}

int oge_gpu_dedup_set_readgroups(oge_gpu_dedup_ctx *c, const char *const *ids, const int16_t *lib_ids, int32_t n,
                                 int16_t unknown_lib_id, int32_t n_libs) {
    if (!c || n < 0 || (n > 0 && (!ids || !lib_ids))) return fail_msg(OGE_ERR_INVALID_ARG, "set_readgroups: bad argument");
    if (n >= (int32_t) RGC_UNKNOWN) return fail_msg(OGE_ERR_INVALID_ARG, "set_readgroups: more than %u read groups", RGC_UNKNOWN - 1);
    if (n_libs < 1 || unknown_lib_id < 1 || unknown_lib_id > n_libs)
        return fail_msg(OGE_ERR_INVALID_ARG, "set_readgroups: library ids must lie in 1..n_libs");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> off(1, 0);
    std::vector<int16_t> lib;
    for (int i = 0; i < n; i++) {
        if (!ids[i]) return fail_msg(OGE_ERR_INVALID_ARG, "set_readgroups: null id");
        if (lib_ids[i] < 1 || lib_ids[i] > n_libs) return fail_msg(OGE_ERR_INVALID_ARG, "set_readgroups: library id %d out of 1..%d", lib_ids[i], n_libs);
        size_t l = strlen(ids[i]);
        bytes.insert(bytes.end(), ids[i], ids[i] + l);
        off.push_back((uint32_t) bytes.size());
        lib.push_back(lib_ids[i]);
    }
    int rc;
    if ((rc = c->rg_bytes.reserve(bytes.size() + 1, false, c->stream))) return rc;
    if ((rc = c->rg_off.reserve(off.size(), false, c->stream))) return rc;
    if ((rc = c->rg_lib.reserve(lib.size() + 1, false, c->stream))) return rc;
    if (!bytes.empty()) OGE_CUDA_TRY(cudaMemcpy(c->rg_bytes.p, bytes.data(), bytes.size(), cudaMemcpyHostToDevice));
    OGE_CUDA_TRY(cudaMemcpy(c->rg_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    if (!lib.empty()) OGE_CUDA_TRY(cudaMemcpy(c->rg_lib.p, lib.data(), lib.size() * 2, cudaMemcpyHostToDevice));
    c->n_rg = n;
    c->unknown_lib = unknown_lib_id;
    c->n_libs = n_libs;
    return OGE_OK;
}

static uint64_t g_bgzf_chunk_bytes = 0;      // 0: per-mode default (oge_gpu_set_bgzf_chunk_bytes)
static uint64_t g_bgzf_stage_bytes = 32ull << 20;      // pageable sources go through two pinned staging buffers of this size; 0: off (oge_gpu_set_bgzf_staging)

int oge_gpu_dedup_push_bgzf(oge_gpu_dedup_ctx *c, const uint8_t *comp, uint64_t comp_bytes, const uint64_t *block_in_off,
                            const uint32_t *block_csize, const uint32_t *block_isize, uint64_t n_blocks, uint64_t header_bytes,
                            uint8_t *host_copy) {
    if (!c || !comp || !block_in_off || !block_csize || !block_isize) return fail_msg(OGE_ERR_INVALID_ARG, "push_bgzf: null argument");
    if (c->n || c->rec_bytes) return fail_msg(OGE_ERR_STATE, "push_bgzf: the context already holds records (oge_gpu_dedup_reset first)");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    std::vector<uint64_t> out_off(n_blocks + 1);
    uint64_t total = 0, prev_end = 0;
    bool in_file_order = true;      // blocks back to back in the order of the table (any BGZF file): the upload can go in pieces
    for (uint64_t b = 0; b < n_blocks; b++) {
        if (block_csize[b] < 26 || block_isize[b] > 65536 || block_in_off[b] + block_csize[b] > comp_bytes)
            return fail_msg(OGE_ERR_BAD_RECORD, "push_bgzf: block %llu does not fit the file or inflates to more than 65536 bytes", (unsigned long long) b);
        if (block_in_off[b] < prev_end) in_file_order = false;
        prev_end = block_in_off[b] + block_csize[b];
        out_off[b] = total;
        total += block_isize[b];
    }
    out_off[n_blocks] = total;
    if (header_bytes > total) return fail_msg(OGE_ERR_INVALID_ARG, "push_bgzf: header_bytes beyond the inflated stream");
    const uint64_t rec_bytes = total - header_bytes;
    const uint64_t lead = (header_bytes + 255) & ~255ull;      // the first record lands on a 256-byte boundary of the buffer
    cudaStream_t s = c->stream, up = c->copy_stream, down = c->side_stream;
    const int mode = inflate_mode_for(c->cfg.device);
    const bool engine = mode == INFLATE_ENGINE;
    int rc;
    // the engine is not bounded by a block's ISIZE (bgzf_inflate.cu): room for the longest operation behind the last block
    if ((rc = c->rec.reserve(lead + rec_bytes + (engine ? engine_inflate_slack(c->cfg.device) : 0), false, s))) return rc;
    // staging of the compressed file and its block table: kept by the context (cudaFree of 10 GB costs 240 ms and cudaMalloc
    // 70 ms on the B200, more than the whole upload saves; oge_gpu_dedup_reset keeps them, destroy frees them)
    DevBuf<uint8_t> &zcomp = c->zcomp;
    DevBuf<uint64_t> &zoff = c->zoff;       // in_off[n_blocks], out_off[n_blocks + 1]
    DevBuf<uint32_t> &zcs = c->zcs;         // csize[n_blocks], err[2], act[n_blocks]
    auto done = [&](int code) { return code; };
    if ((rc = zcomp.reserve(comp_bytes + 64, false, s)) || (rc = zoff.reserve(2 * n_blocks + 1, false, s)) || (rc = zcs.reserve(2 * n_blocks + 2, false, s)))
        return done(rc);
    uint32_t *d_err = zcs.p + n_blocks, *d_act = zcs.p + n_blocks + 2;
    uint8_t *d_out = c->rec.p + (lead - header_bytes);
    // ---- the pieces: the upload of piece k + 1 (copy stream) runs under the inflate of piece k (main stream) and the copy-back
    //      of piece k - 1 (side stream, host_copy only).  Piece size: a few inflate-milliseconds, so that the last piece's inflate
    //      -- the only part of it the upload does not hide -- is small, but enough blocks to fill the machine for the kernels.
    uint64_t piece = g_bgzf_chunk_bytes ? g_bgzf_chunk_bytes : (engine ? (64ull << 20) : (512ull << 20));
    if (!in_file_order) piece = ~0ull;
    // Streams.  The inflates go on a stream of their own that never carries a host-to-device copy: measured on the B200
    // (profiles/r2_push_bgzf_stream_trace.txt), engine work submitted on a stream that has done H2D copies is not started
    // while ANOTHER stream's H2D copies are queued -- piece 0's inflate began when the last piece's upload ended -- whereas on
    // a stream that has only ever seen engine work and memsets it starts the moment its own piece has landed.
    cudaStream_t zs = c->inflate_stream;
    cudaEvent_t t_up0 = c->ev[0], t_up1 = c->ev[1], t_inf0 = c->ev[8], t_inf1 = c->ev[9], t_all0 = c->ev[2], t_all1 = c->ev[3], t_dn0 = c->ev[4], t_dn1 = c->ev[5];
    cudaError_t e = cudaEventRecord(t_all0, s);
    // the three side streams start behind whatever the main stream was doing with these buffers
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_piece[0], s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(up, c->ev_piece[0], 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(down, c->ev_piece[0], 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(zs, c->ev_piece[0], 0);
    if (e == cudaSuccess) e = cudaMemsetAsync(zcomp.p + comp_bytes, 0, 64, zs);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_err, 0, 8, zs);
    if (e == cudaSuccess && engine) e = cudaMemsetAsync(d_act, 0, n_blocks * 4, zs);
    if (e == cudaSuccess) e = cudaEventRecord(t_up0, up);
    if (e == cudaSuccess) e = cudaEventRecord(t_dn0, down);
    // block tables: in front of the pieces on the copy stream (the kernels read them on the device; the engine takes the
    // table from the host and only the size check reads out_off)
    if (e == cudaSuccess) e = cudaMemcpyAsync(zoff.p + n_blocks, out_off.data(), (n_blocks + 1) * 8, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess && !engine) {
        e = cudaMemcpyAsync(zoff.p, block_in_off, n_blocks * 8, cudaMemcpyHostToDevice, up);
        if (e == cudaSuccess) e = cudaMemcpyAsync(zcs.p, block_csize, n_blocks * 4, cudaMemcpyHostToDevice, up);
    }
    if (e != cudaSuccess) return done(fail_cuda(e, "push_bgzf setup", __FILE__, __LINE__));
    uint64_t launches = 0, pieces = 0, engine_ops = 0, n_staged = 0;
    // where the file lies: page-locked memory goes up as it is, pageable memory through the context's staging buffers
    const uint64_t STAGE_BYTES = g_bgzf_stage_bytes;
    constexpr int STAGE_THREADS = 8;
    bool staged = false;
    {
        cudaPointerAttributes pa;
        const cudaError_t q = cudaPointerGetAttributes(&pa, comp);
        if (q != cudaSuccess) cudaGetLastError();
        staged = (q != cudaSuccess || pa.type == cudaMemoryTypeUnregistered) && STAGE_BYTES && comp_bytes >= 2 * STAGE_BYTES;
        if (staged && c->h_stage_bytes != STAGE_BYTES) {      // first use, or the size was changed
            for (auto &h : c->h_stage) { if (h) cudaFreeHost(h); h = nullptr; }
            c->h_stage_bytes = STAGE_BYTES;
        }
        for (int k = 0; staged && k < 2; k++) {
            if (!c->h_stage[k] && cudaHostAlloc((void **) &c->h_stage[k], STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                c->h_stage[k] = nullptr;
                staged = false;      // no pinned memory to be had: the driver's pageable path
            }
        }
    }
#ifdef OGE_TESTING
    const bool trace = getenv("OGE_TRACE_PUSH") != nullptr;      // measurement hook of the -DOGE_TESTING build: per-piece device timestamps on stderr
#else
    const bool trace = false;
#endif
    auto wall = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
    const double w0 = wall();
    std::vector<cudaEvent_t> tr_ev;      // trace: per piece upload done / inflate start / inflate done
    for (uint64_t b0 = 0; b0 < n_blocks && e == cudaSuccess;) {
        uint64_t b1 = b0, bytes = 0;
        while (b1 < n_blocks && (bytes < piece || b1 == b0)) bytes += block_csize[b1++];
        // file bytes of the piece (the whole file when the table is not in file order)
        const uint64_t lo = in_file_order ? block_in_off[b0] : 0, hi = in_file_order ? block_in_off[b1 - 1] + block_csize[b1 - 1] : comp_bytes;
        if (!staged) {
            e = cudaMemcpyAsync(zcomp.p + lo, comp + lo, hi - lo, cudaMemcpyHostToDevice, up);
        } else {
            // pageable source: through the two pinned staging buffers, filled by several host threads while the previous one is
            // on its way (the driver's own path for pageable memory is one staging copy stream: 11 GB/s on the B200 box)
            for (uint64_t a = lo; a < hi && e == cudaSuccess; a += STAGE_BYTES, n_staged++) {
                const uint64_t len = hi - a < STAGE_BYTES ? hi - a : STAGE_BYTES;
                const int slot = (int) (n_staged & 1);
                if (n_staged >= 2) e = cudaEventSynchronize(c->ev_stage[slot]);      // its previous upload has left the buffer
                if (e != cudaSuccess) break;
                uint8_t *dst = c->h_stage[slot];
                const int nt = (int) (len >= (8u << 20) ? STAGE_THREADS : 1);
                const uint64_t per = (len + nt - 1) / nt;
                std::vector<std::thread> pool;
                for (int t = 1; t < nt; t++)
                    pool.emplace_back([=] { const uint64_t x = per * t, y = x + per < len ? x + per : len; if (y > x) memcpy(dst + x, comp + a + x, y - x); });
                memcpy(dst, comp + a, per < len ? per : len);
                for (auto &th : pool) th.join();
                e = cudaMemcpyAsync(zcomp.p + a, dst, len, cudaMemcpyHostToDevice, up);
                if (e == cudaSuccess) e = cudaEventRecord(c->ev_stage[slot], up);
            }
        }
        if (trace) {
            cudaEvent_t ta, tb, tc;
            cudaEventCreate(&ta); cudaEventCreate(&tb); cudaEventCreate(&tc);
            tr_ev.push_back(ta); tr_ev.push_back(tb); tr_ev.push_back(tc);
            cudaEventRecord(ta, up);
        }
        if (e == cudaSuccess) e = cudaEventRecord(c->ev_piece[1], up);      // re-recorded per piece: a wait binds to the record before it
        if (e == cudaSuccess) e = cudaStreamWaitEvent(zs, c->ev_piece[1], 0);
        if (trace) cudaEventRecord(tr_ev[tr_ev.size() - 2], zs);
        if (e == cudaSuccess && pieces == 0) e = cudaEventRecord(t_inf0, zs);
        if (e != cudaSuccess) break;
        if (engine) {
            if ((rc = engine_inflate_submit(zcomp.p, block_in_off, block_csize, block_isize, out_off.data(), d_out, d_act, b0, b1, zs, &engine_ops))) {
                cudaStreamSynchronize(up);
                cudaStreamSynchronize(zs);
                return done(rc);
            }
        } else {
            BgzfParams P;
            P.comp = zcomp.p;
            P.in_off = zoff.p + b0;
            P.out_off = zoff.p + n_blocks + b0;
            P.csize = zcs.p + b0;
            P.n_blocks = b1 - b0;
            P.block_base = b0;
            P.out = d_out;
            P.err = d_err;
            if ((rc = launch_bgzf_inflate(P, mode, c->sms, zs, &launches))) {
                cudaStreamSynchronize(up);
                cudaStreamSynchronize(zs);
                return done(rc);
            }
        }
        if (trace) cudaEventRecord(tr_ev[tr_ev.size() - 1], zs);
        if (host_copy && rec_bytes) {
            // inflated bytes of the piece that belong to the record array -> the caller's buffer
            const uint64_t o0 = std::max(out_off[b0], header_bytes), o1 = std::max(out_off[b1], header_bytes);
            if (o1 > o0) {
                e = cudaEventRecord(c->ev_piece[2], zs);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(down, c->ev_piece[2], 0);
                if (e == cudaSuccess) e = cudaMemcpyAsync(host_copy + (o0 - header_bytes), d_out + o0, o1 - o0, cudaMemcpyDeviceToHost, down);
            }
        }
        pieces++;
        b0 = b1;
    }
    const double w1 = wall();
    uint32_t err[2] = {0, 0};
    if (e == cudaSuccess && pieces == 0) e = cudaEventRecord(t_inf0, zs);      // a file without blocks: the clocks still read
    if (e == cudaSuccess) e = cudaEventRecord(t_up1, up);
    if (e == cudaSuccess) e = cudaEventRecord(t_inf1, zs);
    if (e == cudaSuccess) e = cudaEventRecord(t_dn1, down);
    // back on the main stream: the size check (engine), the verdict, the end
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, t_inf1, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, t_up1, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, t_dn1, 0);
    if (e == cudaSuccess && engine && (rc = launch_bgzf_check_sizes(d_act, zoff.p + n_blocks, n_blocks, d_err, s, &launches))) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMemcpyAsync(err, d_err, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaEventRecord(t_all1, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        cudaStreamSynchronize(up);
        cudaStreamSynchronize(down);
        cudaStreamSynchronize(zs);
        // the engine's verdict on a stream that is not valid deflate is a launch failure, which surfaces at whichever call
        // looks first and takes the context with it (bgzf_inflate.cu)
        if (engine && (e == cudaErrorLaunchFailure || cudaDeviceSynchronize() == cudaErrorLaunchFailure)) {
            cudaGetLastError();
            return done(fail_msg(OGE_ERR_BAD_RECORD, "Zlib inflate failed (a BGZF block is not a valid deflate stream; reported by the hardware decompress "
                                                     "engine as a launch failure, the CUDA context of this process is lost: select a kernel decoder, "
                                                     "OGE_INFLATE_KERNEL=warp, to locate the block)."));
        }
        return done(fail_cuda(e, "push_bgzf inflate", __FILE__, __LINE__));
    }
    if (trace) {
        for (size_t k = 0; k < tr_ev.size() / 3; k++)
            if (k < 4 || k + 2 >= tr_ev.size() / 3 || k == tr_ev.size() / 6)
                fprintf(stderr, "push_bgzf trace: piece %zu: upload done at %.2f ms, inflate start %.2f, inflate done %.2f\n", k, ms_between(t_all0, tr_ev[3 * k]),
                        ms_between(t_all0, tr_ev[3 * k + 1]), ms_between(t_all0, tr_ev[3 * k + 2]));
        for (auto &ev : tr_ev) cudaEventDestroy(ev);
    }
    if (trace)
        fprintf(stderr, "push_bgzf trace: %llu pieces issued in %.2f ms, drained after %.2f ms more\n", (unsigned long long) pieces, w1 - w0, wall() - w1);
    if (err[0]) return done(fail_msg(OGE_ERR_BAD_RECORD, "Zlib inflate failed (BGZF block %u, code %u).", err[1], err[0]));
    c->stats.ms_inflate = ms_between(t_inf0, t_inf1);
    c->stats.ms_inflate_h2d = ms_between(t_up0, t_up1);
    c->stats.ms_inflate_d2h = host_copy ? ms_between(t_dn0, t_dn1) : 0.f;
    c->stats.ms_push_bgzf = ms_between(t_all0, t_all1);
    c->stats.ms_inflate_start = ms_between(t_all0, t_inf0);
    c->stats.inflate_mode = (uint32_t) mode;
    c->stats.inflate_pieces = (uint32_t) pieces;
    c->stats.ms_frame = 0;
    c->stats.frame_repairs = 0;
    c->stats.inflate_blocks = n_blocks;
    c->stats.inflate_bytes_in = comp_bytes;
    c->stats.inflate_bytes_out = total;
    c->rec_lead = lead;
    c->rec_bytes = rec_bytes;
    c->n = 0;
    c->ran = false;
    c->sorted = false;
    return done(OGE_OK);
}

int oge_gpu_dedup_frame(oge_gpu_dedup_ctx *c, uint64_t *nrec_out) {
    if (!c || !nrec_out) return fail_msg(OGE_ERR_INVALID_ARG, "frame: null argument");
    if (c->n || !c->rec_lead) return fail_msg(OGE_ERR_STATE, "frame: follows oge_gpu_dedup_push_bgzf, once");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream;
    *nrec_out = 0;
    if (c->rec_bytes == 0) {
        int rc0 = c->off.reserve(1, false, s);
        if (rc0) return rc0;
        OGE_CUDA_TRY(cudaMemsetAsync(c->off.p, 0, 8, s));
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
        return OGE_OK;
    }
    FrameParams P;
    P.rec = c->recs();
    P.total = c->rec_bytes;
    P.chunk = 1 << 16;
    P.n_chunks = (P.total + P.chunk - 1) / P.chunk;
    P.n_ref = c->cfg.n_ref;
    const uint64_t nc = P.n_chunks;
    // entry, exit, count, base per chunk: kept by the context (a cudaFree in this path was seen to take 100 ms on a busy box)
    DevBuf<uint64_t> &w = c->frame_w;
    DevBuf<uint32_t> &wb = c->frame_wb;
    int rc;
    auto done = [&](int code) { return code; };
    if ((rc = w.reserve(4 * nc, false, s)) || (rc = wb.reserve(nc, false, s))) return done(rc);
    P.entry = w.p;
    P.exit_ = w.p + nc;
    P.count = w.p + 2 * nc;
    uint64_t *d_base = w.p + 3 * nc;
    P.base = d_base;
    P.bad = wb.p;
    P.off = nullptr;
    uint64_t launches = 0;
    cudaEvent_t e0 = c->ev[8], e1 = c->ev[9];
    cudaError_t e = cudaEventRecord(e0, s);
    if (e != cudaSuccess) return done(fail_cuda(e, "frame", __FILE__, __LINE__));
    if ((rc = launch_frame_guess(P, s, &launches)) || (rc = launch_frame_walk(P, 0, nc, 0, s, &launches))) return done(rc);
    std::vector<uint64_t> h(3 * nc);
    std::vector<uint32_t> hb(nc);
    e = cudaMemcpyAsync(h.data(), w.p, 3 * nc * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hb.data(), wb.p, nc * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return done(fail_cuda(e, "frame walk", __FILE__, __LINE__));
    uint64_t *entry = h.data(), *exit_ = h.data() + nc, *count = h.data() + 2 * nc;
    // ---- the proof: every chunk must have been entered where its predecessor left; repair the ones that were not
    uint64_t repairs = 0;
    for (uint64_t k = 0; k < nc; k++) {
        const uint64_t want = k ? exit_[k - 1] : 0;
        if (entry[k] != want) {
            repairs++;
            e = cudaMemcpyAsync(w.p + k, &want, 8, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return done(fail_cuda(e, "frame repair", __FILE__, __LINE__));
            if ((rc = launch_frame_walk(P, k, 1, 0, s, &launches))) return done(rc);
            e = cudaMemcpyAsync(&exit_[k], w.p + nc + k, 8, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(&count[k], w.p + 2 * nc + k, 8, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(&hb[k], wb.p + k, 4, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) return done(fail_cuda(e, "frame repair", __FILE__, __LINE__));
            entry[k] = want;
        }
        if (hb[k]) {      // the TRUE chain breaks here: the reference's messages (util/bam_deserializer.h:155-170)
            if ((hb[k] & 3) == 2) return done(fail_msg(OGE_ERR_BAD_RECORD, "Invalid BAM block size(%u).", hb[k] >> 2));
            return done(fail_msg(OGE_ERR_BAD_RECORD, "Expected more bytes reading BAM core. Is this file truncated or corrupted?"));
        }
    }
    std::vector<uint64_t> base(nc);
    uint64_t n = 0;
    for (uint64_t k = 0; k < nc; k++) {
        base[k] = n;
        n += count[k];
    }
    if (n >= (1ull << 30)) return done(fail_msg(OGE_ERR_TOO_LARGE, "frame: more than 2^30-1 records in one context"));
    if ((rc = c->off.reserve(n + 1, false, s))) return done(rc);
    e = cudaMemcpyAsync(d_base, base.data(), nc * 8, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return done(fail_cuda(e, "frame bases", __FILE__, __LINE__));
    P.off = c->off.p;
    if ((rc = launch_frame_walk(P, 0, nc, 1, s, &launches))) return done(rc);
    e = cudaEventRecord(e1, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return done(fail_cuda(e, "frame write", __FILE__, __LINE__));
    c->stats.ms_frame = ms_between(e0, e1);
    c->stats.frame_repairs = repairs;
    c->n = n;
    c->ran = false;
    c->sorted = false;
    *nrec_out = n;
    return done(OGE_OK);
}

int oge_gpu_dedup_offsets(oge_gpu_dedup_ctx *c, uint64_t *out, uint64_t n_plus_1) {
    if (!c || !out) return fail_msg(OGE_ERR_INVALID_ARG, "offsets: null argument");
    if (n_plus_1 != c->n + 1) return fail_msg(OGE_ERR_INVALID_ARG, "offsets: the context holds %llu records", (unsigned long long) c->n);
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    OGE_CUDA_TRY(cudaMemcpyAsync(out, c->off.p, n_plus_1 * 8, cudaMemcpyDeviceToHost, c->stream));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OGE_OK;
}

int oge_gpu_dedup_set_offsets(oge_gpu_dedup_ctx *c, const uint64_t *offsets, uint64_t nrec) {
    if (!c || !offsets) return fail_msg(OGE_ERR_INVALID_ARG, "set_offsets: null argument");
    if (c->n || !c->rec_lead) return fail_msg(OGE_ERR_STATE, "set_offsets: follows oge_gpu_dedup_push_bgzf, once");
    if (offsets[0] != 0 || offsets[nrec] != c->rec_bytes) return fail_msg(OGE_ERR_BAD_RECORD, "set_offsets: offsets[0] must be 0 and offsets[nrec] the inflated record bytes");
    if (nrec >= (1ull << 30)) return fail_msg(OGE_ERR_TOO_LARGE, "set_offsets: more than 2^30-1 records in one context");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    int rc;
    if ((rc = c->off.reserve(nrec + 1, false, c->copy_stream))) return rc;
    OGE_CUDA_TRY(cudaMemcpyAsync(c->off.p, offsets, (nrec + 1) * 8, cudaMemcpyHostToDevice, c->copy_stream));
    c->n = nrec;
    c->ran = false;
    c->sorted = false;
    return OGE_OK;
}

int oge_gpu_dedup_push(oge_gpu_dedup_ctx *c, const uint8_t *records, uint64_t nbytes, const uint64_t *offsets, uint64_t nrec) {
    if (!c || (nrec && (!records || !offsets))) return fail_msg(OGE_ERR_INVALID_ARG, "push: null argument");
    if (nrec == 0) return OGE_OK;
    if (c->rec_lead) return fail_msg(OGE_ERR_STATE, "push: the context holds an inflated BGZF file (oge_gpu_dedup_reset first)");
    if (offsets[0] != 0 || offsets[nrec] != nbytes) return fail_msg(OGE_ERR_BAD_RECORD, "push: offsets[0] must be 0 and offsets[nrec] == nbytes");
    if (c->n + nrec >= (1ull << 30)) return fail_msg(OGE_ERR_TOO_LARGE, "push: more than 2^30-1 records in one context");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    int rc;
    // growing moves the resident data: drain the copy stream first
    if (c->rec_bytes + nbytes > c->rec.cap || c->n + nrec + 1 > c->off.cap) OGE_CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    if ((rc = c->rec.reserve(c->rec_bytes + nbytes, true, c->copy_stream))) return rc;
    if ((rc = c->off.reserve(c->n + nrec + 1, true, c->copy_stream))) return rc;
    OGE_CUDA_TRY(cudaMemcpyAsync(c->rec.p + c->rec_bytes, records, nbytes, cudaMemcpyHostToDevice, c->copy_stream));
    OGE_CUDA_TRY(cudaMemcpyAsync(c->off.p + c->n, offsets, (nrec + 1) * 8, cudaMemcpyHostToDevice, c->copy_stream));
    if (c->rec_bytes) {
        uint64_t cnt = nrec + 1;
        offsets_rebase<<<(uint32_t) ((cnt + 255) / 256), 256, 0, c->copy_stream>>>(c->off.p + c->n, cnt, c->rec_bytes);
        OGE_CUDA_TRY(cudaGetLastError());
    }
    c->rec_bytes += nbytes;
    c->n += nrec;
    c->ran = false;
    c->sorted = false;
    return OGE_OK;
}

int oge_gpu_dedup_sync(oge_gpu_dedup_ctx *c) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "sync: null context");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OGE_OK;
}

int oge_gpu_dedup_reset(oge_gpu_dedup_ctx *c) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "reset: null context");
    int rc = oge_gpu_dedup_sync(c);
    if (rc) return rc;
    c->n = 0;
    c->rec_bytes = 0;
    c->rec_lead = 0;
    c->ran = false;
    c->sorted = false;
    return OGE_OK;
}

}  // extern "C"

namespace oge {

// K1 end-build + K2 mate join over the context's records: the windowed join inside the CTAs, the global join over what
// they could not settle, the check pass, and -- replay_locally -- the exact path's replay.  The range-sharded path stops
// before the replay: a name that is not a plain couple on this rank is published instead (shard_api.cu), with its local
// pairs retracted by the fix-up as here.  Afterwards: pair lists and counters final for this rank, the mate table
// (c->table, out->n_slots) holding the leftover names, the exact-path list in c->sortbuf (out->n_cplx).
int join_stage(oge_gpu_dedup_ctx *c, bool replay_locally, JoinStage *out, uint64_t *launches_io) {
    cudaStream_t s = c->stream;
    const uint64_t n = c->n;
    EndbuildParams eb;
    eb.rec = c->recs(); eb.off = c->off.p; eb.n = n; eb.idx_base = c->cfg.index_base;
    eb.frag = c->frag.p; eb.hk = c->hk.p; eb.tag = c->tag.p; eb.flag_in = c->flag_in.p;
    eb.counters = c->counters.p; eb.rg = rg_table(c); eb.kl = c->kl;
    if (c->sh.on && c->cfg.world > 1 && c->sh.k1_route_cap) {      // range sharding: boundary fragment ends are listed by K1 itself
        eb.route_out = c->sh.route.p; eb.route_cap = c->sh.k1_route_cap;
        eb.own_lo = c->sh.own_lo; eb.own_hi = c->sh.own_hi;
    }
    const bool fused = !c->cfg.debug_legacy_join;      // windowed join (default) or the whole-file hash join
    uint64_t n_pairs = 0, n_far = 0, n_cplx = 0, n_retracted = 0, n_far_retracted = 0, n_left = 0, n_pe = 0;
    uint64_t &launches = *launches_io;
    int rc;
    JoinParams jp;
    jp.rec = c->recs(); jp.off = c->off.p; jp.n = n; jp.idx_base = c->cfg.index_base;
    jp.frag = c->frag.p; jp.hk = c->hk.p; jp.tag = c->tag.p;
    jp.mate_of = c->mate_of.p; jp.cplx = c->sortbuf.p;
    jp.counters = c->counters.p; jp.rg = rg_table(c); jp.kl = c->kl; jp.verify_names = c->cfg.verify_names;
    jp.list = nullptr; jp.n_list = 0;
    c->k_used = 0;
    c->k_begin(OGE_K_ENDBUILD, s);
    if ((rc = launch_endbuild(eb, (uint32_t) (c->rec_bytes / n), c->sms, s, &launches))) return rc;
    c->k_end(s);
    OGE_CUDA_TRY(cudaEventRecord(c->ev[1], s));
    if (fused) {
        // windowed join: every CTA pairs the reads of its contiguous record range in shared memory (no sizes needed from K1:
        // no host round trip in between); what it cannot settle is listed for the global join below
        const uint64_t pair_cap = n / 2 + 1024, far_cap = n / 2 + 1024;
        if ((rc = c->pair.reserve(pair_cap, false, s)) || (rc = c->pair2.reserve(pair_cap, false, s)) ||
            (rc = c->pairf.reserve(far_cap, false, s)) || (rc = c->pairf2.reserve(far_cap, false, s)) ||
            (rc = c->pair_hk.reserve(pair_cap, false, s)) || (rc = c->pairf_hk.reserve(far_cap, false, s)) ||
            (rc = c->left.reserve(n, false, s)))
            return rc;
        uint32_t lj_grid = 0, lj_tpc = 0;
        local_join_shape(n, c->sms, &lj_grid, &lj_tpc);
        if ((rc = c->couples.reserve((uint64_t) lj_grid * lj_tpc * (LJ_TILE / 2), false, s)) || (rc = c->couple_count.reserve(lj_grid, false, s))) return rc;
        LocalJoinParams lj;
        lj.couples = c->couples.p; lj.couple_count = c->couple_count.p; lj.couples_per_cta = 0;
        lj.pair = c->pair.p; lj.pair_far = c->pairf.p; lj.pair_hk = c->pair_hk.p; lj.pair_far_hk = c->pairf_hk.p;
        lj.pair_cap = (uint32_t) pair_cap; lj.far_cap = (uint32_t) far_cap;
        lj.mate_of = c->mate_of.p; lj.left = c->left.p; lj.n_buckets = 0; lj.tiles_per_cta = 0;
        jp.table = nullptr; jp.n_slots = 0; jp.pair = c->pair.p; jp.pair_far = c->pairf.p; jp.cplx_slots = nullptr;
        if ((rc = launch_local_join(jp, lj, c->sms, s, &launches, c))) return rc;
    }
    OGE_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, s));
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    if ((rc = check_endbuild_errors(c))) return rc;
    const uint64_t n_frag = c->h_counters[CNT_FRAG];
    n_pe = c->h_counters[CNT_PAIR_ELIGIBLE];
    c->h_counters_k1_unpaired = c->h_counters[CNT_UNPAIRED];
    n_left = c->h_counters[CNT_LEFT];

    // ---- K2 mate join: every map-eligible record (legacy form), or what the CTAs could not settle (fused form)
    const uint64_t n_join = fused ? n_left : n_pe;
    const uint32_t n_loc = fused ? c->h_counters[CNT_PAIRS] : 0u, n_loc_far = fused ? c->h_counters[CNT_PAIRS_FAR] : 0u;
    if (n_join) {
        // open addressing with linear probing: at most half full whatever the input (all-singleton names included)
        const uint64_t n_slots = 2 * n_join + 1024;
        if ((rc = c->table.reserve(n_slots, false, s))) return rc;
        if (!fused) {
            if ((rc = c->pair.reserve(n_pe / 2 + 1024, false, s))) return rc;
            if ((rc = c->pair2.reserve(n_pe / 2 + 1024, false, s))) return rc;
            if ((rc = c->pairf.reserve(n_pe / 2 + 1024, false, s))) return rc;      // worst case: every pair is a far pair
            if ((rc = c->pairf2.reserve(n_pe / 2 + 1024, false, s))) return rc;
        } else {
            // pairs the global join adds behind the CTAs' reservations
            const uint64_t need = (uint64_t) n_loc + n_join / 2 + 1, need_far = (uint64_t) n_loc_far + n_join / 2 + 1;
            if ((rc = c->pair.reserve(need, true, s)) || (rc = c->pair2.reserve(need, false, s)) ||
                (rc = c->pairf.reserve(need_far, true, s)) || (rc = c->pairf2.reserve(need_far, false, s)))
                return rc;
        }
        if ((rc = c->cplx_slots.reserve(n_join + 1024, false, s))) return rc;
        OGE_CUDA_TRY(cudaMemsetAsync(c->table.p, 0, n_slots * sizeof(MateSlot), s));
        jp.table = c->table.p; jp.n_slots = n_slots;
        jp.pair = c->pair.p; jp.pair_far = c->pairf.p; jp.cplx_slots = c->cplx_slots.p;
        if (fused) { jp.list = c->left.p; jp.n_list = (uint32_t) n_left; }
        c->k_begin(OGE_K_GLOBAL_JOIN, s);
        if ((rc = launch_mate_join(jp, s, &launches))) return rc;
        c->k_end(s);
        if (fused) {
            c->k_begin(OGE_K_CHECK, s);
            if ((rc = launch_pair_check(jp, c->pair_hk.p, n_loc, false, s, &launches))) return rc;
            if ((rc = launch_pair_check(jp, c->pairf_hk.p, n_loc_far, true, s, &launches))) return rc;
            c->k_end(s);
        }
        OGE_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, s));
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
        if (c->h_counters[CNT_COMPLEX]) {
            // names seen other than twice, or hash-equal couples with different names: the exact path
            if ((rc = launch_mate_fixup(jp, c->h_counters[CNT_COMPLEX_SLOTS], s, &launches))) return rc;
            OGE_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, s));
            OGE_CUDA_TRY(cudaStreamSynchronize(s));
            n_cplx = c->h_counters[CNT_COMPLEX];
            const uint64_t need = c->h_counters[CNT_PAIRS] + n_cplx / 2 + 1, need_far = c->h_counters[CNT_PAIRS_FAR] + n_cplx / 2 + 1;
            if ((rc = c->pair.reserve(need, true, s))) return rc;
            if ((rc = c->pair2.reserve(need, false, s))) return rc;
            if ((rc = c->pairf.reserve(need_far, true, s))) return rc;
            if ((rc = c->pairf2.reserve(need_far, false, s))) return rc;
            jp.pair = c->pair.p;
            jp.pair_far = c->pairf.p;
            if (replay_locally) {
            // sort (hash, ordinal), replay the toggle map per hash value
            E128 *sorted = nullptr;
            if ((rc = c->cplx_sort.reserve(n_cplx, false, s))) return rc;
            if ((rc = radix_sort_128(c->sortbuf.p, c->cplx_sort.p, n_cplx, nullptr, 0, 96, c->scratch.p, s, &sorted, &launches)))
                return rc;
            if ((rc = c->cplx_state.reserve(n_cplx + mate_complex_long_segs_bytes((uint32_t) n_cplx) + 16, false, s))) return rc;
            void *long_segs = c->cplx_state.p + ((n_cplx + 15) & ~15ull);
            if ((rc = launch_mate_complex(jp, sorted, (uint32_t) n_cplx, c->cplx_state.p, long_segs, s, &launches))) return rc;
            OGE_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, s));
            OGE_CUDA_TRY(cudaStreamSynchronize(s));
            }
        }
    }
    n_pairs = c->h_counters[CNT_PAIRS];
    n_far = c->h_counters[CNT_PAIRS_FAR];
    n_retracted = c->h_counters[CNT_PAIRS_RETRACTED];      // fused form: includes the reserved positions the CTAs did not use
    n_far_retracted = c->h_counters[CNT_FAR_RETRACTED];
    out->fused = fused;
    out->n_frag = n_frag; out->n_pe = n_pe; out->n_left = n_left; out->n_loc = n_loc; out->n_loc_far = n_loc_far;
    out->n_pairs = n_pairs; out->n_far = n_far; out->n_retracted = n_retracted; out->n_far_retracted = n_far_retracted; out->n_cplx = n_cplx;
    out->n_slots = n_join ? 2 * n_join + 1024 : 0;
    out->n_unpaired = c->h_counters_k1_unpaired;
    return 0;
}

}  // namespace oge

extern "C" {

int oge_gpu_dedup_run(oge_gpu_dedup_ctx *c) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "run: null context");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream;
    {   // the input-side figures (push_bgzf / frame) describe the resident records, not a run: they survive
        const oge_gpu_dedup_stats keep = c->stats;
        memset(&c->stats, 0, sizeof(c->stats));
        c->stats.ms_inflate = keep.ms_inflate;
        c->stats.inflate_blocks = keep.inflate_blocks;
        c->stats.inflate_bytes_in = keep.inflate_bytes_in;
        c->stats.inflate_bytes_out = keep.inflate_bytes_out;
        c->stats.ms_frame = keep.ms_frame;
        c->stats.frame_repairs = keep.frame_repairs;
        c->stats.ms_inflate_h2d = keep.ms_inflate_h2d;
        c->stats.ms_inflate_d2h = keep.ms_inflate_d2h;
        c->stats.ms_push_bgzf = keep.ms_push_bgzf;
        c->stats.ms_inflate_start = keep.ms_inflate_start;
        c->zfile_bytes = 0;      // a new run: the members made from the last one are stale
        c->stats.inflate_mode = keep.inflate_mode;
        c->stats.inflate_pieces = keep.inflate_pieces;
    }
    c->stats.n_records = c->n;
    if (c->n == 0) {
        c->ran = true;
        return OGE_OK;
    }
    int rc;
    if ((rc = compute_layout(c, &c->kl))) return rc;
    if ((rc = ensure_work(c))) return rc;
    const uint64_t n = c->n;
    uint64_t launches = 0;
    PassTimer timer{c->pass_ev, 48, 0, 0};
    PassTimer *tp = c->cfg.profile_events ? &timer : nullptr;

    // the pushes ran on the copy stream
    OGE_CUDA_TRY(cudaEventRecord(c->copy_done, c->copy_stream));
    OGE_CUDA_TRY(cudaStreamWaitEvent(s, c->copy_done, 0));

    OGE_CUDA_TRY(cudaEventRecord(c->ev[0], s));
    OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, CNT_N * 4, s));
    OGE_CUDA_TRY(cudaMemsetAsync(c->dup.p, 0, n, s));

    // ---- K1 end-build and K2 mate join (shared with the range-sharded path: join_stage)
    JoinStage js;
    if ((rc = join_stage(c, true, &js, &launches))) return rc;
    const uint64_t n_frag = js.n_frag, n_pairs = js.n_pairs, n_far = js.n_far, n_cplx = js.n_cplx, n_retracted = js.n_retracted,
                   n_far_retracted = js.n_far_retracted, n_left = js.n_left;
    const bool fused = js.fused;
    const uint32_t n_loc = js.n_loc, n_loc_far = js.n_loc_far;
    const uint64_t n_dead_k1 = 0;
    OGE_CUDA_TRY(cudaEventRecord(c->ev[2], s));

    // ---- K3 + K4 on the pairs
    SelectParams sp;
    sp.dup = c->dup.p; sp.mate_of = c->mate_of.p; sp.idx_base = c->cfg.index_base; sp.n_records = n;
    sp.counters = c->counters.p; sp.kl = c->kl; sp.n_dev = nullptr;
    sp.fm = nullptr; sp.n_fm = 0; sp.foreign_marks = nullptr; sp.foreign_cap = 0; sp.foreign_counter = c->counters.p + CNT_FOREIGN_MARKS;
    sp.split = nullptr; sp.world = 1; sp.rank = 0;
    // near pairs (short key) and far pairs (full key) are separate lists: equal keys imply equal class
    E128 *sorted_pairs = c->pair.p, *sorted_far = c->pairf.p;
    if (n_pairs) {
        if ((rc = radix_sort_128(c->pair.p, c->pair2.p, n_pairs, nullptr, c->kl.n_delta, c->kl.p_end, c->scratch.p, s, &sorted_pairs,
                                 &launches, tp)))
            return rc;
    }
    if (n_far) {
        if ((rc = radix_sort_128(c->pairf.p, c->pairf2.p, n_far, nullptr, c->kl.p_coord2, c->kl.p_end, c->scratch.p, s, &sorted_far,
                                 &launches, tp)))
            return rc;
    }
    OGE_CUDA_TRY(cudaEventRecord(c->ev[3], s));
    if (n_pairs > n_retracted) {      // retracted provisional pairs are all-ones entries: they sorted to the tail
        sp.sorted = sorted_pairs; sp.n_max = (uint32_t) (n_pairs - n_retracted);
        c->k_begin(OGE_K_SELECT, s);
        if ((rc = launch_select_pairs(sp, false, s, &launches))) return rc;
        c->k_end(s);
    }
    if (n_far > n_far_retracted) {
        sp.sorted = sorted_far; sp.n_max = (uint32_t) (n_far - n_far_retracted);
        c->k_begin(OGE_K_SELECT, s);
        if ((rc = launch_select_pairs(sp, true, s, &launches))) return rc;
        c->k_end(s);
    }
    OGE_CUDA_TRY(cudaEventRecord(c->ev[4], s));

    // ---- K3 + K4 on the fragments.  Only an end that is not part of a pair can be marked here, and an end
    //      of a pair only matters when it shares its key with such an end (fragfilter.cu): sort that subset
    //      when unpaired ends are rare (none at all on clean paired-end data), everything otherwise
    //      (ineligible records carry the all-ones key and sort last).
    const uint64_t n_unp = c->h_counters_k1_unpaired;
    E128 *sorted_frags = c->frag.p;
    uint64_t n_fsel = 0;
    int frag_mode = 2;      // 0 nothing to do, 1 reduced, 2 full
    if (n_frag == 0) frag_mode = 0;
    else if (c->cfg.debug_full_frag_sort) frag_mode = 2;
    else if (n_unp == 0) frag_mode = 0;
    else if (n_unp <= n_frag / 16 && c->kl.f_end - c->kl.f_orient <= 63) frag_mode = 1;
    if (frag_mode == 1) {
        uint64_t ucap = std::max<uint64_t>(n_frag / 4, 4 * n_unp) + 1024;
#ifdef OGE_TESTING
        if (const char *e = getenv("OGE_UFRAG_CAP")) ucap = std::max<uint64_t>(1, (uint64_t) atoll(e));      // test hook: force the fallback
#endif
        uint64_t n_slots = 1024;
        while (n_slots < 4 * n_unp) n_slots <<= 1;
        if ((rc = c->ufrag.reserve(ucap, false, s))) return rc;
        if ((rc = c->ufrag2.reserve(ucap, false, s))) return rc;
        if ((rc = c->uset.reserve(n_slots, false, s))) return rc;
        OGE_CUDA_TRY(cudaMemsetAsync(c->uset.p, 0, n_slots * 8, s));
        OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p + CNT_UFRAG, 0, 4, s));
        if ((rc = launch_ff_collect(c->frag.p, n, c->kl, c->ufrag.p, (uint32_t) ucap, c->counters.p, s, &launches))) return rc;
        if ((rc = launch_ff_set_build(c->ufrag.p, c->counters.p + CNT_UFRAG, (uint32_t) std::min(n_unp, ucap), c->kl, c->uset.p, n_slots, s, &launches))) return rc;
        if ((rc = launch_ff_filter(c->frag.p, n, c->kl, c->uset.p, n_slots, c->ufrag.p, (uint32_t) ucap, c->counters.p, nullptr, s, &launches))) return rc;
        OGE_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, s));
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
        n_fsel = c->h_counters[CNT_UFRAG];
        if (n_fsel > ucap) frag_mode = 2;      // more paired ends share keys with unpaired ones than expected: sort everything
        else if ((rc = radix_sort_128(c->ufrag.p, c->ufrag2.p, n_fsel, nullptr, c->kl.f_orient, c->kl.f_end, c->scratch.p, s, &sorted_frags,
                                      &launches, tp)))
            return rc;
    }
    if (frag_mode == 2) {
        n_fsel = n_frag;
        if ((rc = radix_sort_128(c->frag.p, c->sortbuf.p, n, nullptr, c->kl.f_orient, c->kl.f_end, c->scratch.p, s, &sorted_frags,
                                 &launches, tp)))
            return rc;
    }
    OGE_CUDA_TRY(cudaEventRecord(c->ev[5], s));
    if (frag_mode && n_fsel) {
        sp.sorted = sorted_frags; sp.n_max = (uint32_t) n_fsel;
        c->k_begin(OGE_K_SELECT, s);
        if ((rc = launch_select_frags(sp, s, &launches))) return rc;
        c->k_end(s);
    }
    OGE_CUDA_TRY(cudaEventRecord(c->ev[6], s));

    // ---- K5 flag write
    FlagParams fp;
    fp.rec = c->recs(); fp.off = c->off.p; fp.n = n; fp.flag_in = c->flag_in.p; fp.flag_out = c->flag_out.p;
    fp.dup = c->dup.p; fp.counters = c->counters.p; fp.quiet_index_bug = c->cfg.compat_quiet_index_bug;
    c->k_begin(OGE_K_FLAGS, s);
    if ((rc = launch_flags(fp, s, &launches))) return rc;
    c->k_end(s);
    OGE_CUDA_TRY(cudaEventRecord(c->ev[7], s));
    OGE_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, s));
    OGE_CUDA_TRY(cudaStreamSynchronize(s));
    // the fragment sort moved the end entries: keep them findable for debug_ends
    if (c->cfg.debug_keep_ends && frag_mode == 2 && sorted_frags != c->frag.p && n_frag)
        OGE_CUDA_TRY(cudaMemcpy(c->frag.p, sorted_frags, n * sizeof(E128), cudaMemcpyDeviceToDevice));

#ifdef OGE_TESTING
    if (getenv("OGE_DEBUG_COUNTERS"))
        fprintf(stderr, "[oge] n=%llu frag=%llu pe=%llu near=%llu far=%llu cplx=%llu retracted=%llu/%llu\n", (unsigned long long) n,
                (unsigned long long) n_frag, (unsigned long long) js.n_pe, (unsigned long long) n_pairs, (unsigned long long) n_far,
                (unsigned long long) n_cplx, (unsigned long long) n_retracted, (unsigned long long) n_far_retracted);
#endif
    oge_gpu_dedup_stats &st = c->stats;
    st.n_frag_entries = n_frag;
    st.n_pair_entries = n_pairs - n_retracted + n_far - n_far_retracted;
    st.n_duplicates = c->h_counters[CNT_DUPS];
    st.n_complex_names = n_cplx;
    st.n_local_pairs = fused ? (uint64_t) n_loc - n_dead_k1 + n_loc_far : 0;
    st.n_join_leftovers = n_left;
    st.n_local_retracted = fused ? n_retracted + n_far_retracted - n_dead_k1 : 0;
    st.n_hash_mismatch = c->h_counters[CNT_HASH_MISMATCH];
    st.frag_key_bits = c->kl.f_end - c->kl.f_orient;
    st.pair_key_bits = c->kl.p_end - c->kl.n_delta;      // near pairs (far pairs: p_end - p_coord2)
    st.frag_sort_passes = make_sort_plan(c->kl.f_orient, c->kl.f_end).n_pass;
    st.pair_sort_passes = make_sort_plan(c->kl.n_delta, c->kl.p_end).n_pass;
    st.ms_total = ms_between(c->ev[0], c->ev[7]);
    st.ms_endbuild = ms_between(c->ev[0], c->ev[1]);
    st.ms_join = ms_between(c->ev[1], c->ev[2]);
    st.ms_sort_pair = ms_between(c->ev[2], c->ev[3]);
    st.ms_sort_frag = ms_between(c->ev[4], c->ev[5]);
    st.ms_select = ms_between(c->ev[3], c->ev[4]) + ms_between(c->ev[5], c->ev[6]);
    st.ms_flags = ms_between(c->ev[6], c->ev[7]);
    st.launches = launches;
    for (int i = 0; i < timer.used; i++) st.ms_sort_pass_kernels += ms_between(c->pass_ev[2 * i], c->pass_ev[2 * i + 1]);
    st.sort_pass_launches = timer.used;
    for (int i = 0; i < c->k_used; i++) st.ms_kernel[c->k_slot[i]] += ms_between(c->k_ev[2 * i], c->k_ev[2 * i + 1]);
    st.sort_pass_bytes = timer.bytes;
    c->ran = true;
    return OGE_OK;
}

int oge_gpu_dedup_flags(oge_gpu_dedup_ctx *c, uint16_t *out, uint64_t n) {
    if (!c || (n && !out)) return fail_msg(OGE_ERR_INVALID_ARG, "flags: null argument");
    if (!c->ran) return fail_msg(OGE_ERR_STATE, "flags: call oge_gpu_dedup_run first");
    if (n != c->n) return fail_msg(OGE_ERR_INVALID_ARG, "flags: n=%llu but the context holds %llu records", (unsigned long long) n, (unsigned long long) c->n);
    if (n == 0) return OGE_OK;
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    OGE_CUDA_TRY(cudaMemcpyAsync(out, c->flag_out.p, n * 2, cudaMemcpyDeviceToHost, c->stream));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OGE_OK;
}

int oge_gpu_dedup_pull(oge_gpu_dedup_ctx *c, uint8_t *out_records, uint64_t cap_bytes, uint64_t *out_offsets, uint64_t cap_records,
                       uint64_t *out_bytes, uint64_t *out_nrec) {
    if (!c || !out_bytes || !out_nrec) return fail_msg(OGE_ERR_INVALID_ARG, "pull: null argument");
    // after a run: the flag-patched records; after a sort alone (the ReadSorter drop-in): the records in their new order
    if (!c->ran && !(c->sorted && !c->cfg.remove_duplicates)) return fail_msg(OGE_ERR_STATE, "pull: call oge_gpu_dedup_run first");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream;
    *out_bytes = 0;
    *out_nrec = 0;
    if (c->n == 0) {
        if (out_offsets && cap_records >= 1) out_offsets[0] = 0;
        return OGE_OK;
    }
    if (!c->cfg.remove_duplicates) {
        if (cap_bytes < c->rec_bytes || !out_records) return fail_msg(OGE_ERR_INVALID_ARG, "pull: need %llu bytes", (unsigned long long) c->rec_bytes);
        if (out_offsets && cap_records < c->n + 1) return fail_msg(OGE_ERR_INVALID_ARG, "pull: need %llu offsets", (unsigned long long) c->n + 1);
        OGE_CUDA_TRY(cudaMemcpyAsync(out_records, c->recs(), c->rec_bytes, cudaMemcpyDeviceToHost, s));
        if (out_offsets) OGE_CUDA_TRY(cudaMemcpyAsync(out_offsets, c->off.p, (c->n + 1) * 8, cudaMemcpyDeviceToHost, s));
        OGE_CUDA_TRY(cudaStreamSynchronize(s));
        *out_bytes = c->rec_bytes;
        *out_nrec = c->n;
        return OGE_OK;
    }
    // -r: drop what is flagged after the rewrite (:456-458), compact in order
    DevBuf<uint8_t> tmp_rec;
    DevBuf<uint64_t> tmp_off;
    int rc;
    uint64_t launches = 0;
    if ((rc = tmp_rec.reserve(c->rec_bytes, false, s))) return rc;
    if ((rc = tmp_off.reserve(c->n + 1, false, s))) { tmp_rec.release(); return rc; }
    rc = launch_compact(c->recs(), c->off.p, c->n, c->flag_out.p, 1, tmp_rec.p, tmp_off.p, reinterpret_cast<uint64_t *>(c->scratch.p),
                        c->counters.p, s, &launches);
    uint64_t totals[2] = {0, 0};
    if (!rc && cudaMemcpyAsync(totals, c->scratch.p, 16, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "pull totals", __FILE__, __LINE__);
    if (!rc && cudaStreamSynchronize(s) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "pull sync", __FILE__, __LINE__);
    if (!rc && (totals[1] > cap_bytes || (totals[1] && !out_records))) rc = fail_msg(OGE_ERR_INVALID_ARG, "pull: need %llu bytes", (unsigned long long) totals[1]);
    if (!rc && out_offsets && cap_records < totals[0] + 1) rc = fail_msg(OGE_ERR_INVALID_ARG, "pull: need %llu offsets", (unsigned long long) totals[0] + 1);
    if (!rc && totals[1] && cudaMemcpyAsync(out_records, tmp_rec.p, totals[1], cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "pull copy", __FILE__, __LINE__);
    if (!rc && out_offsets && cudaMemcpyAsync(out_offsets, tmp_off.p, (totals[0] + 1) * 8, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "pull copy", __FILE__, __LINE__);
    if (!rc && cudaStreamSynchronize(s) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "pull sync", __FILE__, __LINE__);
    tmp_rec.release();
    tmp_off.release();
    if (rc) return rc;
    *out_bytes = totals[1];
    *out_nrec = totals[0];
    return OGE_OK;
}

int oge_gpu_dedup_deflate(oge_gpu_dedup_ctx *c, uint64_t *out_bytes, uint64_t *out_blocks, uint64_t *out_nrec) {
    if (!c || !out_bytes || !out_blocks || !out_nrec) return fail_msg(OGE_ERR_INVALID_ARG, "deflate: null argument");
    if (!c->ran) return fail_msg(OGE_ERR_STATE, "deflate: call oge_gpu_dedup_run first");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->stream;
    *out_bytes = 0;
    *out_blocks = 0;
    *out_nrec = 0;
    c->zfile_bytes = 0;
    c->stats.ms_deflate = 0;
    c->stats.deflate_blocks = c->stats.deflate_bytes_in = c->stats.deflate_bytes_out = 0;
    if (c->n == 0) return OGE_OK;
    int rc;
    uint64_t launches = 0;
    cudaEvent_t e0 = c->ev[8], e1 = c->ev[9];
    OGE_CUDA_TRY(cudaEventRecord(e0, s));
    // the writer's bin for every record (bam_serializer.h:112-116), in place: idempotent, and what pull returns afterwards
    // carries it as well
    OGE_CUDA_TRY(cudaMemsetAsync(c->counters.p + CNT_ERR, 0, 4, s));
    if ((rc = launch_fix_bins(c->recs(), c->off.p, c->n, c->counters.p + CNT_ERR, s, &launches))) return rc;
    const uint8_t *in = c->recs();
    uint64_t total = c->rec_bytes, nrec = c->n;
    DevBuf<uint8_t> kept;         // -r: the records that stay, compacted
    DevBuf<uint64_t> kept_off;
    DevBuf<uint8_t> stage;
    DevBuf<uint32_t> sizes;       // dsize[n_blocks], crc[n_blocks], ticket (2 words)
    DevBuf<uint64_t> moff;
    DevBuf<uint8_t> seqs;
    auto done = [&](int code) {
        kept.release(); kept_off.release(); stage.release(); sizes.release(); moff.release(); seqs.release();
        return code;
    };
    // CUDA failures from here on release the temporaries as well
#define OGE_DEFLATE_TRY(expr)                                                            \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) return done(fail_cuda(_e, #expr, __FILE__, __LINE__));    \
    } while (0)
    uint32_t bin_err = 0;
    if (c->cfg.remove_duplicates) {      // :456-458
        if ((rc = kept.reserve(c->rec_bytes, false, s)) || (rc = kept_off.reserve(c->n + 1, false, s))) return done(rc);
        if ((rc = launch_compact(c->recs(), c->off.p, c->n, c->flag_out.p, 1, kept.p, kept_off.p, reinterpret_cast<uint64_t *>(c->scratch.p),
                                 c->counters.p, s, &launches)))
            return done(rc);
        uint64_t totals[2] = {0, 0};
        OGE_DEFLATE_TRY(cudaMemcpyAsync(totals, c->scratch.p, 16, cudaMemcpyDeviceToHost, s));
        OGE_DEFLATE_TRY(cudaMemcpyAsync(&bin_err, c->counters.p + CNT_ERR, 4, cudaMemcpyDeviceToHost, s));
        OGE_DEFLATE_TRY(cudaStreamSynchronize(s));
        nrec = totals[0];
        total = totals[1];
        in = kept.p;
    }
    const uint64_t n_blocks = (total + DEFLATE_PAYLOAD - 1) / DEFLATE_PAYLOAD;
    if (n_blocks) {
        if ((rc = stage.reserve(n_blocks * (uint64_t) DEFLATE_STAGE_STRIDE, false, s)) || (rc = sizes.reserve(2 * n_blocks + 4, false, s)) ||
            (rc = moff.reserve(n_blocks + 1, false, s)) || (rc = seqs.reserve(deflate_seq_bytes(c->sms), false, s)))
            return done(rc);
        DeflateParams P;
        P.in = in;
        P.total = total;
        P.payload = DEFLATE_PAYLOAD;
        P.n_blocks = n_blocks;
        P.stage = stage.p;
        P.dsize = sizes.p;
        P.crc = sizes.p + n_blocks;
        P.seqs = seqs.p;
        P.ticket = reinterpret_cast<unsigned long long *>(sizes.p + ((2 * n_blocks + 1) & ~1ull));
        if ((rc = launch_bgzf_deflate(P, c->sms, s, &launches)) || (rc = launch_scan_sizes(P.dsize, n_blocks, moff.p, s, &launches))) return done(rc);
        uint64_t file_bytes = 0;
        OGE_DEFLATE_TRY(cudaMemcpyAsync(&file_bytes, moff.p + n_blocks, 8, cudaMemcpyDeviceToHost, s));
        if (!c->cfg.remove_duplicates) OGE_DEFLATE_TRY(cudaMemcpyAsync(&bin_err, c->counters.p + CNT_ERR, 4, cudaMemcpyDeviceToHost, s));
        OGE_DEFLATE_TRY(cudaStreamSynchronize(s));
        if ((rc = c->zfile.reserve(file_bytes, false, s))) return done(rc);
        if ((rc = launch_bgzf_assemble(P, moff.p, c->zfile.p, s, &launches))) return done(rc);
        c->zfile_bytes = file_bytes;
    }
    OGE_DEFLATE_TRY(cudaEventRecord(e1, s));
    OGE_DEFLATE_TRY(cudaStreamSynchronize(s));
    if (bin_err) return done(fail_msg(OGE_ERR_BAD_RECORD, "deflate: a record's name and CIGAR overrun the record"));
    c->stats.ms_deflate = ms_between(e0, e1);
    c->stats.deflate_blocks = n_blocks;
    c->stats.deflate_bytes_in = total;
    c->stats.deflate_bytes_out = c->zfile_bytes;
    *out_bytes = c->zfile_bytes;
    *out_blocks = n_blocks;
    *out_nrec = nrec;
    return done(OGE_OK);
#undef OGE_DEFLATE_TRY
}

int oge_gpu_dedup_pull_bgzf(oge_gpu_dedup_ctx *c, uint8_t *out, uint64_t cap_bytes) {
    if (!c || (!out && c && c->zfile_bytes)) return fail_msg(OGE_ERR_INVALID_ARG, "pull_bgzf: null argument");
    if (cap_bytes < c->zfile_bytes) return fail_msg(OGE_ERR_INVALID_ARG, "pull_bgzf: need %llu bytes", (unsigned long long) c->zfile_bytes);
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    if (c->zfile_bytes) {
        OGE_CUDA_TRY(cudaMemcpyAsync(out, c->zfile.p, c->zfile_bytes, cudaMemcpyDeviceToHost, c->stream));
        OGE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return OGE_OK;
}

int oge_gpu_dedup_pull_bgzf_part(oge_gpu_dedup_ctx *c, uint64_t offset, uint64_t nbytes, uint8_t *out) {
    if (!c || (!out && nbytes)) return fail_msg(OGE_ERR_INVALID_ARG, "pull_bgzf_part: null argument");
    if (offset > c->zfile_bytes || nbytes > c->zfile_bytes - offset)
        return fail_msg(OGE_ERR_INVALID_ARG, "pull_bgzf_part: [%llu, +%llu) is outside the %llu bytes of members", (unsigned long long) offset,
                        (unsigned long long) nbytes, (unsigned long long) c->zfile_bytes);
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    if (nbytes) OGE_CUDA_TRY(cudaMemcpyAsync(out, c->zfile.p + offset, nbytes, cudaMemcpyDeviceToHost, c->side_stream));
    return OGE_OK;
}

int oge_gpu_dedup_pull_bgzf_wait(oge_gpu_dedup_ctx *c) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "pull_bgzf_wait: null context");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->side_stream));
    return OGE_OK;
}

int oge_gpu_dedup_flagstats(oge_gpu_dedup_ctx *c, oge_gpu_flagstats *out) {
    if (!c || !out) return fail_msg(OGE_ERR_INVALID_ARG, "flagstats: null argument");
    if (!c->ran) return fail_msg(OGE_ERR_STATE, "flagstats: call oge_gpu_dedup_run first");
    static_assert(sizeof(oge_gpu_flagstats) == FS_N_OUT * sizeof(uint64_t), "oge_gpu_flagstats mirrors the FS_* order");
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    DevBuf<uint8_t> tmp;
    int rc = tmp.reserve(flagstat_scratch_bytes(c->n), false, c->stream);
    if (rc) return rc;
    uint64_t launches = 0;
    rc = launch_flagstats(c->recs(), c->off.p, c->flag_out.p, c->n, tmp.p, c->sms, c->stream, &launches);
    if (!rc && cudaMemcpyAsync(out, tmp.p, sizeof(*out), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        rc = fail_cuda(cudaGetLastError(), "flagstats copy", __FILE__, __LINE__);
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "flagstats sync", __FILE__, __LINE__);
    tmp.release();
    return rc;
}

int oge_gpu_dedup_get_stats(oge_gpu_dedup_ctx *c, oge_gpu_dedup_stats *out) {
    if (!c || !out) return fail_msg(OGE_ERR_INVALID_ARG, "get_stats: null argument");
    *out = c->stats;
    return OGE_OK;
}

int oge_gpu_dedup_debug_ends(oge_gpu_dedup_ctx *c, oge_gpu_end *out, uint64_t n) {
    if (!c || (n && !out)) return fail_msg(OGE_ERR_INVALID_ARG, "debug_ends: null argument");
    if (!c->ran || !c->cfg.debug_keep_ends) return fail_msg(OGE_ERR_STATE, "debug_ends: needs debug_keep_ends and a completed run");
    if (n != c->n) return fail_msg(OGE_ERR_INVALID_ARG, "debug_ends: n mismatch");
    if (n == 0) return OGE_OK;
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    std::vector<E128> ents(n);
    std::vector<uint64_t> hk(n);
    OGE_CUDA_TRY(cudaMemcpy(ents.data(), c->frag.p, n * sizeof(E128), cudaMemcpyDeviceToHost));
    OGE_CUDA_TRY(cudaMemcpy(hk.data(), c->hk.p, n * 8, cudaMemcpyDeviceToHost));
    const KeyLayout &L = c->kl;
    memset(out, 0, n * sizeof(oge_gpu_end));
    for (uint64_t j = 0; j < n; j++) {      // entries are in sorted order: place each by its ordinal
        const E128 &e = ents[j];
        if (bits_get(e, L.f_lib, L.lib_bits) == L.lib_invalid) continue;
        uint64_t i = bits_get(e, L.f_idx, L.idx_bits) - c->cfg.index_base;
        if (i >= n) return fail_msg(OGE_ERR_STATE, "debug_ends: corrupt entry");
        oge_gpu_end &o = out[i];
        o.eligible = 1;
        o.pair_eligible = hk[i] != 0;
        o.ref = (int32_t) bits_get(e, L.f_ref, L.ref_bits);
        o.coord = (int32_t) ((int64_t) bits_get(e, L.f_coord, L.coord_bits) - L.coord_bias);
        o.orientation = bits_get(e, L.f_orient, 1) ? 2 : 1;
        o.read2Sequence = bits_get(e, L.f_paired, 1) ? 0 : -1;
        o.score = (int16_t) (uint16_t) (e.lo & 0xFFFF);
        o.lib = (int16_t) bits_get(e, L.f_lib, L.lib_bits);
    }
    return OGE_OK;
}

int oge_gpu_debug_sort128(int device, void *entries, uint64_t n, int bit_lo, int bit_hi) {
    if ((n && !entries) || bit_lo < 0 || bit_hi > 128 || bit_lo > bit_hi) return fail_msg(OGE_ERR_INVALID_ARG, "debug_sort128: bad argument");
    if (oge_gpu_device_count() <= 0) return fail_msg(OGE_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (n == 0) return OGE_OK;
    OGE_CUDA_TRY(cudaSetDevice(device));
    int rc = radix_sort_init();
    if (rc) return rc;
    DevBuf<E128> a, b;
    DevBuf<uint8_t> scratch;
    uint64_t launches = 0;
    E128 *res = nullptr;
    rc = a.reserve(n, false, 0);
    if (!rc) rc = b.reserve(n, false, 0);
    if (!rc) rc = scratch.reserve(sort_scratch_bytes(n), false, 0);
    if (!rc && cudaMemcpy(a.p, entries, n * sizeof(E128), cudaMemcpyHostToDevice) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "sort h2d", __FILE__, __LINE__);
    if (!rc) rc = radix_sort_128(a.p, b.p, n, nullptr, bit_lo, bit_hi, scratch.p, 0, &res, &launches);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "sort sync", __FILE__, __LINE__);
    if (!rc && cudaMemcpy(entries, res, n * sizeof(E128), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "sort d2h", __FILE__, __LINE__);
    a.release(); b.release(); scratch.release();
    return rc;
}

int oge_gpu_copy_d2d(oge_gpu_dedup_ctx *c, void *dst, const void *src, uint64_t nbytes) {
    if (!c || (nbytes && (!dst || !src))) return fail_msg(OGE_ERR_INVALID_ARG, "copy_d2d: null argument");
    if (!nbytes) return OGE_OK;
    OGE_CUDA_TRY(cudaSetDevice(c->cfg.device));
    OGE_CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, c->stream));
    OGE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OGE_OK;
}

int oge_gpu_dedup_device_ptrs(oge_gpu_dedup_ctx *c, void **records, void **offsets, void **flags) {
    if (!c) return fail_msg(OGE_ERR_INVALID_ARG, "device_ptrs: null context");
    if (records) *records = c->recs();
    if (offsets) *offsets = c->off.p;
    if (flags) *flags = c->flag_out.p;
    return OGE_OK;
}

}  // extern "C"

// ---- K3 measurement hook: device-generated entries, CUDA-event timing, on-device verification ----
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// mode 0: uniform random bits; mode 1: "coordinate sorted": high key bits grow with i, low bits random
__global__ void sb_fill(E128 *e, uint64_t n, uint64_t seed, int mode, int bit_lo, int bit_hi) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    E128 v;
    v.lo = mix64(seed + 2 * i);
    v.hi = mix64(seed + 2 * i + 1);
    if (mode == 1) {
        int kb = bit_hi - bit_lo;
        int top = kb > 24 ? 24 : kb;      // top bits follow the ordinal
        uint64_t t = (uint64_t) ((double) i / (double) n * (double) (1ull << top));
        E128 m;      // clear then set bits [bit_hi - top, bit_hi)
        m.lo = m.hi = 0;
        bits_or(m, bit_hi - top, (1ull << top) - 1);
        v.lo &= ~m.lo; v.hi &= ~m.hi;
        bits_or(v, bit_hi - top, t);
    }
    e[i] = v;
}

// sum and xor of all words (order independent) + count of adjacent key inversions
__global__ void sb_check(const E128 *e, uint64_t n, int bit_lo, int bit_hi, unsigned long long *out /* [4] */) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long sum = 0, x = 0, inv = 0;
    if (i < n) {
        E128 a = e[i];
        sum = a.lo + 3 * a.hi;
        x = a.lo ^ (a.hi * 0x9E3779B97F4A7C15ull);
        if (i + 1 < n) {
            E128 b = e[i + 1];
            E128 ka = bits_from(a, bit_lo), kb = bits_from(b, bit_lo);
            int w = bit_hi - bit_lo;
            if (w < 128) {
                if (w <= 64) { uint64_t m = w == 64 ? ~0ull : (1ull << w) - 1; ka.lo &= m; kb.lo &= m; ka.hi = kb.hi = 0; }
                else { uint64_t m = (1ull << (w - 64)) - 1; ka.hi &= m; kb.hi &= m; }
            }
            inv = (ka.hi > kb.hi || (ka.hi == kb.hi && ka.lo > kb.lo)) ? 1 : 0;
        }
    }
    for (int o = 16; o; o >>= 1) {
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        x ^= __shfl_xor_sync(0xFFFFFFFFu, x, o);
        inv += __shfl_xor_sync(0xFFFFFFFFu, inv, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], sum);
        atomicXor(&out[1], x);
        atomicAdd(&out[2], inv);
    }
}

}  // namespace

extern "C" int oge_gpu_debug_sort_bench(int device, uint64_t n, int bit_lo, int bit_hi, int variant, int mode, int reps,
                                        uint64_t seed, float *ms_pass_avg, float *ms_sort_avg, int *n_pass, int *verified) {
    if (!ms_pass_avg || !ms_sort_avg || !n_pass || !verified || n == 0 || bit_lo < 0 || bit_hi > 128 || bit_lo >= bit_hi || reps < 1)
        return fail_msg(OGE_ERR_INVALID_ARG, "debug_sort_bench: bad argument");
    if (oge_gpu_device_count() <= 0) return fail_msg(OGE_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    OGE_CUDA_TRY(cudaSetDevice(device));
    int rc = radix_sort_init();
    if (rc) return rc;
    const int saved = radix_sort_get_variant();
    radix_sort_set_variant(variant & 0xFF);
    radix_sort_set_prefetch(variant >> 8);
    DevBuf<E128> a, b;
    DevBuf<uint8_t> scratch;
    DevBuf<unsigned long long> chk;
    cudaEvent_t ev[2 * 48 + 2];
    for (auto &e : ev) cudaEventCreate(&e);
    uint64_t launches = 0;
    float pass_ms = 0, sort_ms = 0;
    int passes = 0, ok = 1;
    do {
        if ((rc = a.reserve(n, false, 0)) || (rc = b.reserve(n, false, 0)) || (rc = scratch.reserve(sort_scratch_bytes(n), false, 0)) ||
            (rc = chk.reserve(8, false, 0)))
            break;
        const uint32_t grid = (uint32_t) ((n + 255) / 256);
        for (int r = 0; r < reps + 1 && !rc; r++) {      // first repetition is warm-up
            unsigned long long h0[4], h1[4];
            sb_fill<<<grid, 256>>>(a.p, n, seed + r, mode, bit_lo, bit_hi);
            cudaMemset(chk.p, 0, 64);
            sb_check<<<grid, 256>>>(a.p, n, bit_lo, bit_hi, chk.p);
            cudaMemcpy(h0, chk.p, 32, cudaMemcpyDeviceToHost);
            PassTimer timer{ev + 2, 48, 0, 0};
            E128 *res = nullptr;
            cudaEventRecord(ev[0], 0);
            rc = radix_sort_128(a.p, b.p, n, nullptr, bit_lo, bit_hi, scratch.p, 0, &res, &launches, &timer);
            cudaEventRecord(ev[1], 0);
            if (rc) break;
            if (cudaDeviceSynchronize() != cudaSuccess) { rc = fail_cuda(cudaGetLastError(), "sort bench sync", __FILE__, __LINE__); break; }
            cudaMemset(chk.p, 0, 64);
            sb_check<<<grid, 256>>>(res, n, bit_lo, bit_hi, chk.p);
            cudaMemcpy(h1, chk.p, 32, cudaMemcpyDeviceToHost);
            if (h0[0] != h1[0] || h0[1] != h1[1] || h1[2] != 0) ok = 0;
            if (r > 0) {
                sort_ms += ms_between(ev[0], ev[1]);
                for (int i = 0; i < timer.used; i++) pass_ms += ms_between(ev[2 + 2 * i], ev[2 + 2 * i + 1]);
                passes = timer.used;
            }
        }
    } while (0);
    for (auto &e : ev) cudaEventDestroy(e);
    a.release(); b.release(); scratch.release(); chk.release();
    radix_sort_set_variant(saved);
    if (rc) return rc;
    *ms_pass_avg = passes ? pass_ms / (reps * passes) : 0;
    *ms_sort_avg = sort_ms / reps;
    *n_pass = passes;
    *verified = ok;
    return OGE_OK;
}

extern "C" int oge_gpu_set_inflate_kernel(int kernel) {
    if (kernel < -1 || kernel > 2)
        return fail_msg(OGE_ERR_INVALID_ARG, "set_inflate_kernel: 0 (thread per block), 1 (warp per block), 2 (hardware decompress engine) or -1 (default)");
    oge::set_inflate_kernel(kernel);
    return OGE_OK;
}

extern "C" int oge_gpu_inflate_kernel(int device) {
    if (device < 0 || device >= oge_gpu_device_count()) return fail_msg(OGE_ERR_INVALID_ARG, "inflate_kernel: device %d", device);
    return oge::inflate_mode_for(device);
}

extern "C" int oge_gpu_set_bgzf_chunk_bytes(uint64_t bytes) {
    g_bgzf_chunk_bytes = bytes;
    return OGE_OK;
}

extern "C" int oge_gpu_set_bgzf_staging(uint64_t stage_bytes) {
    g_bgzf_stage_bytes = stage_bytes;
    return OGE_OK;
}

extern "C" int oge_gpu_set_sort_variant(int variant) {
    if (variant < 0 || variant > 0xFFFFFF) return fail_msg(OGE_ERR_INVALID_ARG, "set_sort_variant: %d", variant);
#ifndef OGE_TESTING
    // the product library only selects between its two (equally correct) pass kernels; the measurement knobs in
    // bits 4-7 exist in the -DOGE_TESTING build alone
    if ((variant & 0xFF) != 0 && (variant & 0xFF) != 2) return fail_msg(OGE_ERR_INVALID_ARG, "set_sort_variant: 0 or 2 (+ prefetch distance << 8)");
#endif
    radix_sort_set_variant(variant & 0xFF);
    radix_sort_set_prefetch(variant >> 8);
    return OGE_OK;
}
