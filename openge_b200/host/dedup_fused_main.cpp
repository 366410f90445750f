// `openge dedup` as one fused host path around the GPU: BAM file in, BAM file out, no per-record objects.
//
//   reference (commands/command_dedup.cpp:37-114)                 here
//   FileReader -> MarkDuplicates -> FileWriter, one OGERead        oge_bam_load (parallel BGZF inflate into a pinned buffer,
//   per record on three pipeline threads, temp-file spill            framing) -> oge_gpu_dedup_push / _run / _flags (the CUDA
//                                                                    path) -> oge_bam_apply_flags -> oge_bam_store (parallel
//                                                                    BGZF deflate, byte-identical blocks)
//
// Same flags as the reference's command (commands/commands.cpp:117-132, command_dedup.cpp:29-35):
//   openge dedup [in.bam] -o out.bam [-r] [-v] [-t N] [-c level] [-F bam|rawbam] [--nopg] [--nosplit] [-T dir] [-d]
// BGZF input is inflated ON THE GPU (oge_gpu_dedup_push_bgzf: one warp per block; the compressed file crosses PCIe and the
// records are born in HBM); --cpu-inflate keeps the inflate on the host threads (oge_bam_load).  --pinned puts the host
// copy of the records in page-locked memory.
// --gpu-deflate makes the OUTPUT file's BGZF blocks on the GPU as well (oge_gpu_dedup_deflate: bins, -r filter, one warp per
// block deflate + CRC-32): the records never come back uncompressed and the host's zlib leaves the path; the file then is
// identical to the reference's after decompression instead of byte for byte (block boundaries and deflate streams differ).
// --nosplit, -T and -d are accepted and have no effect: the result is always that of the canonical single-chain run
// (`--nosplit -v`), nothing is spilled to disk, and there is one pipeline.  --stats prints the report of
// the reference's Statistics module (algorithms/statistics.cpp:150-174) for the output stream, counted on the device.
// Errors: message on stderr, exit(-1), as everywhere in the reference.
#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "oge_bam_host.h"
#include "oge_gpu_dedup.h"

#ifndef OGE_VERSION_STRING
#define OGE_VERSION_STRING "0.3-b200"
#endif

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void die(const char *what, const char *msg) {
    fprintf(stderr, "%s: %s Aborting.\n", what, msg);
    exit(-1);
}

// The device-made BGZF members on their way into the file: two pinned buffers, the copy of part k + 1 in flight while the
// writer has part k (oge_bam_store_members_stream pulls the parts through this).
struct MemberStream {
    oge_gpu_dedup_ctx *ctx;
    uint64_t total, issued, handed;
    uint8_t *buf[2];
    uint64_t len[2];
    int cur;
    static constexpr uint64_t PART = 64ull << 20;
    int issue(int which) {
        len[which] = total - issued < PART ? total - issued : PART;
        const int rc = len[which] ? oge_gpu_dedup_pull_bgzf_part(ctx, issued, len[which], buf[which]) : 0;
        issued += len[which];
        return rc;
    }
    static int fill(void *user, const uint8_t **data, uint64_t *nbytes) {
        MemberStream *m = (MemberStream *) user;
        if (m->handed == 0 && m->issued == 0 && m->issue(0)) return -1;
        if (oge_gpu_dedup_pull_bgzf_wait(m->ctx)) return -1;      // the current part has landed
        const int cur = m->cur;
        *data = m->buf[cur];
        *nbytes = m->len[cur];
        m->handed += m->len[cur];
        m->cur ^= 1;
        if (m->issue(m->cur)) return -1;      // the next one travels while the caller writes this one
        return 0;
    }
};

// ---- `--gpus N`: the records range-sharded over N GPUs of this node, one host thread and one context per GPU, the four
// exchanges done by NCCL inside the library (oge_gpu_shard_step).  The flags are those of the
// single-stream run, whatever N (tests/test_gpu_fused.py compares the output files).
struct ShardJob {
    int rank, world, device;
    oge_bam_file *bam;
    const uint64_t *bases;              // world + 1 record ordinals
    const int32_t *split_ref, *split_pos;
    const uint8_t *comm_id;
    uint16_t *flags;                    // the whole file's flag words; this rank fills [bases[rank], bases[rank + 1])
    uint32_t *key_bytes;                // [world]: every rank's longest pairing key
    pthread_barrier_t *barrier;
    std::atomic<int> *failed;           // set by any rank that cannot go on: nobody enters NCCL then
    oge_gpu_dedup_stats stats;
    oge_gpu_shard_step_info info;
    std::string error;
};

static void shard_worker(ShardJob *j) {
    auto fail = [&](const char *what) {
        j->error = std::string(what) + ": " + oge_gpu_last_error();
        j->failed->store(1);
    };
    oge_bam_file *bam = j->bam;
    const uint64_t lo = j->bases[j->rank], hi = j->bases[j->rank + 1], n = hi - lo;
    const uint64_t *off = oge_bam_offsets(bam);
    const uint8_t *rec = oge_bam_records(bam);
    oge_gpu_dedup_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = OGE_GPU_DEDUP_ABI_VERSION;
    cfg.device = j->device;
    cfg.n_ref = oge_bam_n_ref(bam);
    for (int32_t i = 0; i < cfg.n_ref; i++)
        if (oge_bam_ref_len(bam, i) > cfg.max_ref_len) cfg.max_ref_len = oge_bam_ref_len(bam, i);
    cfg.verify_names = -1;
    cfg.rank = j->rank;
    cfg.world = j->world;
    cfg.index_base = lo;
    cfg.capacity_records = n;
    cfg.capacity_bytes = off[hi] - off[lo];
    oge_gpu_dedup_ctx *ctx = NULL;
    bool ok = oge_gpu_dedup_create(&cfg, &ctx) == 0;
    if (!ok) fail("oge_gpu_dedup_create");
    if (ok) {
        const char *const *ids;
        const int16_t *libs;
        int32_t n_rg, n_libs;
        int16_t unknown;
        oge_bam_library_table(bam, &ids, &libs, &n_rg, &unknown, &n_libs);
        if (oge_gpu_dedup_set_readgroups(ctx, ids, libs, n_rg, unknown, n_libs)) { fail("set_readgroups"); ok = false; }
    }
    if (ok && oge_gpu_shard_setup(ctx, j->bases[j->world], j->bases, j->split_ref, j->split_pos)) { fail("oge_gpu_shard_setup"); ok = false; }
    std::vector<uint64_t> local_off;
    if (ok && n) {      // the shard's offsets, relative to its first record
        local_off.resize(n + 1);
        for (uint64_t i = 0; i <= n; i++) local_off[i] = off[lo + i] - off[lo];
        if (oge_gpu_dedup_push(ctx, rec + off[lo], off[hi] - off[lo], local_off.data(), n) || oge_gpu_dedup_sync(ctx)) { fail("oge_gpu_dedup_push"); ok = false; }
    }
    uint32_t kb = 0;
    if (ok && oge_gpu_shard_key_bytes(ctx, &kb)) { fail("oge_gpu_shard_key_bytes"); ok = false; }
    j->key_bytes[j->rank] = kb;
    // every rank reaches the barriers, failed or not: nobody is left waiting
    pthread_barrier_wait(j->barrier);
    uint32_t k = 0;
    for (int r = 0; r < j->world; r++) k = j->key_bytes[r] > k ? j->key_bytes[r] : k;
    const uint32_t entry_bytes = 32 + (k + 31) / 32 * 32 < 64 ? 64 : 32 + (k + 31) / 32 * 32;
    if (ok && oge_gpu_shard_set_entry_bytes(ctx, entry_bytes)) { fail("oge_gpu_shard_set_entry_bytes"); ok = false; }
    pthread_barrier_wait(j->barrier);
    if (j->failed->load()) ok = false;      // a rank that failed before this point would leave the others hanging in NCCL
    if (ok && oge_gpu_shard_comm_init(ctx, j->comm_id)) { fail("oge_gpu_shard_comm_init"); ok = false; }
    if (ok && oge_gpu_shard_step(ctx, &j->info)) { fail("oge_gpu_shard_step"); ok = false; }
    if (ok && n && oge_gpu_dedup_flags(ctx, j->flags + lo, n)) { fail("oge_gpu_dedup_flags"); ok = false; }
    if (ok) oge_gpu_dedup_get_stats(ctx, &j->stats);
    if (ctx) oge_gpu_dedup_destroy(ctx);
}

static void usage() {
    fprintf(stderr,
            "usage: oge_dedup_fused [dedup] in.bam -o out.bam [-r] [-v] [-t threads] [-c level] [-F bam|rawbam] [--nopg] [--stats]\n"
            "                       [--device N] [--cpu-inflate] [--gpu-deflate] [--pinned] [--tidy] [--nosplit] [-T tmpdir] [-d]\n"
            "                       [--sort | -M]   coordinate sort on the GPU in front of the dedup (= openge mergesort -M)\n"
            "                       [--gpus N]      range-shard the records over N GPUs of this node (devices --device .. --device + N - 1)\n");
    exit(-1);
}

int main(int argc, char **argv) {
    std::string in, out, format;
    bool remove_dups = false, verbose = false, nopg = false, stats = false, cpu_inflate = false, pinned = false, sort_first = false, gpu_deflate = false, tidy = false;
    int threads = 0, level = 6, device = 0, gpus = 1;
    std::string command_line = "openge ";      // commands/commands.cpp:36-40
    for (int i = 1; i < argc; i++) {
        command_line += argv[i];
        command_line += " ";
    }
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&]() -> const char * {
            if (i + 1 >= argc) usage();
            return argv[++i];
        };
        if (a == "dedup" && i == 1) continue;
        else if (a == "-o" || a == "--out") out = need();
        else if (a == "-r" || a == "--remove") remove_dups = true;
        else if (a == "-v" || a == "--verbose") verbose = true;
        else if (a == "-t" || a == "--threads") threads = atoi(need());
        else if (a == "-c" || a == "--compression") level = atoi(need());
        else if (a == "-F" || a == "--format") format = need();
        else if (a == "-T" || a == "--tmpdir") need();
        else if (a == "--nopg") nopg = true;
        else if (a == "--nosplit" || a == "-d" || a == "--nothreads") continue;
        else if (a == "--stats") stats = true;
        else if (a == "--sort" || a == "-M") sort_first = true;      // `openge mergesort -M`: coordinate sort in front of the dedup
        else if (a == "--cpu-inflate") cpu_inflate = true;
        else if (a == "--gpu-deflate") gpu_deflate = true;
        else if (a == "--tidy") tidy = true;      // free every buffer before exit (leak checkers); default: leave it to the exit
        else if (a == "--pinned") pinned = true;
        else if (a == "--device") device = atoi(need());
        else if (a == "--gpus") gpus = atoi(need());
        else if (!a.empty() && a[0] == '-') usage();
        else if (in.empty()) in = a;
        else if (out.empty()) out = a;
        else die("oge_dedup_fused", "one input file only (merge inputs with `openge mergesort` first).");
    }
    if (in.empty() || out.empty()) usage();
    if (format == "rawbam") gpu_deflate = false;      // nothing to compress
    if (gpus < 1) usage();
    if (gpus > 1) {
        if (sort_first) die("oge_dedup_fused", "--sort and --gpus do not go together (the device sort works on one GPU).");
        if (stats) die("oge_dedup_fused", "--stats and --gpus do not go together (the counters are reduced on one GPU).");
        if (oge_gpu_device_count() < device + gpus) die("MarkDuplicates (GPU)", "fewer CUDA devices than --gpus asks for.");
        cpu_inflate = true;      // the shards are cut from the framed records on the host
        gpu_deflate = false;     // the output is written by the host writer (byte-identical)
    }

    const double t_start = now_s();
    oge_bam_file *bam = NULL;
    // (Bringing the CUDA context up on a second thread while the file is read was measured and dropped: the two contend for the
    // process's address-space lock -- context creation took 0.70 s next to the read against 0.25 s after it.)
    if (oge_gpu_device_count() < 1) die("MarkDuplicates (GPU)", "no CUDA device: this path has no CPU fallback.");
    oge_bam_alloc_fn alloc_fn = pinned ? oge_gpu_host_alloc : NULL;
    oge_bam_free_fn free_fn = pinned ? oge_gpu_host_free : NULL;
    // BGZF input: open in two stages and let the GPU inflate; an uncompressed stream (or --cpu-inflate) loads on the host
    bool gpu_inflate = !cpu_inflate;
    int rc = 0;
    if (gpu_inflate) {
        rc = oge_bam_open_bgzf(in.c_str(), threads, alloc_fn, free_fn, &bam);
        if (rc == OGE_BAM_ERR_FORMAT && strstr(oge_bam_last_error(), "uncompressed BAM stream")) gpu_inflate = false;
        else if (rc) die("Error reading BAM", oge_bam_last_error());
    }
    if (!gpu_inflate && (rc = oge_bam_load(in.c_str(), threads, alloc_fn, free_fn, &bam))) die("Error reading BAM", oge_bam_last_error());
    const double t_loaded = now_s();

    if (gpus > 1) {
        const uint64_t n = oge_bam_n_records(bam);
        if (verbose) fprintf(stderr, "Read %llu records.\n", (unsigned long long) n);
        std::vector<uint64_t> bases(gpus + 1);
        std::vector<int32_t> split_ref(gpus - 1), split_pos(gpus - 1);
        const uint64_t *off = oge_bam_offsets(bam);
        const uint8_t *rec = oge_bam_records(bam);
        for (int r = 0; r <= gpus; r++) bases[r] = n * (uint64_t) r / gpus;
        for (int r = 1; r < gpus; r++) {      // the key range of rank r starts at its first record; an empty tail shard owns no key
            split_ref[r - 1] = split_pos[r - 1] = -1;
            if (bases[r] < n) {
                memcpy(&split_ref[r - 1], rec + off[bases[r]] + 4, 4);
                memcpy(&split_pos[r - 1], rec + off[bases[r]] + 8, 4);
            }
        }
        uint8_t comm_id[128];
        if (oge_gpu_shard_comm_id(comm_id)) die("MarkDuplicates (GPU): NCCL", oge_gpu_last_error());
        std::vector<uint16_t> flags(n ? n : 1);
        std::vector<uint32_t> key_bytes(gpus, 0);
        pthread_barrier_t barrier;
        pthread_barrier_init(&barrier, NULL, gpus);
        std::atomic<int> failed(0);
        std::vector<ShardJob> jobs(gpus);
        std::vector<std::thread> workers;
        for (int r = 0; r < gpus; r++) {
            ShardJob &j = jobs[r];
            j.rank = r; j.world = gpus; j.device = device + r; j.bam = bam; j.bases = bases.data();
            j.split_ref = split_ref.data(); j.split_pos = split_pos.data(); j.comm_id = comm_id; j.flags = flags.data();
            j.key_bytes = key_bytes.data(); j.barrier = &barrier; j.failed = &failed;
            memset(&j.stats, 0, sizeof(j.stats));
            memset(&j.info, 0, sizeof(j.info));
            workers.emplace_back(shard_worker, &j);
        }
        for (auto &w : workers) w.join();
        pthread_barrier_destroy(&barrier);
        uint64_t dups = 0, published = 0, routed = 0;
        float ms = 0;
        for (int r = 0; r < gpus; r++) {
            if (!jobs[r].error.empty()) die("MarkDuplicates (GPU)", jobs[r].error.c_str());
            dups += jobs[r].stats.n_duplicates;
            published += jobs[r].info.published_in;
            routed += jobs[r].info.routed_in;
            ms = jobs[r].stats.ms_total > ms ? jobs[r].stats.ms_total : ms;
        }
        const double t_gpu = now_s();
        if (verbose) {
            fprintf(stderr, "Range-sharded over %d GPUs: %llu published entries and %llu routed ends crossed ranks; %.3f ms on the slowest device.\n", gpus,
                    (unsigned long long) published, (unsigned long long) routed, ms);
            fprintf(stderr, "Marking %llu records as duplicates.\n", (unsigned long long) dups);
        }
        if ((rc = oge_bam_apply_flags(bam, flags.data(), remove_dups ? 1 : 0, threads))) die("Error rewriting records", oge_bam_last_error());
        if ((rc = oge_bam_store(bam, out.c_str(), format.empty() ? NULL : format.c_str(), level, nopg ? NULL : command_line.c_str(),
                                OGE_VERSION_STRING, threads)))
            die("Error writing BAM", oge_bam_last_error());
        if (verbose) {
            double t[6];
            oge_bam_timings(bam, t, 6);
            fprintf(stderr, "Written %llu records.\n", (unsigned long long) oge_bam_n_records(bam));
            fprintf(stderr, "Timing: load %.3f s | %d gpus %.3f s | rewrite %.3f s | store %.3f s | total %.3f s\n", t_loaded - t_start, gpus, t_gpu - t_loaded,
                    t[4], t[5], now_s() - t_start);
        }
        if (tidy) {
            oge_bam_close(bam);
            return 0;
        }
        fflush(stdout);
        fflush(stderr);
        _exit(0);
    }

    oge_gpu_dedup_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = OGE_GPU_DEDUP_ABI_VERSION;
    cfg.device = device;
    cfg.n_ref = oge_bam_n_ref(bam);
    for (int32_t i = 0; i < cfg.n_ref; i++)
        if (oge_bam_ref_len(bam, i) > cfg.max_ref_len) cfg.max_ref_len = oge_bam_ref_len(bam, i);
    cfg.remove_duplicates = gpu_deflate && remove_dups ? 1 : 0;      // else -r is applied by oge_bam_apply_flags: oge_gpu_dedup_pull returns every record
    cfg.verify_names = -1;
    if (!gpu_inflate) {
        cfg.capacity_records = oge_bam_n_records(bam);
        cfg.capacity_bytes = oge_bam_records_bytes(bam);
    }
    oge_gpu_dedup_ctx *ctx = NULL;
    if ((rc = oge_gpu_dedup_create(&cfg, &ctx))) die("MarkDuplicates (GPU): oge_gpu_dedup_create", oge_gpu_last_error());
    {
        const char *const *ids;
        const int16_t *libs;
        int32_t n_rg, n_libs;
        int16_t unknown;
        oge_bam_library_table(bam, &ids, &libs, &n_rg, &unknown, &n_libs);
        if ((rc = oge_gpu_dedup_set_readgroups(ctx, ids, libs, n_rg, unknown, n_libs))) die("MarkDuplicates (GPU): set_readgroups", oge_gpu_last_error());
    }
    double t_inflated = t_loaded, t_framed = t_loaded;
    uint64_t n = 0;
    std::vector<uint16_t> flags;
    if (gpu_inflate) {
        // compressed bytes up; inflate, framing and dedup on the device; ONE copy of the flag-patched records back
        const uint8_t *comp;
        const uint64_t *in_off;
        const uint32_t *csize, *isize;
        uint64_t comp_bytes, n_blocks, header_bytes;
        oge_bam_bgzf_index(bam, &comp, &comp_bytes, &in_off, &csize, &isize, &n_blocks, &header_bytes);
        uint8_t *host_records = NULL;
        // page the host buffer in meanwhile (not needed when the output is compressed on the device: the records stay there)
        std::thread prefault([&] { if (!gpu_deflate) host_records = oge_bam_records_buffer(bam); });
        if ((rc = oge_gpu_dedup_push_bgzf(ctx, comp, comp_bytes, in_off, csize, isize, n_blocks, header_bytes, NULL))) {
            prefault.join();
            die("Error reading BAM", oge_gpu_last_error());
        }
        t_inflated = now_s();
        if ((rc = oge_gpu_dedup_frame(ctx, &n))) {
            prefault.join();
            die("Error reading BAM", oge_gpu_last_error());
        }
        t_framed = now_s();
        if (verbose) fprintf(stderr, "Read %llu records.\n", (unsigned long long) n);
        if (sort_first && (rc = oge_gpu_dedup_sort(ctx))) die("ReadSorter (GPU): sort", oge_gpu_last_error());
        if ((rc = oge_gpu_dedup_run(ctx))) die("MarkDuplicates (GPU): run", oge_gpu_last_error());
        flags.resize(n ? n : 1);
        if ((rc = oge_gpu_dedup_flags(ctx, flags.data(), n))) die("MarkDuplicates (GPU): flags", oge_gpu_last_error());
        prefault.join();
        if (!gpu_deflate) {
        if (!host_records) die("Error reading BAM", oge_bam_last_error());
        std::vector<uint64_t> offs(n + 1);
        uint64_t got_bytes = 0, got_n = 0;
        uint64_t total = 0;
        for (uint64_t b = 0; b < n_blocks; b++) total += isize[b];
        if ((rc = oge_gpu_dedup_pull(ctx, host_records, total - header_bytes, offs.data(), n + 1, &got_bytes, &got_n))) die("MarkDuplicates (GPU): pull", oge_gpu_last_error());
        if (n == 0) offs[0] = 0;
        if ((rc = oge_bam_adopt_offsets(bam, offs.data(), n))) die("Error reading BAM", oge_bam_last_error());
        }
    } else {
        n = oge_bam_n_records(bam);
        if (verbose) fprintf(stderr, "Read %llu records.\n", (unsigned long long) n);
        if ((rc = oge_gpu_dedup_push(ctx, oge_bam_records(bam), oge_bam_records_bytes(bam), oge_bam_offsets(bam), n)))
            die("MarkDuplicates (GPU): push", oge_gpu_last_error());
        if (sort_first && (rc = oge_gpu_dedup_sort(ctx))) die("ReadSorter (GPU): sort", oge_gpu_last_error());
        if ((rc = oge_gpu_dedup_run(ctx))) die("MarkDuplicates (GPU): run", oge_gpu_last_error());
        flags.resize(n ? n : 1);
        if ((rc = oge_gpu_dedup_flags(ctx, flags.data(), n))) die("MarkDuplicates (GPU): flags", oge_gpu_last_error());
        if (sort_first && n && !gpu_deflate) {      // the records come back in their new order
            std::vector<uint64_t> offs(n + 1);
            uint64_t got_bytes = 0, got_n = 0;
            if ((rc = oge_gpu_dedup_pull(ctx, oge_bam_records(bam), oge_bam_records_bytes(bam), offs.data(), n + 1, &got_bytes, &got_n)))
                die("ReadSorter (GPU): pull", oge_gpu_last_error());
            if ((rc = oge_bam_adopt_offsets(bam, offs.data(), n))) die("Error reading BAM", oge_bam_last_error());
        }
    }
    if (sort_first) {
        oge_bam_set_sort_order(bam, "coordinate");      // read_sorter.cpp:256-258
        if (verbose) {
            uint64_t tied = 0, rounds = 0, sl = 0;
            float ms = 0;
            oge_gpu_dedup_sort_stats(ctx, &tied, &rounds, &sl, &ms);
            fprintf(stderr, "Sorted %llu records by coordinate on the GPU in %.3f ms (%llu tied on position, %llu name rounds, %llu kernel launches).\n",
                    (unsigned long long) n, ms, (unsigned long long) tied, (unsigned long long) rounds, (unsigned long long) sl);
        }
    }
    oge_gpu_dedup_stats st;
    oge_gpu_dedup_get_stats(ctx, &st);
    oge_gpu_flagstats fs;
    memset(&fs, 0, sizeof(fs));
    if (stats && (rc = oge_gpu_dedup_flagstats(ctx, &fs))) die("Statistics (GPU)", oge_gpu_last_error());
    uint64_t members_bytes = 0, member_blocks = 0, n_written = 0;
    double t_gpu = 0, t_mark[4] = {0, 0, 0, 0};
    if (gpu_deflate) {      // the output's BGZF members, made where the records are, streamed into the file
        if ((rc = oge_gpu_dedup_deflate(ctx, &members_bytes, &member_blocks, &n_written))) die("Error writing BAM", oge_gpu_last_error());
        oge_gpu_dedup_get_stats(ctx, &st);
        t_gpu = now_s();
        t_mark[0] = t_gpu;
        MemberStream ms;
        ms.ctx = ctx;
        ms.total = members_bytes;
        ms.issued = ms.handed = 0;
        ms.cur = 0;
        ms.len[0] = ms.len[1] = 0;
        ms.buf[0] = (uint8_t *) oge_gpu_host_alloc(MemberStream::PART);
        ms.buf[1] = (uint8_t *) oge_gpu_host_alloc(MemberStream::PART);
        if (!ms.buf[0] || !ms.buf[1]) die("Error writing BAM", "cannot allocate the output staging buffers.");
        t_mark[1] = now_s();
        if ((rc = oge_bam_store_members_stream(bam, out.c_str(), level, nopg ? NULL : command_line.c_str(), OGE_VERSION_STRING, MemberStream::fill, &ms, members_bytes, threads)))
            die("Error writing BAM", oge_bam_last_error());
        t_mark[2] = now_s();
        oge_gpu_host_free(ms.buf[0]);
        oge_gpu_host_free(ms.buf[1]);
        t_mark[3] = now_s();
    }
    if (tidy) oge_gpu_dedup_destroy(ctx);      // else the process exit does it: cudaFree of tens of GB costs 24 ms per GB
    if (!gpu_deflate) t_gpu = now_s();
    if (verbose) {
        fprintf(stderr, "Sorted %llu pair ends and %llu fragment ends on the GPU in %.3f ms (%llu kernel launches).\n",
                (unsigned long long) st.n_pair_entries, (unsigned long long) st.n_frag_entries, st.ms_total, (unsigned long long) st.launches);
        fprintf(stderr, "Marking %llu records as duplicates.\n", (unsigned long long) st.n_duplicates);
    }

    if (!gpu_deflate) {
        if ((rc = oge_bam_apply_flags(bam, flags.data(), remove_dups ? 1 : 0, threads))) die("Error rewriting records", oge_bam_last_error());
        if ((rc = oge_bam_store(bam, out.c_str(), format.empty() ? NULL : format.c_str(), level, nopg ? NULL : command_line.c_str(),
                                OGE_VERSION_STRING, threads)))
            die("Error writing BAM", oge_bam_last_error());
        n_written = oge_bam_n_records(bam);
    }
    const double t_end = now_s();

    if (stats) {      // the report of algorithms/statistics.cpp:150-174 (with -r the reference's Statistics stage would sit
                      // behind the filter; this one counts before it)
        const double nr = fs.n_reads ? (double) fs.n_reads : 1.0, np = fs.n_paired ? (double) fs.n_paired : 1.0;
        printf("Total reads:       %10llu\n", (unsigned long long) fs.n_reads);
        printf("Mapped reads:      %10llu (%5.1f%%)\n", (unsigned long long) fs.n_mapped, (float) fs.n_mapped / nr * 100);
        printf("Forward strand:    %10llu (%5.1f%%)\n", (unsigned long long) fs.n_forward_strand, (float) fs.n_forward_strand / nr * 100);
        printf("Reverse strand:    %10llu (%5.1f%%)\n", (unsigned long long) fs.n_reverse_strand, (float) fs.n_reverse_strand / nr * 100);
        printf("Failed QC:         %10llu (%5.1f%%)\n", (unsigned long long) fs.n_failed_qc, (float) fs.n_failed_qc / nr * 100);
        printf("Duplicates:        %10llu (%5.1f%%)\n", (unsigned long long) fs.n_duplicates, (float) fs.n_duplicates / nr * 100);
        printf("Paired-end reads:  %10llu (%5.1f%%)\n", (unsigned long long) fs.n_paired, (float) fs.n_paired / nr * 100);
        if (fs.n_paired) {
            printf("'Proper-pairs':    %10llu (%5.1f%%)\n", (unsigned long long) fs.n_proper_pair, (float) fs.n_proper_pair / np * 100);
            printf("Both pairs mapped: %10llu (%5.1f%%)\n", (unsigned long long) fs.n_both_mates_mapped, (float) fs.n_both_mates_mapped / np * 100);
            printf("Read 1:            %10llu\n", (unsigned long long) fs.n_first_mate);
            printf("Read 2:            %10llu\n", (unsigned long long) fs.n_second_mate);
            printf("Singletons:        %10llu (%5.1f%%)\n", (unsigned long long) fs.n_singletons, (float) fs.n_singletons / np * 100);
        }
        printf("Sorted:            %10s\n", fs.sorted ? "Yes" : "No");
    }
    if (verbose) {
        double t[6];
        oge_bam_timings(bam, t, 6);
        fprintf(stderr, "Written %llu records.\n", (unsigned long long) n_written);
        if (gpu_deflate)
            fprintf(stderr, "Timing (output): staging buffers %.3f s, members streamed into the file %.3f s, staging freed %.3f s\n", t_mark[1] - t_mark[0],
                    t_mark[2] - t_mark[1], t_mark[3] - t_mark[2]);
        if (gpu_deflate)
            fprintf(stderr, "gpu deflate: %llu blocks, %.1f MB -> %.1f MB in %.3f ms on the device (%.1f GB/s in).\n", (unsigned long long) member_blocks,
                    st.deflate_bytes_in / 1e6, st.deflate_bytes_out / 1e6, st.ms_deflate, st.ms_deflate > 0 ? st.deflate_bytes_in / 1e6 / st.ms_deflate : 0.0);
        if (gpu_inflate)
            fprintf(stderr,
                    "Timing: open %.3f s (read %.3f, scan %.3f) | gpu inflate %.3f s (upload %.1f ms, kernel %.3f ms = %.1f GB/s out) | gpu frame %.3f s "
                    "(device %.3f ms, %llu repairs) | gpu dedup + copy back %.3f s (device %.3f ms) | rewrite %.3f s | store %.3f s | total %.3f s\n",
                    t_loaded - t_start, t[0], t[1], t_inflated - t_loaded, st.ms_inflate_h2d, st.ms_inflate,
                    st.ms_inflate > 0 ? st.inflate_bytes_out / 1e6 / st.ms_inflate : 0.0, t_framed - t_inflated, st.ms_frame,
                    (unsigned long long) st.frame_repairs, t_gpu - t_framed, st.ms_total, t[4], t[5], t_end - t_start);
        else
            fprintf(stderr,
                    "Timing: load %.3f s (read %.3f, scan %.3f, inflate %.3f, frame %.3f) | gpu %.3f s (device %.3f ms) | rewrite %.3f s | store %.3f s | total %.3f s\n",
                    t_loaded - t_start, t[0], t[1], t[2], t[3], t_gpu - t_loaded, st.ms_total, t[4], t[5], t_end - t_start);
    }
    if (tidy) {
        oge_bam_close(bam);
        return 0;
    }
    fflush(stdout);
    fflush(stderr);
    _exit(0);      // the output file is closed; tearing down the CUDA context and unmapping the buffers is the kernel's job now
}
