/*
 * oge_gpu_dedup.h -- C ABI of the B200 duplicate-marking path (libopenge_b200.so).
 *
 * The reference (adaptivegenome/openge) has no FFI for this path: the operator interface is the
 * C++ class MarkDuplicates : AlgorithmModule (src/algorithms/mark_duplicates.h:27-68,
 * src/algorithms/algorithm_module.h:33-107).  This header is the boundary a drop-in
 * MarkDuplicates (openge_b200/host/mark_duplicates_gpu.{h,cpp}) binds instead of the CPU code.
 * Each entry point names the reference code it stands in for.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on success
 * and a negative OGE_ERR_* code on failure (the reference's convention is message + exit(-1),
 * e.g. util/bam_deserializer.h:155-163; the host shim maps codes back to that);
 * oge_gpu_last_error() gives the message.  A context is used by one host thread at a time
 * (one MarkDuplicates instance = one pthread, algorithm_module.cpp:118-122); several contexts
 * may coexist (split-by-chromosome mode, command_dedup.cpp:71-95).
 *
 * Record format = raw BAM records exactly as the reference (de)serialises them
 * (util/bam_deserializer.h:144-193, util/bam_serializer.h:106-141): 4-byte block_size, 32-byte
 * core, name, cigar, packed bases, qualities, tags; records back to back; offsets[n+1] byte
 * offsets relative to the first record of the batch.
 */
#ifndef OGE_GPU_DEDUP_H
#define OGE_GPU_DEDUP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OGE_GPU_DEDUP_ABI_VERSION 8

enum {
    OGE_OK = 0,
    OGE_ERR_INVALID_ARG = -1,
    OGE_ERR_CUDA = -2,          /* CUDA runtime failure or no usable device (never a CPU fallback) */
    OGE_ERR_NOMEM = -3,
    OGE_ERR_KEY_RANGE = -4,     /* a record's refID/coordinate/library does not fit the key layout */
    OGE_ERR_STATE = -5,         /* call order violated (e.g. flags before run) */
    OGE_ERR_TOO_LARGE = -6,     /* more than 2^30 records in one context */
    OGE_ERR_BAD_RECORD = -7     /* malformed record chain (block_size < 32, offsets not increasing) */
};

typedef struct oge_gpu_dedup_ctx oge_gpu_dedup_ctx;

/* Mirrors the knobs of MarkDuplicates + DedupCommand (mark_duplicates.h:41, command_dedup.cpp:39-48). */
typedef struct oge_gpu_dedup_config {
    int32_t abi_version;            /* OGE_GPU_DEDUP_ABI_VERSION */
    int32_t device;                 /* CUDA device ordinal */
    int32_t n_ref;                  /* number of @SQ entries (BamHeader::getSequences().size()) */
    int32_t max_ref_len;            /* largest @SQ LN; <=0 -> full 32-bit coordinate field */
    int32_t clip_margin;            /* slack for clipped ends beyond [0, max_ref_len); <=0 -> 1<<20 */
    int32_t remove_duplicates;      /* MarkDuplicates::removeDuplicates (-r); affects oge_gpu_dedup_pull only */
    int32_t verify_names;           /* hash-matched mates are confirmed by comparing the RG+":"+name bytes (ReadEndsMap is
                                       keyed by that string, picard_structures.h:82-109, mark_duplicates.cpp:210-214).
                                       Always on in the product library (the field is read by the -DOGE_TESTING build
                                       alone, where 0 measures hash-only pairing) */
    int32_t compat_quiet_index_bug; /* 1: reproduce non-verbose runs of the reference, where the record index
                                       never advances (mark_duplicates.cpp:250).  Default 0. */
    int32_t debug_keep_ends;        /* 1: keep a copy of the per-record end entries for oge_gpu_dedup_debug_ends */
    int32_t profile_events;         /* 1: bracket every radix-sort pass launch with CUDA events (stats.ms_sort_pass_kernels) */
    uint64_t capacity_records;      /* optional preallocation hints (0 = grow on demand) */
    uint64_t capacity_bytes;
    /* multi-GPU range sharding; world <= 1 means single GPU */
    int32_t rank;
    int32_t world;
    uint64_t index_base;            /* global ordinal of this shard's first record */
    int32_t debug_full_frag_sort;   /* 1: always sort every fragment end (measurement / A-B of the reduced fragment pass) */
    int32_t debug_legacy_join;      /* 1: separate end-build and whole-file hash join (the form the range-sharded path uses) instead of
                                       the end-build fused with the in-CTA join (measurement / A-B) */
} oge_gpu_dedup_config;

/* Per-run counters (the reference prints the analogous numbers under -v, mark_duplicates.cpp:261,433). */
typedef struct oge_gpu_dedup_stats {
    uint64_t n_records;
    uint64_t n_frag_entries;        /* fragSort.size() */
    uint64_t n_pair_entries;        /* pairSort.size() */
    uint64_t n_duplicates;          /* records whose 0x400 bit is set after the run (primary only) */
    uint64_t n_complex_names;       /* half-pair entries resolved on the exact slow path (name seen != 2 times) */
    uint64_t n_hash_mismatch;       /* hash-equal mates rejected by the name comparison */
    uint32_t frag_key_bits, pair_key_bits;
    uint32_t frag_sort_passes, pair_sort_passes;
    float ms_total;                 /* device time of the last oge_gpu_dedup_run, CUDA events */
    float ms_endbuild, ms_join, ms_sort_frag, ms_sort_pair, ms_select, ms_flags;
    uint64_t launches;              /* kernels launched by the last run */
    /* with profile_events: the onesweep pass kernel (K3's scatter pass) on its own */
    float ms_sort_pass_kernels;     /* summed duration of the pass launches of the last run */
    uint32_t sort_pass_launches;
    uint64_t sort_pass_bytes;       /* algorithmic bytes of those launches: 32 per entry (16 read + 16 written) */
    /* oge_gpu_dedup_push_bgzf: the inflate kernel on its own (CUDA events around the launch) */
    float ms_inflate;
    float ms_frame;                 /* oge_gpu_dedup_frame: guess + walk + proof + offsets write */
    uint64_t inflate_blocks, inflate_bytes_in, inflate_bytes_out;
    uint64_t frame_repairs;         /* chunks whose guessed entry the proof replaced */
    float ms_inflate_h2d, ms_inflate_d2h;   /* push_bgzf: upload of the compressed file, copy-back of the records (0 without host_copy) */
    /* fused end-build + in-CTA join: pairs settled inside the CTAs, records handed to the global join, local pairs the
     * check pass retracted (their names had leftovers) */
    uint64_t n_local_pairs, n_join_leftovers, n_local_retracted;
    /* with profile_events: the kernels of the last run on their own (CUDA events around each launch, ms).  Order:
     * OGE_K_ENDBUILD, OGE_K_MATCH (windowed join, K2a), OGE_K_EMIT (K2b), OGE_K_GLOBAL_JOIN (leftovers), OGE_K_CHECK,
     * OGE_K_SELECT (all launches), OGE_K_FLAGS, OGE_K_SORT_HIST (histogram + scan of every sort). */
    float ms_kernel[8];
    /* oge_gpu_dedup_push_bgzf as a whole (upload, inflate and copy-back overlapped piece by piece): device time from its
     * first to its last operation; which decoder ran (0 thread per block, 1 warp per block, 2 hardware decompress engine);
     * pieces the file was uploaded in.  With several pieces ms_inflate spans the inflates INCLUDING their waits for the
     * upload; the decoder's own rate shows with one piece (oge_gpu_set_bgzf_chunk_bytes(~0)). */
    float ms_push_bgzf;
    uint32_t inflate_mode, inflate_pieces;
    float ms_inflate_start;         /* device time from the start of push_bgzf to the start of the first piece's inflate */
    /* oge_gpu_dedup_deflate: time on the context's stream from its first to its last operation (bins, -r compaction, deflate,
     * packing, and the device allocations between them), blocks, bytes in (records) and out (BGZF members) */
    float ms_deflate;
    uint64_t deflate_blocks, deflate_bytes_in, deflate_bytes_out;
} oge_gpu_dedup_stats;
enum { OGE_K_ENDBUILD = 0, OGE_K_MATCH, OGE_K_EMIT, OGE_K_GLOBAL_JOIN, OGE_K_CHECK, OGE_K_SELECT, OGE_K_FLAGS, OGE_K_SORT_HIST };

/* Per-record view of the end-building kernel (buildReadEnds, mark_duplicates.cpp:147-164). */
typedef struct oge_gpu_end {
    int32_t eligible;               /* mapped && refID != -1 && primary (:202-205) */
    int32_t pair_eligible;          /* eligible && paired && mate mapped (:209) */
    int32_t ref;                    /* read1Sequence */
    int32_t coord;                  /* read1Coordinate: unclipped 5' end (:88-129) */
    int32_t orientation;            /* 1 = RE_F, 2 = RE_R (picard_structures.h:24-27) */
    int32_t read2Sequence;          /* != -1 iff ReadEnds::isPaired() (picard_structures.h:54); only the
                                       predicate is kept on the device: -1, or 0 for "paired" */
    int16_t score;                  /* getScore (:135-144) */
    int16_t lib;                    /* getLibraryId (:282-294) */
} oge_gpu_end;

/* MarkDuplicates::MarkDuplicates (mark_duplicates.cpp:27-39). */
int oge_gpu_dedup_create(const oge_gpu_dedup_config *cfg, oge_gpu_dedup_ctx **out);
void oge_gpu_dedup_destroy(oge_gpu_dedup_ctx *ctx);

/* The header lookup getLibraryName/getLibraryId does per record (mark_duplicates.cpp:282-318,
 * util/bam_header.h:214-241), resolved once on the host: ids[i] (NUL-terminated @RG ID, first
 * occurrence of an ID wins) -> lib_ids[i] in 1..n_libs; reads with no/unknown RG or an @RG without
 * LB get unknown_lib_id ("Unknown Library"). */
int oge_gpu_dedup_set_readgroups(oge_gpu_dedup_ctx *ctx, const char *const *ids, const int16_t *lib_ids,
                                 int32_t n, int16_t unknown_lib_id, int32_t n_libs);

/* Stands in for the getInputAlignment() loop + temp-file spill of buildSortedReadEndLists
 * (mark_duplicates.cpp:192-256): appends a batch of raw records to the device-resident record
 * array.  `records` should be pinned (oge_gpu_host_alloc) for the copy to overlap; the call
 * returns once the batch has been queued and the host buffer may be reused after
 * oge_gpu_dedup_sync().  May be called repeatedly; record ordinals continue across batches. */
int oge_gpu_dedup_push(oge_gpu_dedup_ctx *ctx, const uint8_t *records, uint64_t nbytes,
                       const uint64_t *offsets, uint64_t nrec);
int oge_gpu_dedup_sync(oge_gpu_dedup_ctx *ctx);

/* The same input side for a BGZF-compressed BAM file, with the inflate on the device: stands in for
 * BgzfInputStream::BgzfBlock::decompress (util/bgzf_input_stream.cpp:65-142; one zlib call per block, raw deflate,
 * window 15; like there the inflated size is checked and the CRC is not) for every block of the file, by the B200's
 * hardware decompress engine (or one of two kernels, oge_gpu_set_inflate_kernel).  The file goes up in pieces
 * (oge_gpu_set_bgzf_chunk_bytes): the upload of one piece runs under the inflate of the one before and under the
 * copy-back of the one before that, like the reference's reader keeps reading while its block jobs decompress
 * (bgzf_input_stream.cpp:144-207).  An invalid deflate stream fails with OGE_ERR_BAD_RECORD "Zlib inflate failed";
 * when the engine was decoding, the CUDA context of the process is lost with it (the reference exit(-1)s there).
 * comp = the whole file in host memory (pinned for the overlap); the block table comes from the file's block headers
 * (oge_bam_bgzf_index in oge_bam_host.h: offset, BSIZE + 1 and ISIZE of every block); header_bytes = inflated bytes in
 * front of the first record (magic, header text, reference list), which stay out of the record array.  host_copy
 * (optional) receives the inflated record bytes (total ISIZE - header_bytes) for the host side of the pipeline
 * (framing, the output file).  Synchronous.  Must be the only push of the context, followed by ONE
 * oge_gpu_dedup_set_offsets with the record chain framed from those bytes (BamDeserializer::read,
 * util/bam_deserializer.h:144-172). */
int oge_gpu_dedup_push_bgzf(oge_gpu_dedup_ctx *ctx, const uint8_t *comp, uint64_t comp_bytes, const uint64_t *block_in_off,
                            const uint32_t *block_csize, const uint32_t *block_isize, uint64_t n_blocks, uint64_t header_bytes,
                            uint8_t *host_copy);
int oge_gpu_dedup_set_offsets(oge_gpu_dedup_ctx *ctx, const uint64_t *offsets, uint64_t nrec);
/* ... or frame them on the device, in place of oge_gpu_dedup_set_offsets: BamDeserializer::read's chain walk
 * (util/bam_deserializer.h:144-172; same limits 32 <= block_size <= 10000, same messages) done speculatively in parallel
 * over 64 KB chunks and then proven -- every chunk must have been entered exactly where its predecessor's walk left,
 * chunk 0 at byte 0; chunks whose guess was wrong are re-walked from the proven position -- so the result is the
 * sequential chain, not a heuristic.  With this the host needs no copy of the records before the dedup runs
 * (host_copy may be NULL in push_bgzf).  oge_gpu_dedup_offsets downloads the n + 1 offsets. */
int oge_gpu_dedup_frame(oge_gpu_dedup_ctx *ctx, uint64_t *nrec_out);
int oge_gpu_dedup_offsets(oge_gpu_dedup_ctx *ctx, uint64_t *out, uint64_t n_plus_1);

/* buildSortedReadEndLists + generateDuplicateIndexes + the flag rewrite of runInternal
 * (mark_duplicates.cpp:185-279, 326-400, 443-465) over everything pushed so far.  Idempotent:
 * may be called again on the same resident records. */
int oge_gpu_dedup_run(oge_gpu_dedup_ctx *ctx);

/* ReadSorter in front of MarkDuplicates (`openge mergesort -M`: commands/command_mergesort.cpp:68-100,
 * algorithms/read_sorter.cpp:203-205,248-): sorts the resident records by coordinate on the device, in the order of
 * Sort::ByPosition (util/bamtools/Sort.h:108-133): refID (records without one last), position, strand (forward first),
 * name, flag.  Where the reference falls through to comparing the ADDRESSES of its heap objects (exact copies, and the
 * order inside the unplaced tail) the input order is kept.  Afterwards the context holds the records in sorted order:
 * run / flags / pull / flagstats refer to that order; oge_gpu_dedup_sort_order gives the permutation
 * (perm[k] = input ordinal of the record now at position k). */
int oge_gpu_dedup_sort(oge_gpu_dedup_ctx *ctx);
int oge_gpu_dedup_sort_order(oge_gpu_dedup_ctx *ctx, uint32_t *perm_out, uint64_t n);
/* what the last sort did: records that tied on (refID, position, strand), name-refinement rounds, kernel launches, device ms */
int oge_gpu_dedup_sort_stats(oge_gpu_dedup_ctx *ctx, uint64_t *n_tied, uint64_t *rounds, uint64_t *launches, float *ms);

/* Output side of runInternal (:443-465): the flag word of every record, in input order. */
int oge_gpu_dedup_flags(oge_gpu_dedup_ctx *ctx, uint16_t *out, uint64_t n);

/* Output side with records: copies the (flag-patched) records back; with remove_duplicates set,
 * records whose flag has 0x400 are dropped (:456-458) and the rest are compacted in order.
 * out_offsets (optional) receives out_nrec+1 offsets. */
int oge_gpu_dedup_pull(oge_gpu_dedup_ctx *ctx, uint8_t *out_records, uint64_t cap_bytes,
                       uint64_t *out_offsets, uint64_t cap_records, uint64_t *out_bytes, uint64_t *out_nrec);

/* Output side as a finished file body: the flag-patched records (with remove_duplicates: the ones that stay, :456-458), the
 * bin of every record recomputed as the reference's serialiser does (util/bam_serializer.h:88-126), cut into blocks of 65280
 * bytes and compressed into BGZF members ON THE DEVICE -- 18-byte header, raw deflate stream, CRC32, ISIZE, back to back: the
 * part of BgzfOutputStream (util/bgzf_output_stream.cpp:59-144, 170-250) between the BAM header and the empty end-of-file
 * block, which the host writes around it (oge_bam_store_members).  The deflate streams are NOT zlib's (a warp-parallel match
 * finder, its own prefix codes; compression in zlib level 1's class): the file is identical to the reference's after
 * decompression, not byte for byte -- for that, oge_gpu_dedup_pull + oge_bam_store.  deflate leaves the members resident and
 * reports their size, the number of blocks and of records; pull_bgzf copies them out. */
int oge_gpu_dedup_deflate(oge_gpu_dedup_ctx *ctx, uint64_t *out_bytes, uint64_t *out_blocks, uint64_t *out_nrec);
int oge_gpu_dedup_pull_bgzf(oge_gpu_dedup_ctx *ctx, uint8_t *out, uint64_t cap_bytes);
/* The same in parts, for a writer that streams the members through small pinned buffers into the file while the next part
 * is on its way (the reference's writer thread does the same with its block queue, bgzf_output_stream.cpp:170-223):
 * _part queues the copy of members[offset, offset + nbytes) on a side stream and returns; _wait returns when every queued
 * part has landed.  Any split is fine for a file (parts need not end on member boundaries). */
int oge_gpu_dedup_pull_bgzf_part(oge_gpu_dedup_ctx *ctx, uint64_t offset, uint64_t nbytes, uint8_t *out);
int oge_gpu_dedup_pull_bgzf_wait(oge_gpu_dedup_ctx *ctx);

/* The counters of the reference's Statistics module (algorithms/statistics.cpp:77-162: what `openge stats` prints,
 * and what a Statistics stage placed behind MarkDuplicates would count) over the resident records and their
 * flag words after the run, as device reductions; `sorted` is the "Sorted:" verdict (:89-101), including the
 * reference's exemption of the first record of every contig. */
typedef struct oge_gpu_flagstats {
    uint64_t n_reads, n_mapped, n_forward_strand, n_reverse_strand, n_failed_qc, n_duplicates;
    uint64_t n_paired, n_proper_pair, n_both_mates_mapped, n_first_mate, n_second_mate, n_singletons;
    uint64_t sorted;                /* 1 = "Yes", 0 = "No" */
} oge_gpu_flagstats;
int oge_gpu_dedup_flagstats(oge_gpu_dedup_ctx *ctx, oge_gpu_flagstats *out);

/* Forget the pushed records (keeps allocations); a context can then take the next file. */
int oge_gpu_dedup_reset(oge_gpu_dedup_ctx *ctx);

int oge_gpu_dedup_get_stats(oge_gpu_dedup_ctx *ctx, oge_gpu_dedup_stats *out);
int oge_gpu_dedup_debug_ends(oge_gpu_dedup_ctx *ctx, oge_gpu_end *out, uint64_t n);

/* Device-to-device copy on the context's stream, completed on return: how a caller takes a list that one of the
 * oge_gpu_shard_* calls handed out (valid only until the next call) into its own exchange buffer. */
int oge_gpu_copy_d2d(oge_gpu_dedup_ctx *ctx, void *dst, const void *src, uint64_t nbytes);

/* Device pointers of the resident arrays, for callers that keep working on the GPU
 * (and for bench.py's device-resident timing): records, offsets (u64, n+1), flags (u16, n). */
int oge_gpu_dedup_device_ptrs(oge_gpu_dedup_ctx *ctx, void **records, void **offsets, void **flags);

/* Test hook for the radix sort alone (K3): sorts n 16-byte little-endian 128-bit entries held in
 * host memory by the bit range [bit_lo, bit_hi), stably, on `device`. */
int oge_gpu_debug_sort128(int device, void *entries, uint64_t n, int bit_lo, int bit_hi);

/* ---- range sharding across GPUs (one context per rank; cfg.rank / cfg.world / cfg.index_base) -------------
 * The reference is a single process; these entry points exist so that G ranks, each holding a contiguous
 * record range of one coordinate-sorted file, produce exactly the flags of its single-stream run
 * (`openge dedup --nosplit -v`; the reference's own split mode, command_dedup.cpp:71-95, does not: SURVEY F2).
 * Between the calls the host moves small lists with all-to-all exchanges (NCCL all_to_all_single with uneven
 * splits): every output list comes ORDERED BY DESTINATION RANK with its per-destination item counts, every
 * input list is what the other ranks (and this one) addressed to this rank, concatenated.  Only the 8-byte key
 * hashes of the published entries go to every rank.  All pointers are DEVICE pointers, valid until the next
 * call on the context.  Items: published entry = entry_bytes (below), routed end entry = 32 bytes, mark = 4,
 * hash = 8.  Order: setup, key_bytes + set_entry_bytes once; then per run
 * begin -> probe -> replay -> finish -> apply; afterwards oge_gpu_dedup_flags / _pull / _get_stats. */
int oge_gpu_shard_setup(oge_gpu_dedup_ctx *ctx, uint64_t global_n, const uint64_t *bases /* world + 1 */,
                        const int32_t *split_ref, const int32_t *split_pos /* world - 1: first record of ranks 1.. */);
/* A published entry carries its whole pairing key RG + ":" + name (mark_duplicates.cpp:210-214), so that the rank
 * that owns the name compares keys byte by byte whatever their length: key_bytes = the longest key among this
 * rank's records; every rank then sets entry_bytes = 32 + the largest key_bytes of all ranks rounded up to 32. */
int oge_gpu_shard_key_bytes(oge_gpu_dedup_ctx *ctx, uint32_t *max_key_bytes);
int oge_gpu_shard_set_entry_bytes(oge_gpu_dedup_ctx *ctx, uint32_t entry_bytes);
/* K1 + mate join of the shard (mark_duplicates.cpp:192-256 on its records).  out: the records whose RG:name key
 * was not seen exactly twice on this rank -> the rank that owns the name (hash mod world); their key hashes ->
 * every rank; copies of the fragment ends whose (refID, unclipped coordinate) lies in another rank's key range
 * -> that rank. */
int oge_gpu_shard_begin(oge_gpu_dedup_ctx *ctx, void **pub_dev, uint64_t *pub_counts /* world */, void **hash_dev, uint64_t *n_hash,
                        void **frag_route_dev, uint64_t *frag_route_counts /* world */);
/* in: the key hashes the OTHER ranks published; the fragment ends routed to this rank.  A local pair of a name
 * published elsewhere is retracted and its two records published (round 2, to the name's owner); the fragment
 * K3 + K4 (mark_duplicates.cpp:262-271, 371-390) start on a side stream; out also: local pair ends whose key
 * lies in another rank's range -> that rank. */
int oge_gpu_shard_probe(oge_gpu_dedup_ctx *ctx, const void *hash_in_dev, uint64_t n_hash_in, const void *frag_route_in_dev,
                        uint64_t n_frag_route_in, void **pub2_dev, uint64_t *pub2_counts, void **pair_route_dev, uint64_t *pair_route_counts);
/* in: every published entry (both rounds, all ranks) of the names this rank owns.  Replays ReadEndsMap's toggle
 * (picard_structures.h:87-96) over them in global file order; pairs whose key range this rank owns join its
 * lists, out: the others -> the rank that owns their key range. */
int oge_gpu_shard_replay(oge_gpu_dedup_ctx *ctx, const void *pub_in_dev, uint64_t n_pub_in, void **owner_route_dev, uint64_t *owner_route_counts);
/* in: the pair ends routed to this rank (probe and replay outputs of all ranks).  Pair K3 + K4 (:336-355, 488-507).
 * out: global ordinals to mark -> the rank that holds the record. */
int oge_gpu_shard_finish(oge_gpu_dedup_ctx *ctx, const void *pair_route_in_dev, uint64_t n_pair_route_in, void **marks_dev,
                         uint64_t *marks_counts);
/* in: the marks addressed to this rank.  K5 (mark_duplicates.cpp:443-465). */
int oge_gpu_shard_apply(oge_gpu_dedup_ctx *ctx, const void *marks_in_dev, uint64_t n_marks_in);

/* The same step driven from C++, the four exchanges done by NCCL directly (grouped ncclSend / ncclRecv per peer = an
 * all-to-all with uneven splits, on the library's stream; NCCL is loaded at run time, libnccl.so.2).  One rank creates an id
 * (ncclGetUniqueId) and hands its 128 bytes to the others by any means; every rank then joins with its context (cfg.rank /
 * cfg.world); oge_gpu_shard_step runs begin -> ... -> apply.  comm_destroy is also done by oge_gpu_dedup_destroy. */
typedef struct oge_gpu_shard_step_info {
    uint64_t published_in, routed_in, marks_in;   /* items this rank received */
    uint64_t exchanges, bytes_sent;
} oge_gpu_shard_step_info;
int oge_gpu_shard_comm_id(uint8_t *id128);
int oge_gpu_shard_comm_init(oge_gpu_dedup_ctx *ctx, const uint8_t *id128);
void oge_gpu_shard_comm_destroy(oge_gpu_dedup_ctx *ctx);
int oge_gpu_shard_step(oge_gpu_dedup_ctx *ctx, oge_gpu_shard_step_info *info /* may be NULL */);

/* Measurement hook for K3 alone: sorts n device-generated entries (mode 0 uniform random, 1 = high key
 * bits follow the ordinal like a coordinate-sorted file) `reps` times after one warm-up; reports the
 * average CUDA-event time of one pass launch and of the whole sort, and verifies the result on the
 * device (no key inversion on [bit_lo, bit_hi), order-independent checksums unchanged). */
int oge_gpu_debug_sort_bench(int device, uint64_t n, int bit_lo, int bit_hi, int variant, int mode, int reps, uint64_t seed,
                             float *ms_pass_avg, float *ms_sort_avg, int *n_pass, int *verified);
/* Tuning hook: pass-kernel variant (0: 2048-entry tiles of 256 threads; 2: 4096-entry tiles of 512 threads, default). */
int oge_gpu_set_sort_variant(int variant);

/* Which BGZF decoder oge_gpu_dedup_push_bgzf uses: 2 = the hardware decompress engine (default where the device and
 * driver have one), 1 = one warp per block (default elsewhere), 0 = one thread per block (32 streams per warp as a
 * converged state machine), -1 = back to the default.  Also OGE_INFLATE_KERNEL=engine|warp|threads.
 * oge_gpu_inflate_kernel(device) -> the decoder push_bgzf would use on that device now (or a negative error). */
int oge_gpu_set_inflate_kernel(int kernel);
int oge_gpu_inflate_kernel(int device);
/* Tuning hook: compressed bytes per piece of push_bgzf's overlapped upload (0 = default: 64 MB for the engine, 512 MB for the
 * kernels; ~0 = the whole file in one piece, i.e. upload, then inflate, then copy-back). */
int oge_gpu_set_bgzf_chunk_bytes(uint64_t bytes);
/* A compressed file in PAGEABLE host memory goes up through two pinned staging buffers of the context (32 MB each by
 * default), filled by several host threads while the previous one is on its way.  stage_bytes = their size; 0 = leave
 * pageable sources to the driver's own staging (one copy stream: 11 GB/s on the B200 box against 55 GB/s from page-locked
 * memory).  Files smaller than two buffers go up directly. */
int oge_gpu_set_bgzf_staging(uint64_t stage_bytes);

/* Pinned host memory for push/pull buffers. */
void *oge_gpu_host_alloc(size_t nbytes);
void oge_gpu_host_free(void *p);

/* Number of CUDA devices visible (0 if none / driver missing). */
int oge_gpu_device_count(void);

const char *oge_gpu_last_error(void);
int oge_gpu_abi_version(void);
/* sizeof of the structs of this header as the library was compiled: 0 config, 1 stats, 2 oge_gpu_end, 3 flagstats
 * (a binding checks its own layout against these) */
int oge_gpu_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif
