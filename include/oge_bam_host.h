/*
 * oge_bam_host.h -- C ABI of the host-side BAM streaming layer around the GPU dedup path (liboge_bamhost.so).
 *
 * SURVEY 8(f) rows f1 + f2: what sits either side of MarkDuplicates in `openge dedup in.bam -o out.bam`
 * (commands/command_dedup.cpp:48-69: FileReader -> MarkDuplicates -> FileWriter), rebuilt as flat buffers
 * instead of one heap object per record:
 *
 *   reference                                                           here
 *   BgzfInputStream (util/bgzf_input_stream.cpp:65-142)                 oge_bam_load: the file's BGZF blocks are
 *     + BamDeserializer::open/read (util/bam_deserializer.h:40-193)       indexed, inflated in parallel straight into
 *                                                                         ONE (pinned) buffer, the header is parsed
 *                                                                         and the record chain framed into offsets[]:
 *                                                                         exactly what oge_gpu_dedup_push takes
 *   the flag rewrite + -r filter of runInternal                         oge_bam_apply_flags: flag words from
 *     (algorithms/mark_duplicates.cpp:443-465) and the bin the            oge_gpu_dedup_flags patched into the records
 *     writer recomputes per record (util/bam_serializer.h:106-126)        in place, bins recomputed, -r compaction
 *   FileWriter (algorithms/file_writer.cpp:69-170) + BamSerializer      oge_bam_store: header re-rendered the way
 *     ::open (util/bam_serializer.h:46-79) + BgzfOutputStream             BamHeader::toString does (util/bam_header.cpp:
 *     (util/bgzf_output_stream.cpp:59-250)                                184-262), optional @PG line, the stream cut
 *                                                                         into the reference's 65536-byte blocks and
 *                                                                         deflated in parallel with the reference's
 *                                                                         zlib parameters: the output FILE is byte-
 *                                                                         identical to the reference's
 *
 * No CUDA in this library: buffers come from an injected allocator (the CLI passes oge_gpu_host_alloc so that the
 * inflated records are pinned).  Nothing here computes duplicate flags -- that is libopenge_b200.so's job.
 * Conventions as in oge_gpu_dedup.h: 0 on success, negative on failure, oge_bam_last_error() has the message
 * (the reference prints a message and exit(-1)s on every one of these conditions; the CLI maps back to that).
 * Not rebuilt: SAM/FASTQ output, stdin/stdout streams.  (No .bai either -- and none in the reference: its BAM writer
 * collects index data for coordinate-sorted output, but BamIndex::writeFile returns before writing, util/bam_index.cpp:243.)
 */
#ifndef OGE_BAM_HOST_H
#define OGE_BAM_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    OGE_BAM_OK = 0,
    OGE_BAM_ERR_IO = -1,        /* cannot open / read / write */
    OGE_BAM_ERR_FORMAT = -2,    /* not BGZF/BAM, corrupt block, bad record chain, bad header line */
    OGE_BAM_ERR_NOMEM = -3,
    OGE_BAM_ERR_ARG = -4
};

typedef struct oge_bam_file oge_bam_file;
typedef void *(*oge_bam_alloc_fn)(size_t);
typedef void (*oge_bam_free_fn)(void *);

/* Reads `path` (BGZF-compressed BAM, or an uncompressed BAM stream = the reference's "rawbam"), inflates it with
 * `threads` workers into one buffer obtained from alloc_fn (NULL = malloc/free), parses the header and frames the
 * records.  Limits are the reference's: BGZF blocks must carry the 6-byte BC extra field
 * (bgzf_input_stream.cpp:84-98); 32 <= block_size <= 10000 (bam_deserializer.h:160). */
int oge_bam_load(const char *path, int threads, oge_bam_alloc_fn alloc_fn, oge_bam_free_fn free_fn, oge_bam_file **out);
void oge_bam_close(oge_bam_file *f);

/* Two-stage open for a caller that inflates elsewhere -- the GPU, oge_gpu_dedup_push_bgzf (oge_gpu_dedup.h):
 *   oge_bam_open_bgzf      reads the file, indexes its BGZF blocks and inflates only the leading block(s) on the host,
 *                          as far as the header reaches: header text, reference list and the stream offset of the
 *                          first record are known afterwards (everything below except records/offsets works);
 *   oge_bam_bgzf_index     the compressed bytes and the block table to hand to the inflater; header_bytes = inflated
 *                          bytes in front of the first record;
 *   oge_bam_records_buffer the (alloc_fn) buffer that has to receive the inflated bytes from header_bytes on
 *                          (total inflated size - header_bytes of them);
 *   oge_bam_frame_records  frames the record chain in that buffer (BamDeserializer::read, :144-172) and frees the
 *                          compressed copy.  From here on the object is the same as after oge_bam_load. */
int oge_bam_open_bgzf(const char *path, int threads, oge_bam_alloc_fn alloc_fn, oge_bam_free_fn free_fn, oge_bam_file **out);
int oge_bam_bgzf_index(const oge_bam_file *f, const uint8_t **comp, uint64_t *comp_bytes, const uint64_t **block_in_off,
                       const uint32_t **block_csize, const uint32_t **block_isize, uint64_t *n_blocks, uint64_t *header_bytes);
uint8_t *oge_bam_records_buffer(oge_bam_file *f);
int oge_bam_frame_records(oge_bam_file *f);
/* ... or take the offsets from whoever framed the records already (oge_gpu_dedup_frame + oge_gpu_dedup_offsets). */
int oge_bam_adopt_offsets(oge_bam_file *f, const uint64_t *offsets, uint64_t nrec);

/* What ReadSorter does to the header it hands on (algorithms/read_sorter.cpp:256-258: setSortOrder): the @HD SO value the
 * stored file will carry -- "coordinate" after oge_gpu_dedup_sort. */
int oge_bam_set_sort_order(oge_bam_file *f, const char *so);

const char *oge_bam_header_text(const oge_bam_file *f);        /* as stored in the file */
int32_t oge_bam_n_ref(const oge_bam_file *f);
const char *oge_bam_ref_name(const oge_bam_file *f, int32_t i);
int32_t oge_bam_ref_len(const oge_bam_file *f, int32_t i);
uint8_t *oge_bam_records(oge_bam_file *f);                     /* raw records back to back (block_size included) */
uint64_t oge_bam_records_bytes(const oge_bam_file *f);
const uint64_t *oge_bam_offsets(const oge_bam_file *f);        /* n + 1, relative to oge_bam_records() */
uint64_t oge_bam_n_records(const oge_bam_file *f);

/* @RG ID -> LB -> library id the way MarkDuplicates resolves them (mark_duplicates.cpp:282-318,
 * util/bam_header.h:214-241): the arguments of oge_gpu_dedup_set_readgroups.  Pointers stay valid until close. */
int oge_bam_library_table(oge_bam_file *f, const char *const **ids, const int16_t **lib_ids, int32_t *n,
                          int16_t *unknown_lib_id, int32_t *n_libs);

/* flags[n]: the output of oge_gpu_dedup_flags.  Writes every record's flag word and recomputes its bin from
 * (pos, end) as the reference's writer does for every record it serialises; with remove_duplicates the records
 * whose flag has 0x400 are dropped and the rest compacted in order. */
int oge_bam_apply_flags(oge_bam_file *f, const uint16_t *flags, int remove_duplicates, int threads);

/* format: "bam" (BGZF, compression `level`, default 6 in the reference: commands.cpp:122), "rawbam", or NULL =
 * by file name (".bam" / anything else -> bam, as FileWriter's default).  pg_command_line: NULL = --nopg; else the
 * CL: value of the @PG ID:openge line the reference's writer appends (file_writer.cpp:76-89). */
int oge_bam_store(oge_bam_file *f, const char *path, const char *format, int level, const char *pg_command_line,
                  const char *pg_version, int threads);

/* The same file with the record part already compressed: `members` = BGZF members holding the records (what
 * oge_gpu_dedup_deflate makes on the device).  Writes the re-rendered header in members of its own (zlib, `level`), the
 * given members, and the empty member that ends a BGZF file (util/bgzf_output_stream.cpp:225-250).  The file equals
 * oge_bam_store's (and the reference's) after decompression; its block boundaries and deflate streams differ. */
int oge_bam_store_members(oge_bam_file *f, const char *path, int level, const char *pg_command_line, const char *pg_version,
                          const uint8_t *members, uint64_t members_bytes);
/* ... with the members arriving in pieces: fill(user, &data, &nbytes) hands out the next piece (valid until the next call),
 * nbytes = 0 at the end, a non-zero return aborts.  Pieces need not end on member boundaries.  members_bytes_hint = the
 * total of all pieces when known (it must then be exact), 0 otherwise: with it the file is sized first and filled through
 * a shared mapping by up to `threads` threads (0 = all), without it the pieces are written one after the other. */
typedef int (*oge_bam_fill_fn)(void *user, const uint8_t **data, uint64_t *nbytes);
int oge_bam_store_members_stream(oge_bam_file *f, const char *path, int level, const char *pg_command_line, const char *pg_version,
                                 oge_bam_fill_fn fill, void *user, uint64_t members_bytes_hint, int threads);

/* seconds: [0] read file, [1] block scan, [2] inflate, [3] header + framing, [4] apply_flags, [5] store */
int oge_bam_timings(const oge_bam_file *f, double *out, int n);

/* ---- the codec and the header model on their own (tests, other callers) ---- */
/* BGZF <-> bytes.  *out is malloc'ed; free with oge_bam_buffer_free. */
int oge_bgzf_decompress(const uint8_t *in, size_t n, int threads, uint8_t **out, size_t *out_n);
int oge_bgzf_compress(const uint8_t *in, size_t n, int level, int threads, uint8_t **out, size_t *out_n);
/* BamHeader(text).toString() (util/bam_header.cpp:107-262).  *out is malloc'ed. */
int oge_bam_header_render(const char *text, char **out);
void oge_bam_buffer_free(void *p);

const char *oge_bam_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
