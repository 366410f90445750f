// Kernel parameter blocks and launchers of the dedup path (one .cu per stage).
#pragma once
#include "common.cuh"

struct oge_gpu_dedup_ctx;      // ctx.cuh

namespace oge {

// ---- K1 end-build (endbuild.cu) ---------------------------------------------------------------
constexpr int EB_THREADS = 128;     // records per tile = threads per CTA
constexpr int EB_STAGES = 2;
static_assert(EB_STAGES >= 2, "stage metadata is only ordered by the next iteration's barrier");

constexpr uint32_t RGC_UNKNOWN = 0xFFFEu;   // RG value not listed in the header
constexpr uint32_t RGC_ABSENT = 0xFFFFu;    // no RG tag, or an empty value

// Host-resolved @RG table (reference: util/bam_header.h:214-241 lookups done per record by
// mark_duplicates.cpp:282-318): ids concatenated in `bytes`, id i = bytes[off[i], off[i+1]).
struct RgTable {
    const uint8_t *bytes;
    const uint32_t *off;
    const int16_t *lib;
    int n;
    int16_t unknown_lib;
};

// Compact copy of what the pairing key is made of (mark_duplicates.cpp:210-214): read-group code,
// name length and the first 29 name bytes, zero padded, in one 32-byte sector per record.  Two keys
// with l_name - 1 <= 29 are equal iff their tags are bit-equal (and the code is a listed read group);
// longer names are compared tag first, then tail against tail in the records.  The join reads
// these instead of the records: 32 B per record, and the mate's tag is an L2 hit.
struct __align__(32) NameTag {
    uint32_t w[8];      // w[0] = rgcode | l_name << 16 | name[0] << 24, w[1..7] = name[1..28]
};
constexpr uint32_t NAME_TAG_BYTES = 29;

struct EndbuildParams {
    const uint8_t *rec;
    const uint64_t *off;
    uint64_t n;
    uint64_t idx_base;      // global ordinal of record 0 (multi-GPU shards)
    E128 *frag;
    uint64_t *hk;
    NameTag *tag;           // [n] compact pairing-key copy, written for records that enter the mate map
    uint16_t *flag_in;
    uint32_t *counters;
    RgTable rg;
    KeyLayout kl;
    // range sharding: copies of the fragment ends whose packed key (ref << coord_bits | biased coord) lies outside this
    // rank's key range [own_lo, own_hi) are listed for their owners right here (RouteEntry list, counters[CNT_ROUTE])
    void *route_out = nullptr;
    uint32_t route_cap = 0;
    uint64_t own_lo = 0, own_hi = ~0ull;
};

int launch_endbuild(const EndbuildParams &P, uint32_t avg_rec_bytes, int sms, cudaStream_t stream, uint64_t *launches);

// ---- K2, windowed form (join.cu: local_join_kernel) -------------------------
constexpr int LJ_WAYS = 8;                   // ways per bucket of the in-CTA table
constexpr int LJ_ENTRY_BYTES = 4 + 12;       // key word + payload
constexpr int LJ_CTAS_PER_SM = 6;
constexpr int LJ_BUCKETS = 256;              // 2048 entries = 32 KB of shared memory per CTA
constexpr int LJ_THREADS = 256;
constexpr int LJ_ITEMS = 4;
constexpr int LJ_TILE = LJ_THREADS * LJ_ITEMS;
constexpr uint32_t LJ_HORIZON = 16384;       // an entry unmatched for this many records leaves for the global join
constexpr uint32_t LJ_SWEEP_TILES = 4;       // how often the table is swept for such entries

struct LocalJoinParams {
    E128 *pair, *pair_far;              // pair lists (counters[CNT_PAIRS], counters[CNT_PAIRS_FAR])
    uint64_t *pair_hk, *pair_far_hk;    // key hash of every pair entry, by list position
    uint32_t pair_cap, far_cap;
    uint32_t *mate_of;                  // [n] GLOBAL ordinal of the pair's other record, indexed by idx1's local ordinal
    uint32_t *left;                     // [n] local ordinals handed to the global join (counters[CNT_LEFT])
    uint4 *couples;                     // (taker, entry, hash lo, hash hi): CTA b owns [b * couples_per_cta, + couple_count[b])
    uint32_t *couple_count;             // [grid]
    uint32_t couples_per_cta;
    uint32_t n_buckets;                 // per CTA
    uint32_t tiles_per_cta;
};
void local_join_shape(uint64_t n, int sms, uint32_t *grid_out, uint32_t *tiles_per_cta);

// ---- K2 mate join (join.cu) -------------------------------------------------------------------
struct __align__(32) MateSlot {      // one 32-byte sector
    uint64_t key;       // 64-bit hash of RG + ":" + name; 0 = empty
    uint64_t val;       // (arrivals << 32) + sum of (ordinal + 1) over the arrivals
    uint64_t who;       // written by the second arrival: (its ordinal << 32) | the first arrival's
    uint32_t pair_pos;  // where it put their pair entry, or SLOT_NO_PAIR
    uint32_t pad;
};

struct JoinParams {
    const uint8_t *rec;
    const uint64_t *off;
    uint64_t n;
    uint64_t idx_base;
    const E128 *frag;
    const uint64_t *hk;
    const NameTag *tag;
    MateSlot *table;
    uint64_t n_slots;
    E128 *pair;             // output near-pair entries (appended; counters[CNT_PAIRS])
    E128 *pair_far;         // output far-pair entries (appended; counters[CNT_PAIRS_FAR])
    uint32_t *mate_of;      // [n] GLOBAL ordinal of the pair's other record, indexed by idx1's local ordinal
    E128 *cplx;             // output: (hash << 32 | local ordinal) of records for the exact path
    uint32_t *cplx_slots;   // output: slots that saw a third arrival (counters[CNT_COMPLEX_SLOTS])
    const uint32_t *list = nullptr;   // when not null: the records to join are list[0 .. n_list) (local ordinals), not 0 .. n
    uint32_t n_list = 0;
    uint32_t *counters;
    RgTable rg;
    KeyLayout kl;
    int verify_names;
};

int launch_mate_join(const JoinParams &P, cudaStream_t stream, uint64_t *launches);
// windowed form: pairs settled inside the CTAs' contiguous record ranges; the rest goes on J.left
int launch_local_join(const JoinParams &P, LocalJoinParams J, int sms, cudaStream_t stream, uint64_t *launches,
                      ::oge_gpu_dedup_ctx *timing = nullptr /* profile_events: brackets the two launches */);
// fused form: is the key hash of a pair formed inside a CTA among the records the global join has seen?  Then the pair
// is retracted and its two records (with whatever the slot held) go to the exact path.
int launch_pair_check(const JoinParams &P, const uint64_t *pair_hk, uint32_t n_pairs, bool far, cudaStream_t stream, uint64_t *launches);
int launch_mate_fixup(const JoinParams &P, uint32_t n_slots_listed, cudaStream_t stream, uint64_t *launches);
// exact path over the sorted complex list (sorted by hash then ordinal); state = n_cplx bytes of scratch
int launch_mate_complex(const JoinParams &P, const E128 *sorted_cplx, uint32_t n_cplx, uint8_t *state, void *long_segs, cudaStream_t stream,
                        uint64_t *launches);
size_t mate_complex_long_segs_bytes(uint32_t n_cplx);      // room for the list of long segments

// ---- K4 group-and-select (select.cu) ----------------------------------------------------------
struct SelectParams {
    const E128 *sorted;
    uint32_t n_max;
    const uint32_t *n_dev;      // device count (pairs) or nullptr
    uint8_t *dup;               // [n records] duplicate set, indexed by local ordinal
    const uint32_t *mate_of;    // global ordinal of idx2, indexed by idx1's local ordinal
    uint64_t idx_base;          // global ordinal of local record 0
    uint64_t n_records;         // local records: ordinals outside [idx_base, idx_base + n_records) belong to other ranks
    uint32_t *counters;
    KeyLayout kl;
    // range sharding only (null / 0 on one GPU)
    const uint64_t *fm;         // (idx1 << 32 | idx2) sorted by idx1: mates of pair entries whose idx1 is not local
    uint32_t n_fm;
    uint32_t *foreign_marks;    // out: global ordinals to mark on other ranks
    uint32_t foreign_cap;
    uint32_t *foreign_counter;  // how many of them
    const uint64_t *split;      // world - 1 packed first keys of ranks 1.. (fragment ends whose key another rank owns are left alone)
    int world, rank;
};

int launch_select_pairs(const SelectParams &P, bool far, cudaStream_t stream, uint64_t *launches);
int launch_select_frags(const SelectParams &P, cudaStream_t stream, uint64_t *launches);

// ---- reduced fragment pass (fragfilter.cu): only unpaired ends and the paired ends sharing a key with one matter
int launch_ff_collect(const E128 *frag, uint64_t n, const KeyLayout &L, E128 *out, uint32_t cap, uint32_t *counters, cudaStream_t s,
                      uint64_t *launches);
int launch_ff_set_build(const E128 *list, const uint32_t *n_dev, uint32_t n_max, const KeyLayout &L, unsigned long long *set, uint64_t n_slots,
                        cudaStream_t s, uint64_t *launches);
int launch_ff_filter(const E128 *frag, uint64_t n, const KeyLayout &L, const unsigned long long *set, uint64_t n_slots, E128 *out, uint32_t cap,
                     uint32_t *counters, const uint32_t *n_set /* device: entries in the set, or null */, cudaStream_t s, uint64_t *launches);

// ---- K5 flag write (flags.cu) -----------------------------------------------------------------
struct FlagParams {
    uint8_t *rec;
    const uint64_t *off;
    uint64_t n;
    const uint16_t *flag_in;
    uint16_t *flag_out;
    const uint8_t *dup;
    uint32_t *counters;
    int quiet_index_bug;        // compat F1: the duplicate set collapses to {0}
};

int launch_flags(const FlagParams &P, cudaStream_t stream, uint64_t *launches);

// ---- BGZF inflate (bgzf_inflate.cu): the hardware decompress engine, or one of two SIMT decoders
struct BgzfParams {
    const uint8_t *comp;        // the compressed file
    const uint64_t *in_off;     // [n_blocks] byte offset of every block in comp
    const uint32_t *csize;      // [n_blocks] BSIZE + 1
    const uint64_t *out_off;    // [n_blocks + 1] byte offset of every block's payload in the inflated stream
    uint64_t n_blocks;
    uint64_t block_base;        // index of block 0 of this launch in the file (error reports)
    uint8_t *out;               // where byte 0 of the inflated stream goes
    uint32_t *err;              // [2]: first inflate error code, block index
};
enum { INFLATE_THREADS = 0, INFLATE_WARP = 1, INFLATE_ENGINE = 2 };
int launch_bgzf_inflate(const BgzfParams &P, int mode, int sms, cudaStream_t stream, uint64_t *launches);
void set_inflate_kernel(int mode);      // INFLATE_*; -1: back to the default (OGE_INFLATE_KERNEL, else the engine where the device has one)
int inflate_mode_for(int device);       // the mode push_bgzf uses on this device
// Hardware decompress engine (cuMemBatchDecompressAsync): one operation per block of [b0, b1), host-side block table;
// the engine writes the byte count of every block to act[b].  Blocks that inflate to nothing are not submitted.
int engine_inflate_submit(const uint8_t *d_comp, const uint64_t *h_in_off, const uint32_t *h_csize, const uint32_t *h_isize,
                          const uint64_t *h_out_off, uint8_t *d_out, uint32_t *d_act, uint64_t b0, uint64_t b1, cudaStream_t stream,
                          uint64_t *submitted);
// act[b] against ISIZE for every block (the reference's assert(zs.total_out == uncompressed_size)) -> err[0..1]
int launch_bgzf_check_sizes(const uint32_t *act, const uint64_t *out_off, uint64_t n_blocks, uint32_t *err, cudaStream_t stream, uint64_t *launches);
uint64_t engine_inflate_slack(int device);      // bytes the destination buffer keeps free behind the last block

// ---- BGZF members made on the device (bgzf_deflate.cu): bins, one warp per block deflate + CRC-32, packing
constexpr uint32_t DEFLATE_PAYLOAD = 65280;          // inflated bytes per block (any split is a valid BGZF file; htslib's)
constexpr uint32_t DEFLATE_STAGE_STRIDE = 65536;     // staging slot per block: the stream is at most payload + 5 bytes (+ 3 of word padding)
struct DeflateParams {
    const uint8_t *in;          // the byte stream to compress, 4-byte aligned, readable 8 bytes beyond its end
    uint64_t total;
    uint32_t payload;
    uint64_t n_blocks;
    uint8_t *stage;             // [n_blocks][DEFLATE_STAGE_STRIDE]
    uint32_t *dsize;            // [n_blocks] bytes of every block's deflate stream
    uint32_t *crc;              // [n_blocks] CRC-32 of every block's payload
    void *seqs;                 // parse scratch, deflate_seq_bytes(sms)
    unsigned long long *ticket; // block dispenser
};
size_t deflate_seq_bytes(int sms);
int launch_fix_bins(uint8_t *rec, const uint64_t *off, uint64_t n, uint32_t *err, cudaStream_t stream, uint64_t *launches);
int launch_bgzf_deflate(const DeflateParams &P, int sms, cudaStream_t stream, uint64_t *launches);
int launch_scan_sizes(const uint32_t *dsize, uint64_t n, uint64_t *moff /* n + 1 */, cudaStream_t stream, uint64_t *launches);
int launch_bgzf_assemble(const DeflateParams &P, uint64_t *moff, uint8_t *out, cudaStream_t stream, uint64_t *launches);

// ---- record framing on the device (frame.cu): speculative parallel chain walk + proof
struct FrameParams {
    const uint8_t *rec;         // first record
    uint64_t total;             // record bytes
    uint64_t chunk;             // bytes per chunk
    uint64_t n_chunks;
    int32_t n_ref;
    uint64_t *entry;            // [n_chunks] first record start at or after the chunk start (guess, then proven)
    uint64_t *exit_;            // [n_chunks] first record start at or after the chunk end, walking from entry
    uint64_t *count;            // [n_chunks] records starting inside the chunk
    uint32_t *bad;              // [n_chunks] 0, 1 = truncated, 2 | block_size << 2 = invalid block size
    const uint64_t *base;       // [n_chunks] rank of the chunk's first record (write pass)
    uint64_t *off;              // offsets out
};
int launch_frame_guess(const FrameParams &P, cudaStream_t s, uint64_t *launches);
int launch_frame_walk(const FrameParams &P, uint64_t first_chunk, uint64_t n_walk, int write, cudaStream_t s, uint64_t *launches);

// ---- flag statistics (flagstat.cu): Statistics::runInternal's counters over the resident records
enum {
    FS_READS = 0, FS_MAPPED, FS_FORWARD, FS_REVERSE, FS_FAILED_QC, FS_DUPLICATES, FS_PAIRED, FS_PROPER_PAIR,
    FS_BOTH_MAPPED, FS_FIRST_MATE, FS_SECOND_MATE, FS_SINGLETONS,
    FS_N_COUNTERS,
    FS_SORTED = FS_N_COUNTERS,
    FS_N_OUT
};
struct SortTile {      // what a 1024-record tile tells the seam check about its considered records
    int32_t n_valid, bad;      // 0, 1, 2 (= two or more)
    int32_t first_ref, first_pos, second_ref, second_pos, last2_ref, last2_pos, last_ref, last_pos;
};
size_t flagstat_scratch_bytes(uint64_t n);
int launch_flagstats(const uint8_t *rec, const uint64_t *off, const uint16_t *flags, uint64_t n, void *scratch, int sms,
                     cudaStream_t stream, uint64_t *launches);

// pull with remove_duplicates: compaction of the kept records
// ---- range sharding (shard.cu) ------------------------------------------------------------------
// What leaves a rank about one record whose name was not seen exactly twice locally.
struct __align__(16) PubEntry {
    E128 frag;          // its fragment end entry (global ordinal inside)
    uint64_t hk;        // pairing-key hash
    uint64_t rsv;
    NameTag tag;        // read-group code + name, for the exact comparison on the receiving side
};
static_assert(sizeof(PubEntry) == 64, "PubEntry is 64 bytes on the wire");

// An end entry handed to the rank that owns its key range.
struct __align__(16) RouteEntry {
    E128 e;
    uint32_t idx2;      // pairs: global ordinal of the second record
    uint32_t kind;      // 0 fragment, 1 near pair, 2 far pair
    uint64_t rsv;
};
static_assert(sizeof(RouteEntry) == 32, "RouteEntry is 32 bytes on the wire");

struct ShardParams {
    const uint64_t *split;      // world - 1 packed keys (ref << coord_bits | biased coord): first key of ranks 1..
    int world, rank;
    uint64_t idx_base, n;       // local ordinal range
    KeyLayout kl;
    uint32_t *counters;
};

int launch_compact(const uint8_t *rec, const uint64_t *off, uint64_t n, const uint16_t *flag_out, int remove_dups,
                   uint8_t *out_rec, uint64_t *out_off, uint64_t *scratch /* n+1 + blocks */, uint32_t *counters,
                   cudaStream_t stream, uint64_t *launches);
size_t compact_scratch_bytes(uint64_t n);

}  // namespace oge
