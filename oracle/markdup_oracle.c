/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's duplicate-marking path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * The product path (openge_b200/csrc) never calls it and has no CPU fallback.
 *
 * Parity status: PINNED against the compiled reference itself (oracle/_ref/oge_ref_dedup,
 * built in place from /root/reference by oracle/ref_build/Makefile): tests/test_oracle.py
 * checks this restatement against the reference's flags on test/data/208.yhet.bam
 * (6642 of 15419 flagged), on the SURVEY A.3 known-answer fixtures and on seeded synthetic
 * BAMs.  The reference's own tests hold no golden vector for dedup (test/CMakeLists.txt:32
 * is exit-code only).
 *
 * Plain C, single threaded, sequential -- it follows the reference line by line rather
 * than the GPU design.  Citations are into /root/reference/openge/src/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* One 5' end, as util/picard_structures.h:29-54 (ReadEnds). */
typedef struct {
    int16_t libraryId;
    int16_t score;
    int32_t orientation;       /* RE_NONE=0,F=1,R=2,FF=3,RR=4,FR=5,RF=6  picard_structures.h:24-27 */
    int32_t read1Sequence;
    int32_t read1Coordinate;
    int64_t read1IndexInFile;
    int32_t read2Sequence;
    int32_t read2Coordinate;
    int64_t read2IndexInFile;
} ReadEnds;

/* Per-record view of the end-building step, exported for field-by-field parity of K1. */
typedef struct {
    int32_t eligible;          /* mapped && refID != -1 && primary          mark_duplicates.cpp:202-205 */
    int32_t pair_eligible;     /* eligible && paired && mate mapped         mark_duplicates.cpp:209 */
    int32_t ref;               /* read1Sequence */
    int32_t coord;             /* read1Coordinate (unclipped 5' end) */
    int32_t orientation;       /* RE_F / RE_R */
    int32_t read2Sequence;     /* mate refID when paired && mate mapped, else -1 */
    int16_t score;
    int16_t lib;
} OracleEnd;

enum { RE_NONE, RE_F, RE_R, RE_FF, RE_RR, RE_FR, RE_RF };

static int32_t rd_i32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
static uint32_t rd_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd_u16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

/* Decoded view of one raw BAM record (util/bam_deserializer.h:144-193). */
typedef struct {
    const uint8_t *base;       /* points at block_size */
    uint32_t block_size;
    int32_t ref_id, pos, mate_ref, mate_pos;
    uint32_t l_read_name, n_cigar, l_seq;
    uint16_t flag;
    const uint8_t *name, *cigar, *qual, *tags;
    uint32_t tags_len;
} Rec;

static void decode(const uint8_t *p, Rec *r) {
    r->base = p;
    r->block_size = rd_u32(p);
    r->ref_id = rd_i32(p + 4);
    r->pos = rd_i32(p + 8);
    r->l_read_name = p[12];
    r->n_cigar = rd_u16(p + 16);
    r->flag = rd_u16(p + 18);
    r->l_seq = rd_u32(p + 20);
    r->mate_ref = rd_i32(p + 24);
    r->mate_pos = rd_i32(p + 28);
    r->name = p + 36;
    r->cigar = r->name + r->l_read_name;
    r->qual = r->cigar + 4u * r->n_cigar + (r->l_seq + 1) / 2;
    r->tags = r->qual + r->l_seq;
    {
        int64_t used = (int64_t)(r->tags - (p + 4));
        int64_t left = (int64_t) r->block_size - used;
        r->tags_len = left > 0 ? (uint32_t) left : 0;
    }
}

/* bamtools/BamConstants.h:48: op index into "MIDNSHP=X" */
static int cig_op(const Rec *r, uint32_t i) { return (int)(rd_u32(r->cigar + 4 * i) & 0xf); }
static int32_t cig_len(const Rec *r, uint32_t i) { return (int32_t)(rd_u32(r->cigar + 4 * i) >> 4); }

/* mark_duplicates.cpp:44-61  getReferenceLength: sum of M D N = X */
static int32_t reference_length(const Rec *r) {
    int32_t len = 0;
    for (uint32_t i = 0; i < r->n_cigar; i++) {
        int op = cig_op(r, i);
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) len += cig_len(r, i);
    }
    return len;
}

/* mark_duplicates.cpp:88-103  getUnclippedStart */
static int32_t unclipped_start(const Rec *r) {
    int32_t pos = r->pos;
    for (uint32_t i = 0; i < r->n_cigar; i++) {
        int op = cig_op(r, i);
        if (op == 4 || op == 5) pos -= cig_len(r, i); else break;
    }
    return pos;
}

/* mark_duplicates.cpp:73-79,112-129  getAlignmentEnd + getUnclippedEnd */
static int32_t unclipped_end(const Rec *r) {
    int32_t pos = (r->flag & 0x4) ? -1 : r->pos + reference_length(r) - 1;
    for (int64_t i = (int64_t) r->n_cigar - 1; i >= 0; i--) {
        int op = cig_op(r, (uint32_t) i);
        if (op == 4 || op == 5) pos += cig_len(r, (uint32_t) i); else break;
    }
    return pos;
}

/* mark_duplicates.cpp:135-144 + bamtools/BamAlignment.cpp:846-851: short accumulator,
 * every raw quality byte b (uint8) with b >= 15 is added; wraps mod 2^16. */
static int16_t score_of(const Rec *r) {
    int16_t score = 0;
    for (uint32_t i = 0; i < r->l_seq; i++) {
        uint8_t b = r->qual[i];
        if (b >= 15) score = (int16_t)(score + b);
    }
    return score;
}

/* bamtools/BamAlignment.cpp:270-294 FindTag + :699-786 SkipToNextTag, for tag "RG".
 * Returns pointer to the value (any type code: GetTag<string> never checks it,
 * BamAlignment.h:575-606) and its strlen bounded by the record end. */
static const uint8_t *find_rg(const Rec *r, uint32_t *len) {
    const uint8_t *p = r->tags;
    uint32_t n = r->tags_len, parsed = 0;
    *len = 0;
    if (n == 0) return NULL;
    while (parsed < n) {
        const uint8_t *name = p;
        uint8_t type;
        if (n - parsed < 3) return NULL;      /* truncated tag header: the reference reads past the end here */
        type = p[2];
        p += 3; parsed += 3;
        if (name[0] == 'R' && name[1] == 'G') {
            uint32_t l = 0;
            while (parsed + l < n && p[l]) l++;
            *len = l;
            return p;
        }
        if (type == 0) return NULL;
        switch (type) {
            case 'A': case 'c': case 'C': p += 1; parsed += 1; break;
            case 's': case 'S': p += 2; parsed += 2; break;
            case 'f': case 'i': case 'I': p += 4; parsed += 4; break;
            case 'Z': case 'H':
                while (parsed < n && *p) { p++; parsed++; }
                p++; parsed++;
                break;
            case 'B': {
                uint8_t at;
                int32_t cnt;
                int64_t skip;
                if (parsed + 5 > n) return NULL;
                at = *p; p++; parsed++;
                cnt = rd_i32(p); p += 4; parsed += 4;
                if (at == 'c' || at == 'C') skip = cnt;
                else if (at == 's' || at == 'S') skip = 2 * (int64_t) cnt;
                else if (at == 'f' || at == 'i' || at == 'I') skip = 4 * (int64_t) cnt;
                else return NULL;
                if (skip < 0 || parsed + skip > n) return NULL;
                p += skip; parsed += (uint32_t) skip;
                break;
            }
            default: return NULL;
        }
        if (parsed >= n) return NULL;
        if (*p == 0) return NULL;
    }
    return NULL;
}

/* mark_duplicates.cpp:282-318 getLibraryId/getLibraryName with the host-resolved table
 * (rg id -> library id; see openge_b200/header.py for the @RG parsing rules). */
typedef struct {
    const char *const *rg_ids;
    const int16_t *rg_lib;
    int32_t n_rg;
    int16_t unknown_lib;
} LibTable;

static int16_t library_id(const LibTable *t, const uint8_t *rg, uint32_t rg_len) {
    if (rg && rg_len > 0) {
        for (int32_t i = 0; i < t->n_rg; i++) {
            if (strlen(t->rg_ids[i]) == rg_len && memcmp(t->rg_ids[i], rg, rg_len) == 0)
                return t->rg_lib[i];
        }
    }
    return t->unknown_lib;
}

/* mark_duplicates.cpp:147-164 buildReadEnds */
static void build_read_ends(const LibTable *t, int64_t index, const Rec *r, ReadEnds *e) {
    uint32_t rg_len;
    const uint8_t *rg = find_rg(r, &rg_len);
    int rev = (r->flag & 0x10) != 0;
    e->libraryId = -1; e->score = -1; e->orientation = RE_NONE;
    e->read1Sequence = -1; e->read1Coordinate = -1; e->read1IndexInFile = -1;
    e->read2Sequence = -1; e->read2Coordinate = -1; e->read2IndexInFile = -1;
    e->read1Sequence = r->ref_id;
    e->read1Coordinate = rev ? unclipped_end(r) : unclipped_start(r);
    e->orientation = rev ? RE_R : RE_F;
    e->read1IndexInFile = index;
    e->score = score_of(r);
    if ((r->flag & 0x1) && !(r->flag & 0x8)) e->read2Sequence = r->mate_ref;
    e->libraryId = library_id(t, rg, rg_len);
}

/* mark_duplicates.cpp:169-178 getOrientationByte */
static int orientation_byte(int neg1, int neg2) {
    if (neg1) return neg2 ? RE_RR : RE_RF;
    return neg2 ? RE_FR : RE_FF;
}

/* picard_structures.h:56-68 ReadEnds::compare (exact three-way compares in place of the
 * reference's int subtraction; identical wherever the subtraction does not overflow). */
#define CMP(a, b) do { if ((a) < (b)) return -1; if ((a) > (b)) return 1; } while (0)
static int ends_compare(const void *pa, const void *pb) {
    const ReadEnds *l = *(const ReadEnds *const *) pa, *r = *(const ReadEnds *const *) pb;
    CMP(l->libraryId, r->libraryId);
    CMP(l->read1Sequence, r->read1Sequence);
    CMP(l->read1Coordinate, r->read1Coordinate);
    CMP(l->orientation, r->orientation);
    CMP(l->read2Sequence, r->read2Sequence);
    CMP(l->read2Coordinate, r->read2Coordinate);
    CMP(l->read1IndexInFile, r->read1IndexInFile);
    CMP(l->read2IndexInFile, r->read2IndexInFile);
    return 0;
}

/* ReadEndsMap (picard_structures.h:82-109): exact string-keyed map, key = RG + ":" + name
 * (mark_duplicates.cpp:210-214).  Open addressing with tombstone-free backward-shift delete;
 * equality is on the full key bytes, the hash only picks the probe start. */
typedef struct { const uint8_t *rg; uint32_t rg_len; const uint8_t *name; uint32_t name_len; ReadEnds *val; } Slot;
typedef struct { Slot *s; uint64_t cap, used; } Map;

static uint64_t key_hash(const uint8_t *rg, uint32_t rg_len, const uint8_t *name, uint32_t name_len) {
    uint64_t h = 1469598103934665603ULL;
    uint32_t i;
    for (i = 0; i < rg_len; i++) { h ^= rg[i]; h *= 1099511628211ULL; }
    h ^= ':'; h *= 1099511628211ULL;
    for (i = 0; i < name_len; i++) { h ^= name[i]; h *= 1099511628211ULL; }
    return h ^ (h >> 29);
}

/* Compare the logical concatenations RG + ":" + name byte by byte. */
static int key_equal(const Slot *s, const uint8_t *rg, uint32_t rg_len, const uint8_t *name, uint32_t name_len) {
    uint32_t la = s->rg_len + 1 + s->name_len, lb = rg_len + 1 + name_len, i;
    if (la != lb) return 0;
    for (i = 0; i < la; i++) {
        uint8_t a = i < s->rg_len ? s->rg[i] : (i == s->rg_len ? ':' : s->name[i - s->rg_len - 1]);
        uint8_t b = i < rg_len ? rg[i] : (i == rg_len ? ':' : name[i - rg_len - 1]);
        if (a != b) return 0;
    }
    return 1;
}

static void map_init(Map *m, uint64_t cap) {
    m->cap = 1024; while (m->cap < cap) m->cap <<= 1;
    m->s = (Slot *) calloc(m->cap, sizeof(Slot));
    m->used = 0;
}

static void map_grow(Map *m);

static Slot *map_find(Map *m, const uint8_t *rg, uint32_t rg_len, const uint8_t *name, uint32_t name_len) {
    uint64_t i = key_hash(rg, rg_len, name, name_len) & (m->cap - 1);
    while (m->s[i].val) {
        if (key_equal(&m->s[i], rg, rg_len, name, name_len)) return &m->s[i];
        i = (i + 1) & (m->cap - 1);
    }
    return &m->s[i];
}

static void map_put(Map *m, const uint8_t *rg, uint32_t rg_len, const uint8_t *name, uint32_t name_len, ReadEnds *v) {
    Slot *s;
    if ((m->used + 1) * 2 > m->cap) map_grow(m);
    s = map_find(m, rg, rg_len, name, name_len);
    if (!s->val) m->used++;
    s->rg = rg; s->rg_len = rg_len; s->name = name; s->name_len = name_len; s->val = v;
}

static void map_grow(Map *m) {
    Slot *old = m->s; uint64_t oc = m->cap, i;
    m->cap <<= 1; m->s = (Slot *) calloc(m->cap, sizeof(Slot)); m->used = 0;
    for (i = 0; i < oc; i++) if (old[i].val) map_put(m, old[i].rg, old[i].rg_len, old[i].name, old[i].name_len, old[i].val);
    free(old);
}

static void map_erase(Map *m, Slot *s) {
    uint64_t mask = m->cap - 1, i = (uint64_t)(s - m->s), j = i;
    m->s[i].val = NULL; m->used--;
    for (;;) {
        uint64_t k;
        j = (j + 1) & mask;
        if (!m->s[j].val) break;
        k = key_hash(m->s[j].rg, m->s[j].rg_len, m->s[j].name, m->s[j].name_len) & mask;
        if ((i <= j) ? (i < k && k <= j) : (i < k || k <= j)) continue;
        m->s[i] = m->s[j]; m->s[j].val = NULL; i = j;
    }
}

typedef struct { ReadEnds **v; uint64_t n, cap; } Vec;
static void vec_push(Vec *v, ReadEnds *e) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->v = (ReadEnds **) realloc(v->v, v->cap * sizeof(*v->v)); }
    v->v[v->n++] = e;
}

/* mark_duplicates.cpp:402-414 */
static int comparable(const ReadEnds *l, const ReadEnds *r, int compare_read2) {
    int ret = l->libraryId == r->libraryId && l->read1Sequence == r->read1Sequence &&
              l->read1Coordinate == r->read1Coordinate && l->orientation == r->orientation;
    if (ret && compare_read2) ret = l->read2Sequence == r->read2Sequence && l->read2Coordinate == r->read2Coordinate;
    return ret;
}

/* mark_duplicates.cpp:477-480: std::set<int> insert (long -> int truncation). */
static void add_dup(uint8_t *dup, uint64_t n, int64_t idx, uint64_t *calls) {
    int32_t t = (int32_t) idx;
    (*calls)++;
    if (t >= 0 && (uint64_t) t < n) dup[t] = 1;
}

/* mark_duplicates.cpp:488-507 */
static void mark_pairs(ReadEnds **list, uint64_t cnt, uint8_t *dup, uint64_t n, uint64_t *calls) {
    int16_t max_score = 0; ReadEnds *best = NULL; uint64_t i;
    for (i = 0; i < cnt; i++) if (list[i]->score > max_score || best == NULL) { max_score = list[i]->score; best = list[i]; }
    for (i = 0; i < cnt; i++) if (list[i] != best) { add_dup(dup, n, list[i]->read1IndexInFile, calls); add_dup(dup, n, list[i]->read2IndexInFile, calls); }
}

/* mark_duplicates.cpp:515-540 */
static void mark_frags(ReadEnds **list, uint64_t cnt, int contains_pairs, uint8_t *dup, uint64_t n, uint64_t *calls) {
    uint64_t i;
    if (contains_pairs) {
        for (i = 0; i < cnt; i++) if (list[i]->read2Sequence == -1) add_dup(dup, n, list[i]->read1IndexInFile, calls);
    } else {
        int16_t max_score = 0; ReadEnds *best = NULL;
        for (i = 0; i < cnt; i++) if (list[i]->score > max_score || best == NULL) { max_score = list[i]->score; best = list[i]; }
        for (i = 0; i < cnt; i++) if (list[i] != best) add_dup(dup, n, list[i]->read1IndexInFile, calls);
    }
}

/*
 * The whole path: MarkDuplicates::runInternal (mark_duplicates.cpp:422-475).
 *   records/offsets : raw BAM records back to back, offsets[n+1]
 *   compat_quiet    : reproduce the reference's non-verbose behaviour (index never
 *                     increments, mark_duplicates.cpp:250).  0 = canonical (-v).
 *   flags_out       : n u16 flag words after the rewrite loop (:443-465)
 *   ends_out        : optional, n OracleEnd
 *   stats_out       : optional [4]: frag entries, pair entries, addIndexAsDuplicate calls, unmatched
 * returns 0, or -1 on allocation failure.
 */

/* ---------------------------------------------------------------------------------------------
 * Header side of getLibraryName / getLibraryId (mark_duplicates.cpp:282-318), restated from the
 * reference's own header model so that the oracle does not lean on the product's header code:
 *   BamHeader::BamHeader(text)        util/bam_header.cpp:107-180: getline() until the stream is no
 *                                     longer good -- a final line without '\n' is never seen;
 *                                     "@RG\t" lines become BamReadGroupRecords in file order
 *   BamReadGroupRecord(line)          util/bam_header.cpp:81-106: tab-split segments, tag = first two
 *                                     characters, data from the fourth; a later ID / LB on the line
 *                                     overwrites an earlier one
 *   BamReadGroupRecords::operator[]   util/bam_header.h:216-241: the FIRST record with the ID
 *   library ids                       handed out per distinct library NAME (:282-294); an RG without
 *                                     LB, an unknown ID or no RG tag give "Unknown Library" (:301-318)
 * out: ids[i] points into `text` (length id_len[i]); lib[i] = 1-based id of that read group's library
 * name; *unknown = id of "Unknown Library".  Returns the number of read groups, or -1.
 */
#define ORACLE_MAX_RG 4096
static int parse_read_groups(const char *text, const char **ids, uint32_t *id_len, int16_t *lib, int16_t *unknown) {
    static const char UNK[] = "Unknown Library";
    const char *names[ORACLE_MAX_RG + 1];
    uint32_t name_len[ORACLE_MAX_RG + 1];
    int n_names = 0, n_rg = 0;
    const char *p = text;
    names[0] = UNK; name_len[0] = (uint32_t) strlen(UNK); n_names = 1;
    *unknown = 1;
    while (*p) {
        const char *eol = strchr(p, '\n');
        if (!eol) break;                                   /* unterminated last line: dropped (:111-116) */
        if (eol - p >= 4 && p[0] == '@' && p[1] == 'R' && p[2] == 'G' && p[3] == '\t') {
            const char *seg = p + 4, *id = NULL, *lb = NULL;
            uint32_t idl = 0, lbl = 0;
            while (seg <= eol) {
                const char *end = memchr(seg, '\t', (size_t)(eol - seg));
                if (!end) end = eol;
                if (end - seg >= 3) {
                    if (seg[0] == 'I' && seg[1] == 'D') { id = seg + 3; idl = (uint32_t)(end - seg - 3); }
                    else if (seg[0] == 'L' && seg[1] == 'B') { lb = seg + 3; lbl = (uint32_t)(end - seg - 3); }
                } else if (end - seg == 2) {               /* "ID" with no data: substr(3) would throw; treat as empty */
                    if (seg[0] == 'I' && seg[1] == 'D') { id = seg + 2; idl = 0; }
                    else if (seg[0] == 'L' && seg[1] == 'B') { lb = seg + 2; lbl = 0; }
                }
                if (end == eol) break;
                seg = end + 1;
            }
            if (id && n_rg < ORACLE_MAX_RG) {
                int dup = 0, i, k = -1;
                for (i = 0; i < n_rg; i++)
                    if (id_len[i] == idl && !memcmp(ids[i], id, idl)) { dup = 1; break; }   /* first record with the ID wins */
                if (!dup) {
                    if (!lb || lbl == 0) k = 0;
                    else {
                        for (i = 0; i < n_names; i++)
                            if (name_len[i] == lbl && !memcmp(names[i], lb, lbl)) { k = i; break; }
                        if (k < 0) { names[n_names] = lb; name_len[n_names] = lbl; k = n_names++; }
                    }
                    ids[n_rg] = id; id_len[n_rg] = idl; lib[n_rg] = (int16_t)(k + 1);
                    n_rg++;
                }
            }
        }
        p = eol + 1;
    }
    return n_rg;
}

/* The whole path from the header TEXT: what tests and bench use. */
int oge_oracle_markdup_text(const uint8_t *records, const uint64_t *offsets, uint64_t n, const char *header_text,
                            int compat_quiet, uint16_t *flags_out, OracleEnd *ends_out, uint64_t *stats_out);

int oge_oracle_markdup(const uint8_t *records, const uint64_t *offsets, uint64_t n,
                       const char *const *rg_ids, const int16_t *rg_lib, int32_t n_rg, int16_t unknown_lib,
                       int compat_quiet, uint16_t *flags_out, OracleEnd *ends_out, uint64_t *stats_out) {
    LibTable lt;
    Map tmp;
    Vec pair_sort = {0, 0, 0}, frag_sort = {0, 0, 0};
    uint8_t *dup = (uint8_t *) calloc(n ? n : 1, 1);
    uint64_t calls = 0, i;
    int64_t index = 0;
    if (!dup) return -1;
    lt.rg_ids = rg_ids; lt.rg_lib = rg_lib; lt.n_rg = n_rg; lt.unknown_lib = unknown_lib;
    map_init(&tmp, 1 << 16);

    /* buildSortedReadEndLists, mark_duplicates.cpp:185-279 */
    for (i = 0; i < n; i++) {
        Rec r;
        decode(records + offsets[i], &r);
        if (ends_out) memset(&ends_out[i], 0, sizeof(OracleEnd));
        if ((r.flag & 0x4) || r.ref_id == -1) {
            /* unmapped / no coordinate: passes through (:202-204) */
        } else if (!(r.flag & 0x100)) {
            ReadEnds *frag = (ReadEnds *) malloc(sizeof(ReadEnds));
            build_read_ends(&lt, index, &r, frag);
            vec_push(&frag_sort, frag);
            if (ends_out) {
                OracleEnd *o = &ends_out[i];
                o->eligible = 1; o->ref = frag->read1Sequence; o->coord = frag->read1Coordinate;
                o->orientation = frag->orientation; o->read2Sequence = frag->read2Sequence;
                o->score = frag->score; o->lib = frag->libraryId;
                o->pair_eligible = (r.flag & 0x1) && !(r.flag & 0x8);
            }
            if ((r.flag & 0x1) && !(r.flag & 0x8)) {
                uint32_t rg_len; const uint8_t *rg = find_rg(&r, &rg_len);
                uint32_t name_len = r.l_read_name ? r.l_read_name - 1 : 0;
                Slot *s = map_find(&tmp, rg, rg_len, r.name, name_len);
                ReadEnds *paired = s->val;
                if (paired) map_erase(&tmp, s);
                if (paired == NULL) {
                    paired = (ReadEnds *) malloc(sizeof(ReadEnds));
                    build_read_ends(&lt, index, &r, paired);
                    map_put(&tmp, rg, rg_len, r.name, name_len, paired);
                } else {
                    int32_t sequence = frag->read1Sequence, coordinate = frag->read1Coordinate;
                    int rev = (r.flag & 0x10) != 0;
                    if (sequence > paired->read1Sequence ||
                        (sequence == paired->read1Sequence && coordinate >= paired->read1Coordinate)) {
                        paired->read2Sequence = sequence;
                        paired->read2Coordinate = coordinate;
                        paired->read2IndexInFile = index;
                        paired->orientation = orientation_byte(paired->orientation == RE_R, rev);
                    } else {
                        paired->read2Sequence = paired->read1Sequence;
                        paired->read2Coordinate = paired->read1Coordinate;
                        paired->read2IndexInFile = paired->read1IndexInFile;
                        paired->read1Sequence = sequence;
                        paired->read1Coordinate = coordinate;
                        paired->read1IndexInFile = index;
                        paired->orientation = orientation_byte(rev, paired->orientation == RE_R);
                    }
                    paired->score = (int16_t)(paired->score + score_of(&r));
                    vec_push(&pair_sort, paired);
                }
            }
        }
        if (!compat_quiet) ++index;   /* :250 -- only advances when verbose */
    }

    if (pair_sort.n) qsort(pair_sort.v, pair_sort.n, sizeof(ReadEnds *), ends_compare);   /* :262-265 */
    if (frag_sort.n) qsort(frag_sort.v, frag_sort.n, sizeof(ReadEnds *), ends_compare);   /* :267-271 */

    if (stats_out) { stats_out[0] = frag_sort.n; stats_out[1] = pair_sort.n; stats_out[3] = tmp.used; }

    /* generateDuplicateIndexes, mark_duplicates.cpp:326-400 */
    {
        ReadEnds *first = NULL;
        uint64_t chunk_start = 0, chunk_n = 0;
        int contains_pairs = 0, contains_frags = 0;
        for (i = 0; i < pair_sort.n; i++) {
            ReadEnds *next = pair_sort.v[i];
            if (first == NULL) { first = next; chunk_start = i; chunk_n = 1; }
            else if (comparable(first, next, 1)) chunk_n++;
            else {
                if (chunk_n > 1) mark_pairs(pair_sort.v + chunk_start, chunk_n, dup, n, &calls);
                chunk_start = i; chunk_n = 1; first = next;
            }
        }
        mark_pairs(pair_sort.v + chunk_start, chunk_n, dup, n, &calls);

        first = NULL; chunk_start = 0; chunk_n = 0;
        for (i = 0; i < frag_sort.n; i++) {
            ReadEnds *next = frag_sort.v[i];
            if (first != NULL && comparable(first, next, 0)) {
                chunk_n++;
                contains_pairs = contains_pairs || next->read2Sequence != -1;
                contains_frags = contains_frags || next->read2Sequence == -1;
            } else {
                if (chunk_n > 1 && contains_frags) mark_frags(frag_sort.v + chunk_start, chunk_n, contains_pairs, dup, n, &calls);
                chunk_start = i; chunk_n = 1; first = next;
                contains_pairs = next->read2Sequence != -1;
                contains_frags = next->read2Sequence == -1;
            }
        }
        mark_frags(frag_sort.v + chunk_start, chunk_n, contains_pairs, dup, n, &calls);
    }
    if (stats_out) stats_out[2] = calls;

    /* rewrite loop, mark_duplicates.cpp:443-465 + BamAlignment.cpp:600-603 */
    for (i = 0; i < n; i++) {
        uint16_t flag = rd_u16(records + offsets[i] + 18);
        if (!(flag & 0x100)) flag = dup[i] ? (uint16_t)(flag | 0x400) : (uint16_t)(flag & ~0x400);
        flags_out[i] = flag;
    }

    for (i = 0; i < pair_sort.n; i++) free(pair_sort.v[i]);
    for (i = 0; i < frag_sort.n; i++) free(frag_sort.v[i]);
    for (i = 0; i < tmp.cap; i++) if (tmp.s[i].val) free(tmp.s[i].val);   /* :274-278 */
    free(pair_sort.v); free(frag_sort.v); free(tmp.s); free(dup);
    return 0;
}

int oge_oracle_markdup_text(const uint8_t *records, const uint64_t *offsets, uint64_t n, const char *header_text,
                            int compat_quiet, uint16_t *flags_out, OracleEnd *ends_out, uint64_t *stats_out) {
    const char *ids[ORACLE_MAX_RG];
    uint32_t id_len[ORACLE_MAX_RG];
    int16_t lib[ORACLE_MAX_RG], unknown = 1;
    char *idz[ORACLE_MAX_RG];
    int n_rg = parse_read_groups(header_text, ids, id_len, lib, &unknown), i, rc;
    if (n_rg < 0) return -1;
    for (i = 0; i < n_rg; i++) {                            /* NUL-terminated copies for the table */
        idz[i] = (char *) malloc(id_len[i] + 1);
        memcpy(idz[i], ids[i], id_len[i]);
        idz[i][id_len[i]] = 0;
    }
    rc = oge_oracle_markdup(records, offsets, n, (const char *const *) idz, lib, n_rg, unknown, compat_quiet, flags_out, ends_out, stats_out);
    for (i = 0; i < n_rg; i++) free(idz[i]);
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * Flag statistics: the counting loop of Statistics::runInternal (algorithms/statistics.cpp:77-162),
 * over framed records.  `flags` (optional) overrides the flag word of the records (the dedup
 * output); out[0..11] = numReads, numMapped, numForwardStrand, numReverseStrand, numFailedQC,
 * numDuplicates, numPaired, numProperPair, numBothMatesMapped, numFirstMate, numSecondMate,
 * numSingletons; out[12] = the "Sorted:" verdict (1 Yes / 0 No).
 * Flag predicates: util/bamtools/BamAlignment.cpp:444-516 (bit tests of BamConstants.h:35-45).
 */
int oge_oracle_flagstats(const uint8_t *records, const uint64_t *offsets, uint64_t n, const uint16_t *flags, uint64_t *out) {
    uint64_t i;
    int sorted = 1, last_rid = -1, last_position = -1;          /* statistics.cpp:82-84 */
    memset(out, 0, 13 * sizeof(uint64_t));
    for (i = 0; i < n; i++) {
        const uint8_t *p = records + offsets[i];
        const int32_t rid = rd_i32(p + 4), pos = rd_i32(p + 8);
        const uint32_t f = flags ? flags[i] : rd_u16(p + 18);
        if (sorted && rid != -1 && pos != -1) {                  /* :89-101 */
            if (last_rid > rid) sorted = 0;
            else if (last_rid < rid) { last_rid = rid; last_position = -1; }
            else {
                if (last_position > pos) sorted = 0;
                else last_position = pos;
            }
        }
        out[0]++;                                                /* :104 */
        if (f & 0x400) out[5]++;                                 /* :107 IsDuplicate */
        if (f & 0x200) out[4]++;                                 /* :108 IsFailedQC */
        if (!(f & 0x4)) out[1]++;                                /* :109 IsMapped */
        if (f & 0x10) out[3]++; else out[2]++;                   /* :112-115 */
        if (f & 0x1) {                                           /* :118 IsPaired */
            out[6]++;
            if (f & 0x40) out[9]++;                              /* :124 */
            if (f & 0x80) out[10]++;                             /* :125 */
            if (!(f & 0x4)) {                                    /* :128-135 */
                if (!(f & 0x8)) out[8]++;
                else out[11]++;
            }
            if (f & 0x2) out[7]++;                               /* :138 */
        }
    }
    out[12] = (uint64_t) sorted;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Coordinate order: the order `openge mergesort` puts records in before MarkDuplicates
 * (algorithms/read_sorter.cpp: runs sorted with Sort::ByPosition, merged through a multiset of the same
 * comparator, util/read_stream_reader.h:88-106).  ByPosition (util/bamtools/Sort.h) compares refID (-1 after
 * everything else), position, strand (forward first), name (std::string order), flag, and finally the ADDRESSES
 * of the two objects; records with refID -1 are all equivalent.  So the reference's order is defined only up to
 * permutations inside groups that tie on the five fields and inside the unplaced tail; this restatement breaks
 * those ties by input order (which is what a stable sort would do) and tests compare accordingly.
 * perm[k] = input ordinal of the record at output position k.
 */
typedef struct {
    const uint8_t *p;      /* record */
    uint32_t idx;
} SortItem;

static int by_position(const void *pa, const void *pb) {
    const SortItem *a = (const SortItem *) pa, *b = (const SortItem *) pb;
    const int32_t ra = rd_i32(a->p + 4), rb = rd_i32(b->p + 4);
    if (ra == -1 || rb == -1) {
        if (ra == -1 && rb == -1) return a->idx < b->idx ? -1 : 1;
        return ra == -1 ? 1 : -1;
    }
    if (ra != rb) return ra < rb ? -1 : 1;
    const int32_t qa = rd_i32(a->p + 8), qb = rd_i32(b->p + 8);
    if (qa != qb) return qa < qb ? -1 : 1;
    const uint32_t fa = rd_u16(a->p + 18), fb = rd_u16(b->p + 18);
    if ((fa & 0x10) != (fb & 0x10)) return (fa & 0x10) ? 1 : -1;
    {   /* names without their terminating NUL, std::string::compare */
        const uint32_t la = a->p[12] ? a->p[12] - 1u : 0, lb = b->p[12] ? b->p[12] - 1u : 0;
        const int c = memcmp(a->p + 36, b->p + 36, la < lb ? la : lb);
        if (c) return c < 0 ? -1 : 1;
        if (la != lb) return la < lb ? -1 : 1;
    }
    if (fa != fb) return fa < fb ? -1 : 1;
    return a->idx < b->idx ? -1 : (a->idx > b->idx ? 1 : 0);
}

int oge_oracle_coordinate_order(const uint8_t *records, const uint64_t *offsets, uint64_t n, uint32_t *perm, uint8_t *tied) {
    uint64_t i;
    SortItem *it = (SortItem *) malloc((n ? n : 1) * sizeof(SortItem));
    if (!it) return -1;
    for (i = 0; i < n; i++) {
        it[i].p = records + offsets[i];
        it[i].idx = (uint32_t) i;
    }
    qsort(it, n, sizeof(SortItem), by_position);
    for (i = 0; i < n; i++) perm[i] = it[i].idx;
    if (tied) {      /* tied[k] = 1: output position k belongs to a group whose internal order the reference does not define */
        for (i = 0; i < n; i++) {
            int t = rd_i32(it[i].p + 4) == -1;
            if (!t && i > 0) {
                SortItem x = it[i - 1], y = it[i];
                x.idx = y.idx = 0;
                t = by_position(&x, &y) == 0;
            }
            if (!t && i + 1 < n) {
                SortItem x = it[i], y = it[i + 1];
                x.idx = y.idx = 0;
                t = by_position(&x, &y) == 0;
            }
            tied[i] = (uint8_t) t;
        }
    }
    free(it);
    return 0;
}
