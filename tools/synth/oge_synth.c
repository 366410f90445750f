/*
 * Synthetic coordinate-sorted BAM record generator for the dedup workloads C1..C5
 *  + a fast block_size-chain framer.  Host tooling for tests and bench.py;
 * not on the device path.  Output = raw BAM records back to back (each with its leading
 * block_size, layout of bam_deserializer.h:144-193 in the reference) + n+1 u64 offsets.
 *
 * Two-phase API so the (large) record buffer can be caller-owned pinned memory:
 *     h = oge_synth_plan(&cfg, &n_records, &n_bytes);
 *     oge_synth_emit(h, records, offsets, nthreads);
 *     oge_synth_free(h);
 * Generation is deterministic in cfg.seed and independent of nthreads.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t seed;
    uint64_t n_templates;        /* templates (a pair, a single read, ...) to draw */
    int32_t n_contigs;
    int32_t contig_len[256];
    int32_t read_len;
    int32_t insert_lo, insert_hi;   /* outer distance between unclipped starts + read_len */
    double dup_frac;             /* P(template copies the geometry of an earlier template) */
    double softclip_frac;        /* P(a read end gets a soft clip), per end */
    double hardclip_frac;        /* P(a clip is H instead of S) */
    double indel_frac;           /* P(read has a D / N / I op) */
    double rf_frac;              /* P(pair is RF rather than FR); ff_frac likewise */
    double ff_frac;
    double single_frac;          /* P(template is a single-end read) */
    double mate_unmapped_frac;   /* P(template is a pair whose second mate is unmapped) */
    double cross_contig_frac;    /* P(pair's mate lies on another contig) */
    double secondary_frac;       /* P(template spawns an extra secondary (0x100) record) */
    double supplementary_frac;   /* P(template spawns an extra 0x800 record (primary to the reference)) */
    double unmapped_pair_frac;   /* P(template is a fully unmapped pair, refID -1, file tail) */
    double predup_frac;          /* P(a record arrives with 0x400 already set) */
    double no_rg_frac;           /* P(read carries no RG tag) */
    double unknown_rg_frac;      /* P(read carries an RG id missing from the header) */
    double const_qual_frac;      /* P(read has constant qualities -> score ties) */
    double extra_tag_frac;       /* P(read has other tags in front of RG) */
    int32_t n_rg;                /* read groups rg1..rgN (ids "rg%d") */
    int32_t hot_loci;            /* >0: duplicates copy from this many hot templates, Zipf-skewed (C4) */
    int32_t dup_same_rg;         /* 1: a duplicate keeps its source's read group (else redrawn) */
    int32_t contig_lo, contig_hi;   /* hi > lo: templates are drawn on contigs [lo, hi) only (range shards) */
} SynthCfg;

typedef struct {
    int32_t ref, pos, mate_ref, mate_pos, tlen;
    uint32_t l_seq;
    uint64_t name_id, qseed;
    uint16_t flag, a, b, x, d;
    uint8_t rg, kind, clip_h, qual_mode, tagmode, qconst;
} Desc;

typedef struct {
    SynthCfg cfg;
    Desc *d;
    uint64_t n;
    uint64_t *off;       /* n+1 */
} Plan;

/* ---------------------------------------------------------------- RNG */
static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double urand(uint64_t *s) { return (splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t irand(uint64_t *s, uint64_t n) { return n ? (uint64_t)(((__uint128_t) splitmix(s) * n) >> 64) : 0; }

/* ---------------------------------------------------------------- geometry */
enum { K_PLAIN = 0, K_DEL = 1, K_SKIP = 2, K_INS = 3, K_NONE = 4 };
#define RG_NONE 0xFF
#define RG_UNKNOWN 0xFE

typedef struct {            /* one read's geometry in unclipped coordinates */
    int32_t u;              /* unclipped start */
    uint16_t a, b, x, d;
    uint8_t kind, clip_h, rev;
} ReadGeom;

typedef struct {            /* one template's geometry, the thing a duplicate copies */
    int32_t contig, contig2;
    ReadGeom r1, r2;
    uint8_t type;           /* 0 pair, 1 single, 2 pair with unmapped mate, 3 unmapped pair */
    uint8_t rg;
} Tmpl;

static int32_t ref_len_of(const ReadGeom *g, int32_t L) {
    int32_t m = L - g->a - g->b;
    if (g->kind == K_DEL || g->kind == K_SKIP) return m + g->d;
    if (g->kind == K_INS) return m - g->d;
    return m;
}

static void draw_clips(const SynthCfg *c, uint64_t *s, ReadGeom *g, int32_t L) {
    g->a = g->b = 0; g->clip_h = 0;
    if (urand(s) < c->softclip_frac) { g->a = (uint16_t)(1 + irand(s, L / 3)); if (urand(s) < c->hardclip_frac) g->clip_h |= 1; }
    if (urand(s) < c->softclip_frac) { g->b = (uint16_t)(1 + irand(s, L / 3)); if (urand(s) < c->hardclip_frac) g->clip_h |= 2; }
}

static void draw_read(const SynthCfg *c, uint64_t *s, ReadGeom *g, int32_t u, int rev) {
    int32_t L = c->read_len;
    g->u = u; g->rev = (uint8_t) rev; g->kind = K_PLAIN; g->x = g->d = 0;
    draw_clips(c, s, g, L);
    if (urand(s) < c->indel_frac) {
        int32_t m = L - g->a - g->b;
        if (m >= 12) {
            g->kind = (uint8_t)(1 + irand(s, 3));
            g->d = (uint16_t)(1 + irand(s, g->kind == K_SKIP ? 400 : 5));
            if (g->kind == K_INS && g->d > m - 8) g->d = 2;
            g->x = (uint16_t)(4 + irand(s, m - 8 - (g->kind == K_INS ? g->d : 0)));
        }
    }
}

static uint8_t draw_rg(const SynthCfg *c, uint64_t *s) {
    double r = urand(s);
    if (r < c->no_rg_frac) return RG_NONE;
    if (r < c->no_rg_frac + c->unknown_rg_frac) return RG_UNKNOWN;
    return (uint8_t) irand(s, c->n_rg > 0 ? c->n_rg : 1);
}

static int32_t draw_contig(const SynthCfg *c, uint64_t *s, const double *cum) {
    int32_t lo = 0, hi = c->n_contigs - 1;
    double base = 0.0, r;
    if (c->contig_hi > c->contig_lo) {
        lo = c->contig_lo; hi = c->contig_hi - 1;
        base = lo ? cum[lo - 1] : 0.0;
    }
    r = base + urand(s) * (cum[hi] - base);
    while (lo < hi) { int32_t mid = (lo + hi) / 2; if (cum[mid] > r) hi = mid; else lo = mid + 1; }
    return lo;
}

static void fill_desc(const SynthCfg *c, Desc *o, const ReadGeom *g, int32_t contig, uint64_t name_id,
                      uint64_t *s, uint8_t rg) {
    int32_t L = c->read_len;
    memset(o, 0, sizeof(*o));
    o->ref = contig;
    o->pos = g->u + g->a;
    o->a = g->a; o->b = g->b; o->x = g->x; o->d = g->d; o->kind = g->kind; o->clip_h = g->clip_h;
    o->l_seq = (uint32_t)(L - ((g->clip_h & 1) ? g->a : 0) - ((g->clip_h & 2) ? g->b : 0));
    o->name_id = name_id;
    o->qseed = splitmix(s);
    o->rg = rg;
    o->qual_mode = urand(s) < c->const_qual_frac;
    o->qconst = (uint8_t)(o->qual_mode ? (urand(s) < 0.7 ? 30 : 10 + irand(s, 31)) : 0);
    o->tagmode = urand(s) < c->extra_tag_frac ? (uint8_t)(1 + irand(s, 2)) : 0;
    if (urand(s) < c->predup_frac) o->flag |= 0x400;
    if (g->rev) o->flag |= 0x10;
}

static int desc_cmp(const void *pa, const void *pb) {
    const Desc *a = (const Desc *) pa, *b = (const Desc *) pb;
    uint32_t ra = (uint32_t) a->ref, rb = (uint32_t) b->ref;   /* -1 sorts last, as in a coordinate-sorted BAM */
    if (ra != rb) return ra < rb ? -1 : 1;
    if (a->pos != b->pos) return a->pos < b->pos ? -1 : 1;
    if (a->name_id != b->name_id) return a->name_id < b->name_id ? -1 : 1;
    return (int) (a->flag & 0xC0) - (int) (b->flag & 0xC0);
}

/* ---------------------------------------------------------------- sorting the plan
 * The plan is sorted by (ref, pos, name id, mate bits) -- qsort over 100 M descriptors was a third of the planning time of a
 * rank's shard in the N > 1 bench.  Same comparator, several threads: chunks sorted with qsort side by side, then merged
 * pairwise (left chunk first on equal keys).  Descriptors that compare equal are a measure-zero event of the generator
 * (same name, same mate, same position); the golden files of tests/ pin the output for the seeds they use. */
typedef struct { Desc *d, *tmp; uint64_t lo, mid, hi; } SortJob;

static void *sort_chunk_worker(void *arg) {
    SortJob *j = (SortJob *) arg;
    qsort(j->d + j->lo, j->hi - j->lo, sizeof(Desc), desc_cmp);
    return NULL;
}

static void *merge_worker(void *arg) {
    SortJob *j = (SortJob *) arg;
    uint64_t a = j->lo, b = j->mid, o = j->lo;
    while (a < j->mid && b < j->hi) {
        if (desc_cmp(&j->d[b], &j->d[a]) < 0) j->tmp[o++] = j->d[b++];
        else j->tmp[o++] = j->d[a++];
    }
    while (a < j->mid) j->tmp[o++] = j->d[a++];
    while (b < j->hi) j->tmp[o++] = j->d[b++];
    return NULL;
}

static void sort_descs(Desc *d, uint64_t n) {
    enum { T = 16 };
    pthread_t th[T];
    SortJob job[T];
    uint64_t bound[T + 1];
    Desc *tmp, *src = d, *dst;
    int t, parts = T, width;
    if (n < 200000 || !(tmp = (Desc *) malloc(n * sizeof(Desc)))) {
        qsort(d, n, sizeof(Desc), desc_cmp);
        return;
    }
    for (t = 0; t <= T; t++) bound[t] = n * (uint64_t) t / T;
    for (t = 0; t < T; t++) {
        job[t].d = d; job[t].tmp = tmp; job[t].lo = bound[t]; job[t].mid = bound[t]; job[t].hi = bound[t + 1];
        pthread_create(&th[t], NULL, sort_chunk_worker, &job[t]);
    }
    for (t = 0; t < T; t++) pthread_join(th[t], NULL);
    dst = tmp;
    for (width = 1; width < T; width *= 2) {      /* runs of `width` chunks -> runs of 2 * width */
        int k = 0;
        for (t = 0; t < T; t += 2 * width, k++) {
            job[k].d = src; job[k].tmp = dst; job[k].lo = bound[t]; job[k].mid = bound[t + width]; job[k].hi = bound[t + 2 * width];
            pthread_create(&th[k], NULL, merge_worker, &job[k]);
        }
        for (t = 0; t < k; t++) pthread_join(th[t], NULL);
        { Desc *x = src; src = dst; dst = x; }
        parts /= 2;
    }
    (void) parts;
    if (src != d) memcpy(d, src, n * sizeof(Desc));
    free(tmp);
}

static uint32_t n_cigar_of(const Desc *d) {
    uint32_t n;
    if (d->kind == K_NONE) return 0;
    n = 1 + (d->a ? 1 : 0) + (d->b ? 1 : 0);
    if (d->kind != K_PLAIN) n += 2;
    return n;
}

static const char *RG_UNKNOWN_ID = "rgX";

static uint32_t tags_len_of(const Desc *d) {
    uint32_t n = 0;
    char tmp[32];
    if (d->tagmode == 1) n += 4 + 7;                 /* NM:C + AS:i */
    if (d->tagmode == 2) n += (3 + 1 + 4 + 6) + (3 + 4);   /* XB:B:s,3 + MD:Z:150\0 */
    if (d->rg != RG_NONE) {
        if (d->rg == RG_UNKNOWN) n += 3 + (uint32_t) strlen(RG_UNKNOWN_ID) + 1;
        else n += 3 + (uint32_t) snprintf(tmp, sizeof tmp, "rg%d", d->rg + 1) + 1;
    }
    if (d->tagmode == 2) n += 4;                     /* XT:A:U after RG */
    return n;
}

#define NAME_LEN 24u
static uint32_t rec_size_of(const Desc *d) {
    return 4 + 32 + NAME_LEN + 4 * n_cigar_of(d) + (d->l_seq + 1) / 2 + d->l_seq + tags_len_of(d);
}

void *oge_synth_plan(const SynthCfg *cfg, uint64_t *n_records, uint64_t *n_bytes) {
    Plan *p = (Plan *) calloc(1, sizeof(Plan));
    const SynthCfg *c = cfg;
    uint64_t s = mix64(cfg->seed * 0x9E3779B97F4A7C15ULL + 12345), t, n = 0, cap;
    double cum[256];
    Tmpl *src;          /* geometry of non-duplicate templates, for duplicates to copy */
    uint64_t n_src = 0;
    int32_t i;
    if (!p) return NULL;
    p->cfg = *cfg;
    for (i = 0; i < c->n_contigs; i++) cum[i] = (i ? cum[i - 1] : 0.0) + (double) c->contig_len[i];
    cap = c->n_templates * 2 + c->n_templates / 4 + 16;
    p->d = (Desc *) malloc(cap * sizeof(Desc));
    src = (Tmpl *) malloc((c->n_templates + 1) * sizeof(Tmpl));
    if (!p->d || !src) { free(p->d); free(src); free(p); return NULL; }

    for (t = 0; t < c->n_templates; t++) {
        Tmpl g;
        int is_dup = n_src > 0 && urand(&s) < c->dup_frac;
        int32_t L = c->read_len;
        if (is_dup) {
            uint64_t k;
            if (c->hot_loci > 0) {
                uint64_t H = (uint64_t) c->hot_loci < n_src ? (uint64_t) c->hot_loci : n_src;
                double z = urand(&s);
                k = (uint64_t)((double) H * z * z * z);      /* skewed: low indices get long runs */
                if (k >= H) k = H - 1;
            } else k = irand(&s, n_src);
            g = src[k];
            if (!c->dup_same_rg) g.rg = draw_rg(c, &s);
            /* fresh clips on plain reads: the unclipped 5' coordinates do not move */
            if (g.r1.kind == K_PLAIN) draw_clips(c, &s, &g.r1, L);
            if (g.type == 0 && g.r2.kind == K_PLAIN) draw_clips(c, &s, &g.r2, L);
        } else {
            double r = urand(&s);
            int32_t ins, span;
            memset(&g, 0, sizeof g);
            g.rg = draw_rg(c, &s);
            if (r < c->single_frac) g.type = 1;
            else if (r < c->single_frac + c->mate_unmapped_frac) g.type = 2;
            else if (r < c->single_frac + c->mate_unmapped_frac + c->unmapped_pair_frac) g.type = 3;
            else g.type = 0;
            g.contig = draw_contig(c, &s, cum);
            ins = c->insert_lo + (int32_t) irand(&s, (uint64_t)(c->insert_hi - c->insert_lo + 1));
            if (ins < L) ins = L;
            span = c->contig_len[g.contig] - ins - 1200;
            if (span < 1) span = 1;
            {
                int32_t u1 = 600 + (int32_t) irand(&s, (uint64_t) span);
                int rev1 = 0, rev2 = 1;
                double o = urand(&s);
                if (g.type == 0) {
                    if (o < c->rf_frac) { rev1 = 1; rev2 = 0; }
                    else if (o < c->rf_frac + c->ff_frac) { rev1 = 0; rev2 = 0; }
                } else rev1 = urand(&s) < 0.5;
                draw_read(c, &s, &g.r1, u1, rev1);
                g.contig2 = g.contig;
                if (g.type == 0) {
                    int32_t u2 = u1 + ins - L;
                    int32_t c_lo = c->contig_hi > c->contig_lo ? c->contig_lo : 0;
                    int32_t c_n = c->contig_hi > c->contig_lo ? c->contig_hi - c->contig_lo : c->n_contigs;
                    if (urand(&s) < c->cross_contig_frac && c_n > 1) {
                        do g.contig2 = c_lo + (int32_t) irand(&s, (uint64_t) c_n); while (g.contig2 == g.contig);
                        span = c->contig_len[g.contig2] - 2 * L - 1200; if (span < 1) span = 1;
                        u2 = 600 + (int32_t) irand(&s, (uint64_t) span);
                    }
                    draw_read(c, &s, &g.r2, u2, rev2);
                }
            }
            src[n_src++] = g;
        }

        /* ---- emit descriptors for this template */
        {
            uint64_t name_id = t;
            if (n + 4 > cap) { cap = cap * 2; p->d = (Desc *) realloc(p->d, cap * sizeof(Desc)); }
            if (g.type == 1) {
                fill_desc(c, &p->d[n], &g.r1, g.contig, name_id, &s, g.rg);
                p->d[n].mate_ref = -1; p->d[n].mate_pos = -1;
                n++;
            } else if (g.type == 3) {
                int k;
                for (k = 0; k < 2; k++) {
                    Desc *o = &p->d[n];
                    ReadGeom z; memset(&z, 0, sizeof z); z.kind = K_NONE;
                    fill_desc(c, o, &z, -1, name_id, &s, g.rg);
                    o->kind = K_NONE; o->pos = -1; o->mate_ref = -1; o->mate_pos = -1;
                    o->flag = (uint16_t)((o->flag & 0x400) | 0x1 | 0x4 | 0x8 | (k ? 0x80 : 0x40));
                    n++;
                }
            } else if (g.type == 2) {
                Desc *m = &p->d[n], *um = &p->d[n + 1];
                ReadGeom z; memset(&z, 0, sizeof z); z.kind = K_NONE;
                fill_desc(c, m, &g.r1, g.contig, name_id, &s, g.rg);
                m->flag |= 0x1 | 0x8 | 0x40;                           /* 73 / 89 */
                m->mate_ref = m->ref; m->mate_pos = m->pos;
                fill_desc(c, um, &z, g.contig, name_id, &s, g.rg);
                um->kind = K_NONE; um->pos = m->pos;
                um->flag = (uint16_t)((um->flag & 0x400) | 0x1 | 0x4 | 0x80 | (g.r1.rev ? 0x20 : 0));   /* 133 / 165 */
                um->mate_ref = m->ref; um->mate_pos = m->pos;
                n += 2;
            } else {
                Desc *a = &p->d[n], *b = &p->d[n + 1];
                int32_t e1, e2;
                fill_desc(c, a, &g.r1, g.contig, name_id, &s, g.rg);
                fill_desc(c, b, &g.r2, g.contig2, name_id, &s, g.rg);
                a->flag |= 0x1 | 0x2 | 0x40 | (g.r2.rev ? 0x20 : 0);
                b->flag |= 0x1 | 0x2 | 0x80 | (g.r1.rev ? 0x20 : 0);
                a->mate_ref = b->ref; a->mate_pos = b->pos;
                b->mate_ref = a->ref; b->mate_pos = a->pos;
                if (a->ref == b->ref) {
                    e1 = a->pos + ref_len_of(&g.r1, L); e2 = b->pos + ref_len_of(&g.r2, L);
                    a->tlen = (e2 > e1 ? e2 : e1) - (a->pos < b->pos ? a->pos : b->pos);
                    if (a->pos > b->pos) a->tlen = -a->tlen;
                    b->tlen = -a->tlen;
                }
                n += 2;
                if (urand(&s) < c->supplementary_frac) {
                    Desc *x = &p->d[n];
                    ReadGeom z;
                    draw_read(c, &s, &z, g.r1.u + 50 + (int32_t) irand(&s, 200), g.r1.rev);
                    fill_desc(c, x, &z, g.contig, name_id, &s, g.rg);
                    x->flag |= (uint16_t)(0x800 | (a->flag & (0x1 | 0x2 | 0x20 | 0x40)));
                    x->mate_ref = a->mate_ref; x->mate_pos = a->mate_pos;
                    n++;
                }
            }
            if (g.type != 3 && urand(&s) < c->secondary_frac) {
                Desc *x = &p->d[n];
                ReadGeom z;
                int32_t ct = draw_contig(c, &s, cum);
                int32_t span = c->contig_len[ct] - 2 * L - 1200; if (span < 1) span = 1;
                draw_read(c, &s, &z, 600 + (int32_t) irand(&s, (uint64_t) span), urand(&s) < 0.5);
                fill_desc(c, x, &z, ct, name_id, &s, g.rg);
                x->flag |= 0x100;
                if (g.type == 0) { x->flag |= 0x1 | 0x40; x->mate_ref = g.contig2; x->mate_pos = g.r2.u + g.r2.a; }
                else { x->mate_ref = -1; x->mate_pos = -1; }
                n++;
            }
        }
    }
    free(src);
    sort_descs(p->d, n);
    p->n = n;
    p->off = (uint64_t *) malloc((n + 1) * sizeof(uint64_t));
    p->off[0] = 0;
    for (t = 0; t < n; t++) p->off[t + 1] = p->off[t] + rec_size_of(&p->d[t]);
    *n_records = n;
    *n_bytes = p->off[n];
    return p;
}

/* ---------------------------------------------------------------- emission */
static inline void put32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
static inline void put16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }

static uint32_t reg2bin(int32_t beg, int32_t end) {
    --end;
    if (beg >> 14 == end >> 14) return 4681 + (beg >> 14);
    if (beg >> 17 == end >> 17) return 585 + (beg >> 17);
    if (beg >> 20 == end >> 20) return 73 + (beg >> 20);
    if (beg >> 23 == end >> 23) return 9 + (beg >> 23);
    if (beg >> 26 == end >> 26) return 1 + (beg >> 26);
    return 0;
}

static void emit_one(const SynthCfg *c, const Desc *d, uint8_t *p, uint32_t size) {
    static const char HEX[] = "0123456789abcdef";
    uint32_t nc = n_cigar_of(d), i, L = (uint32_t) c->read_len;
    uint8_t *q;
    uint64_t s = d->qseed, h;
    int32_t reflen = 0;
    put32(p, size - 4);
    put32(p + 4, (uint32_t) d->ref);
    put32(p + 8, (uint32_t) d->pos);
    p[12] = NAME_LEN; p[13] = (d->flag & 0x4) ? 0 : 60;
    put16(p + 16, (uint16_t) nc);
    put16(p + 18, d->flag);
    put32(p + 20, d->l_seq);
    put32(p + 24, (uint32_t) d->mate_ref);
    put32(p + 28, (uint32_t) d->mate_pos);
    put32(p + 32, (uint32_t) d->tlen);
    q = p + 36;
    memcpy(q, "OGB200:", 7);
    h = mix64(d->name_id * 0xD6E8FEB86659FD93ULL + c->seed);
    for (i = 0; i < 16; i++) q[7 + i] = (uint8_t) HEX[(h >> (4 * i)) & 15];
    q[23] = 0;
    q += NAME_LEN;
    if (d->kind != K_NONE) {
        uint32_t m = L - d->a - d->b;
        if (d->a) { put32(q, ((uint32_t) d->a << 4) | ((d->clip_h & 1) ? 5u : 4u)); q += 4; }
        if (d->kind == K_PLAIN) { put32(q, (m << 4) | 0u); q += 4; reflen = (int32_t) m; }
        else if (d->kind == K_INS) {
            put32(q, ((uint32_t) d->x << 4) | 0u); put32(q + 4, ((uint32_t) d->d << 4) | 1u);
            put32(q + 8, ((m - d->x - d->d) << 4) | 0u); q += 12; reflen = (int32_t)(m - d->d);
        } else {
            put32(q, ((uint32_t) d->x << 4) | 0u); put32(q + 4, ((uint32_t) d->d << 4) | (d->kind == K_DEL ? 2u : 3u));
            put32(q + 8, ((m - d->x) << 4) | 0u); q += 12; reflen = (int32_t)(m + d->d);
        }
        if (d->b) { put32(q, ((uint32_t) d->b << 4) | ((d->clip_h & 2) ? 5u : 4u)); q += 4; }
    }
    {
        /* the reference's writer recomputes bin from (pos, GetEndPosition()) with end == pos for
         * CIGAR-less records (bam_serializer.h:112-116); emit that form so outputs compare bytewise */
        put16(p + 14, (uint16_t) reg2bin(d->pos, d->pos + reflen));
    }
    /* packed bases: random A/C/G/T nibbles */
    {
        uint32_t nb = (d->l_seq + 1) / 2;
        for (i = 0; i < nb; i += 8) {
            uint64_t r = splitmix(&s), v = 0;
            uint32_t k, lim = nb - i < 8 ? nb - i : 8;
            for (k = 0; k < 8; k++) v |= (uint64_t)((1u << ((r >> (4 * k)) & 3)) << 4 | (1u << ((r >> (4 * k + 2)) & 3))) << (8 * k);
            memcpy(q + i, &v, lim);
        }
        if (d->l_seq & 1) q[nb - 1] &= 0xF0;
        q += nb;
    }
    if (d->qual_mode) memset(q, d->qconst, d->l_seq);
    else {
        for (i = 0; i < d->l_seq; i += 8) {
            uint64_t r = splitmix(&s), v = 0;
            uint32_t k, lim = d->l_seq - i < 8 ? d->l_seq - i : 8;
            for (k = 0; k < 8; k++) v |= (uint64_t)(2 + ((((r >> (8 * k)) & 255) * 39) >> 8)) << (8 * k);
            memcpy(q + i, &v, lim);
        }
    }
    q += d->l_seq;
    if (d->tagmode == 1) {
        memcpy(q, "NMC", 3); q[3] = (uint8_t)(s & 7); q += 4;
        memcpy(q, "ASi", 3); put32(q + 3, (uint32_t)(100 + (s >> 8 & 63))); q += 7;
    } else if (d->tagmode == 2) {
        memcpy(q, "XBBs", 4); put32(q + 4, 3); put16(q + 8, 1); put16(q + 10, 2); put16(q + 12, 3); q += 14;
        memcpy(q, "MDZ150", 6); q[6] = 0; q += 7;
    }
    if (d->rg != RG_NONE) {
        char id[32];
        int l;
        if (d->rg == RG_UNKNOWN) l = snprintf(id, sizeof id, "%s", RG_UNKNOWN_ID);
        else l = snprintf(id, sizeof id, "rg%d", d->rg + 1);
        memcpy(q, "RGZ", 3); memcpy(q + 3, id, (size_t) l + 1); q += 3 + l + 1;
    }
    if (d->tagmode == 2) { memcpy(q, "XTAU", 4); q += 4; }
}

typedef struct { Plan *p; uint8_t *rec; uint64_t lo, hi; } Job;

static void *emit_worker(void *arg) {
    Job *j = (Job *) arg;
    uint64_t i;
    for (i = j->lo; i < j->hi; i++)
        emit_one(&j->p->cfg, &j->p->d[i], j->rec + j->p->off[i], (uint32_t)(j->p->off[i + 1] - j->p->off[i]));
    return NULL;
}

int oge_synth_emit(void *handle, uint8_t *records, uint64_t *offsets, int nthreads) {
    Plan *p = (Plan *) handle;
    pthread_t th[64];
    Job jobs[64];
    int t;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    memcpy(offsets, p->off, (p->n + 1) * sizeof(uint64_t));
    for (t = 0; t < nthreads; t++) {
        jobs[t].p = p; jobs[t].rec = records;
        jobs[t].lo = p->n * (uint64_t) t / (uint64_t) nthreads;
        jobs[t].hi = p->n * (uint64_t)(t + 1) / (uint64_t) nthreads;
        pthread_create(&th[t], NULL, emit_worker, &jobs[t]);
    }
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

void oge_synth_free(void *handle) {
    Plan *p = (Plan *) handle;
    if (!p) return;
    free(p->d); free(p->off); free(p);
}

/* Walk the block_size chain (reference: bam_deserializer.h:147-168).  offsets has room for
 * cap entries; returns the record count, or -1 on a malformed chain / cap overflow.
 * With offsets == NULL only counts. */
int64_t oge_frame_records(const uint8_t *buf, uint64_t len, uint64_t *offsets, uint64_t cap) {
    uint64_t pos = 0, n = 0;
    if (offsets) { if (cap == 0) return -1; offsets[0] = 0; }
    while (pos < len) {
        uint32_t bs;
        if (pos + 4 > len) return -1;
        memcpy(&bs, buf + pos, 4);
        if (bs < 32) return -1;
        pos += 4 + (uint64_t) bs;
        if (pos > len) return -1;
        n++;
        if (offsets) { if (n >= cap) return -1; offsets[n] = pos; }
    }
    return (int64_t) n;
}

/* (refID, pos) order of a coordinate-sorted file: unmapped (refID -1) last. */
static inline uint64_t sort_key_of(const uint8_t *rec) {
    int32_t ref, pos;
    memcpy(&ref, rec + 4, 4);
    memcpy(&pos, rec + 8, 4);
    if (ref < 0) return ~0ULL;
    return ((uint64_t)(uint32_t) ref << 32) | (uint32_t)(pos + 1);
}

/* Merge two coordinate-sorted record chains A and B' (B restricted to the records with keep[i] != 0)
 * into out_rec / out_off (n_out + 1 offsets); ties keep A first.  With out_rec == NULL only sizes
 * are computed.  Returns the number of records. */
uint64_t oge_merge_sorted(const uint8_t *a, const uint64_t *a_off, uint64_t na, const uint8_t *b, const uint64_t *b_off,
                          uint64_t nb, const uint8_t *keep, uint8_t *out_rec, uint64_t *out_off, uint64_t *out_bytes) {
    uint64_t i = 0, j = 0, n = 0, pos = 0;
    while (j < nb && !keep[j]) j++;
    if (out_off) out_off[0] = 0;
    while (i < na || j < nb) {
        int take_a = j >= nb || (i < na && sort_key_of(a + a_off[i]) <= sort_key_of(b + b_off[j]));
        const uint8_t *src = take_a ? a + a_off[i] : b + b_off[j];
        uint64_t len = take_a ? a_off[i + 1] - a_off[i] : b_off[j + 1] - b_off[j];
        if (out_rec) memcpy(out_rec + pos, src, len);
        pos += len;
        n++;
        if (out_off) out_off[n] = pos;
        if (take_a) i++;
        else { j++; while (j < nb && !keep[j]) j++; }
    }
    if (out_bytes) *out_bytes = pos;
    return n;
}

/* Range shards that cut INSIDE contigs: a rank draws its reads on "pieces" (stretches of real contigs) as if every piece
 * were a contig of its own, and this maps the records onto the real contigs afterwards: refID piece -> piece_ref[piece],
 * pos += piece_start[piece], same for the mate fields.  Records without a reference (refID -1) are left alone.
 * The bin field keeps its piece-relative value (nothing on the dedup path reads it). */
typedef struct {
    uint8_t *rec;
    const uint64_t *off;
    uint64_t lo, hi;
    const int32_t *piece_ref, *piece_start;
    int32_t n_pieces;
} RemapJob;

static void *remap_worker(void *arg) {
    RemapJob *j = (RemapJob *) arg;
    uint64_t i;
    for (i = j->lo; i < j->hi; i++) {
        uint8_t *p = j->rec + j->off[i];
        int32_t ref, pos, mref, mpos;
        memcpy(&ref, p + 4, 4); memcpy(&pos, p + 8, 4); memcpy(&mref, p + 24, 4); memcpy(&mpos, p + 28, 4);
        if (ref >= 0 && ref < j->n_pieces) { pos += j->piece_start[ref]; ref = j->piece_ref[ref]; }
        if (mref >= 0 && mref < j->n_pieces) { mpos += j->piece_start[mref]; mref = j->piece_ref[mref]; }
        memcpy(p + 4, &ref, 4); memcpy(p + 8, &pos, 4); memcpy(p + 24, &mref, 4); memcpy(p + 28, &mpos, 4);
    }
    return NULL;
}

int oge_synth_remap(uint8_t *records, const uint64_t *offsets, uint64_t n, const int32_t *piece_ref, const int32_t *piece_start,
                    int32_t n_pieces, int nthreads) {
    pthread_t th[64];
    RemapJob jobs[64];
    int t, nt = nthreads < 1 ? 1 : (nthreads > 64 ? 64 : nthreads);
    for (t = 0; t < nt; t++) {
        jobs[t].rec = records; jobs[t].off = offsets; jobs[t].lo = n * (uint64_t) t / nt; jobs[t].hi = n * (uint64_t)(t + 1) / nt;
        jobs[t].piece_ref = piece_ref; jobs[t].piece_start = piece_start; jobs[t].n_pieces = n_pieces;
        if (pthread_create(&th[t], NULL, remap_worker, &jobs[t])) return -1;
    }
    for (t = 0; t < nt; t++) pthread_join(th[t], NULL);
    return 0;
}
