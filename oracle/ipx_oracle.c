/* ipx_oracle.c -- CPU oracle (plain C) for the junction-extraction path.
 * TEST INFRASTRUCTURE ONLY -- see ipx_oracle.h.  Sequential, one record at a time, in
 * stream order, exactly as the reference processes them; citations are file:line into
 * the reference's src/.
 */
#define _GNU_SOURCE
#include "ipx_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- small utilities */
typedef struct {
    char *s;
    size_t n, cap;
} sbuf;

static void sb_reserve(sbuf *b, size_t extra) {
    if (b->n + extra + 1 > b->cap) {
        size_t c = b->cap ? b->cap * 2 : 256;
        while (c < b->n + extra + 1) c *= 2;
        b->s = (char *)realloc(b->s, c);
        b->cap = c;
    }
}

static void sb_printf(sbuf *b, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    char tmp[512];
    int k = vsnprintf(tmp, sizeof(tmp), fmt, ap);
    va_end(ap);
    if (k < 0) return;
    if ((size_t)k >= sizeof(tmp)) {
        sb_reserve(b, (size_t)k);
        va_start(ap, fmt);
        vsnprintf(b->s + b->n, (size_t)k + 1, fmt, ap);
        va_end(ap);
    } else {
        sb_reserve(b, (size_t)k);
        memcpy(b->s + b->n, tmp, (size_t)k + 1);
    }
    b->n += (size_t)k;
}

/* ---------------------------------------------------------------- junction store */
typedef struct {
    int32_t chrom;
    uint32_t intron_start, intron_end; /* JunctionKey, junctions.hh:57-70 (uint32) */
    uint32_t anchor_start, anchor_end; /* Junction, junctions.hh:109-110 */
    char strand;
    char sites[8];
    uint32_t ijunc;
    uint64_t creation; /* (record stream index << 20) | candidate ordinal of the creating event */
    uint64_t counts[3];
    uint16_t *offsets; /* read_offsets, junctions.hh:120 */
    size_t n_off, cap_off;
    float confidence; /* only used by the merge fold */
    int used;
} ojunc;

struct ipxo_state {
    ipx_params p;
    int32_t n_targets;
    char **targets;
    /* genome */
    int32_t n_seqs;
    char **seq_names;
    uint64_t *seq_len;
    char **seqs;
    int32_t *target_seq; /* target -> genome sequence index or -1 */
    int have_genome;
    /* open-addressing table keyed (chrom, intron_start, intron_end) */
    ojunc *tab;
    size_t cap, used;
    uint32_t ijunc; /* ijunc_, junctions_extractor.hh:129 */
    /* missing genome warnings, junctions_extractor.cc:411-417 */
    int32_t *missing;
    size_t n_missing;
    uint64_t n_events, cigar_live;
    sbuf err, log;
    /* deferred fatal error */
    int status;
    uint64_t error_seq;
    int32_t error_tid;
    char error_text[256];
};

static uint64_t mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

static size_t slot_of(int32_t chrom, uint32_t s, uint32_t e, size_t cap) {
    return (size_t)(mix(((uint64_t)(uint32_t)chrom << 40) ^ ((uint64_t)s << 20) ^ e ^ ((uint64_t)e << 44)) & (cap - 1));
}

static void tab_grow(ipxo_state *st) {
    size_t ncap = st->cap ? st->cap * 2 : 1024;
    ojunc *nt = (ojunc *)calloc(ncap, sizeof(ojunc));
    for (size_t i = 0; i < st->cap; i++) {
        if (!st->tab[i].used) continue;
        size_t k = slot_of(st->tab[i].chrom, st->tab[i].intron_start, st->tab[i].intron_end, ncap);
        while (nt[k].used) k = (k + 1) & (ncap - 1);
        nt[k] = st->tab[i];
    }
    free(st->tab);
    st->tab = nt;
    st->cap = ncap;
}

/* returns existing junction or NULL */
static ojunc *tab_find(ipxo_state *st, int32_t chrom, uint32_t s, uint32_t e) {
    if (st->cap == 0) return NULL;
    size_t k = slot_of(chrom, s, e, st->cap);
    while (st->tab[k].used) {
        ojunc *j = &st->tab[k];
        if (j->chrom == chrom && j->intron_start == s && j->intron_end == e) return j;
        k = (k + 1) & (st->cap - 1);
    }
    return NULL;
}

static ojunc *tab_insert(ipxo_state *st, int32_t chrom, uint32_t s, uint32_t e) {
    if ((st->used + 1) * 2 > st->cap) tab_grow(st);
    size_t k = slot_of(chrom, s, e, st->cap);
    while (st->tab[k].used) k = (k + 1) & (st->cap - 1);
    ojunc *j = &st->tab[k];
    memset(j, 0, sizeof(*j));
    j->used = 1;
    j->chrom = chrom;
    j->intron_start = s;
    j->intron_end = e;
    st->used++;
    return j;
}

static void tab_clear(ipxo_state *st) {
    for (size_t i = 0; i < st->cap; i++)
        if (st->tab[i].used) free(st->tab[i].offsets);
    free(st->tab);
    st->tab = NULL;
    st->cap = st->used = 0;
}

/* Junction::count_read, junctions.hh:163-167 (offset truncates to uint16) */
static void count_read(ojunc *j, int cat, uint32_t read_offset) {
    j->counts[cat] += 1;
    if (j->n_off == j->cap_off) {
        j->cap_off = j->cap_off ? j->cap_off * 2 : 4;
        j->offsets = (uint16_t *)realloc(j->offsets, j->cap_off * sizeof(uint16_t));
    }
    j->offsets[j->n_off++] = (uint16_t)read_offset;
}

/* ---------------------------------------------------------------- lifecycle */
ipxo_state *ipxo_create(const ipx_params *params) {
    ipxo_state *st = (ipxo_state *)calloc(1, sizeof(*st));
    st->p = *params;
    return st;
}

static void free_genome(ipxo_state *st) {
    for (int32_t i = 0; i < st->n_seqs; i++) {
        free(st->seq_names[i]);
        free(st->seqs[i]);
    }
    free(st->seq_names);
    free(st->seq_len);
    free(st->seqs);
    st->n_seqs = 0;
}

void ipxo_destroy(ipxo_state *st) {
    if (!st) return;
    tab_clear(st);
    for (int32_t i = 0; i < st->n_targets; i++) free(st->targets[i]);
    free(st->targets);
    free(st->target_seq);
    free_genome(st);
    free(st->missing);
    free(st->err.s);
    free(st->log.s);
    free(st);
}

const char *ipxo_last_error(const ipxo_state *st) {
    return st->err.s ? st->err.s : "";
}

const char *ipxo_stderr_text(const ipxo_state *st) {
    return st->log.s ? st->log.s : "";
}

uint64_t ipxo_event_count(const ipxo_state *st) {
    return st->n_events;
}

uint64_t ipxo_cigar_live(const ipxo_state *st) {
    return st->cigar_live;
}

static void map_targets_to_genome(ipxo_state *st) {
    free(st->target_seq);
    st->target_seq = (int32_t *)malloc(sizeof(int32_t) * (size_t)(st->n_targets > 0 ? st->n_targets : 1));
    for (int32_t t = 0; t < st->n_targets; t++) {
        st->target_seq[t] = -1;
        for (int32_t g = 0; g < st->n_seqs; g++)
            if (strcmp(st->targets[t], st->seq_names[g]) == 0) {
                st->target_seq[t] = g;
                break;
            }
    }
}

/* save_targets, junctions_extractor.cc:564-568 */
int ipxo_set_targets(ipxo_state *st, int32_t n, const char *const *names) {
    for (int32_t i = 0; i < st->n_targets; i++) free(st->targets[i]);
    free(st->targets);
    st->targets = (char **)malloc(sizeof(char *) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; i++) st->targets[i] = strdup(names[i]);
    st->n_targets = n;
    map_targets_to_genome(st);
    return IPX_OK;
}

/* Genome::Genome, genome.cc:44-50 */
int ipxo_set_genome(ipxo_state *st, int32_t n, const char *const *names, const uint64_t *lengths,
                    const char *const *seqs) {
    free_genome(st);
    st->seq_names = (char **)malloc(sizeof(char *) * (size_t)(n > 0 ? n : 1));
    st->seqs = (char **)malloc(sizeof(char *) * (size_t)(n > 0 ? n : 1));
    st->seq_len = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; i++) {
        st->seq_names[i] = strdup(names[i]);
        st->seq_len[i] = lengths[i];
        st->seqs[i] = (char *)malloc(lengths[i] + 1);
        memcpy(st->seqs[i], seqs[i], lengths[i]);
        st->seqs[i][lengths[i]] = 0;
    }
    st->n_seqs = n;
    st->have_genome = 1;
    map_targets_to_genome(st);
    return IPX_OK;
}

/* ---------------------------------------------------------------- motif helpers */
/* is_canonical, splice_juncs.cc:25-38: GT/AG, GC/AG, AT/AC in any letter case */
static int is_canonical(const char *s) {
    if (strlen(s) != 5) return 0;
    char u[6];
    for (int i = 0; i < 5; i++) u[i] = (s[i] >= 'a' && s[i] <= 'z') ? (char)(s[i] - 32) : s[i];
    u[5] = 0;
    return strcmp(u, "GT/AG") == 0 || strcmp(u, "GC/AG") == 0 || strcmp(u, "AT/AC") == 0;
}

/* dna_reverse_complement, genome.cc:14-42: only ACGTacgt are complemented */
static void revcomp(const char *in, char *out) {
    size_t n = strlen(in);
    for (size_t i = 0; i < n; i++) {
        char c = in[i], r = c;
        switch (c) {
            case 'A': r = 'T'; break;
            case 'T': r = 'A'; break;
            case 'G': r = 'C'; break;
            case 'C': r = 'G'; break;
            case 'a': r = 't'; break;
            case 't': r = 'a'; break;
            case 'g': r = 'c'; break;
            case 'c': r = 'g'; break;
            default: break;
        }
        out[n - 1 - i] = r;
    }
    out[n] = 0;
}

static void str_upper(char *s) {
    for (; *s; s++)
        if (*s >= 'a' && *s <= 'z') *s = (char)(*s - 32);
}

static void str_lower(char *s) {
    for (; *s; s++)
        if (*s >= 'A' && *s <= 'Z') *s = (char)(*s + 32);
}

/* Genome::fetch, genome.cc:62-74 on top of faidx_fetch_seq clipping; returns 0 ok, else error set */
static int genome_fetch(ipxo_state *st, int32_t tid, int64_t start, int64_t end, char *out) {
    int32_t g = st->target_seq[tid];
    if (g < 0) {
        snprintf(st->error_text, sizeof(st->error_text), "can't load genome sequence for: %s:%d-%d",
                 st->targets[tid], (int)start, (int)end);
        return IPX_ERR_GENOME;
    }
    int64_t len = (int64_t)st->seq_len[g];
    int64_t b = start, e = end - 1;
    if (e < b - 1) e = b - 1;
    if (b < 0) b = 0;
    if (b >= len) b = len;
    if (e < 0) e = 0; else if (e >= len) e = len - 1;
    int64_t got = e - b + 1;
    if (got < 0) got = 0;
    if (got != end - start) {
        snprintf(st->error_text, sizeof(st->error_text), "unexpected length returned for: %s:%d-%d: %d",
                 st->targets[tid], (int)start, (int)end, (int)got);
        return IPX_ERR_GENOME;
    }
    memcpy(out, st->seqs[g] + b, (size_t)got);
    out[got] = 0;
    return IPX_OK;
}

/* get_splice_sites, junctions_extractor.cc:405-428 */
static int get_splice_sites(ipxo_state *st, int32_t tid, char strand, int64_t i0, int64_t i1, char *sites) {
    sites[0] = 0;
    if (!st->have_genome) return IPX_OK;
    if (st->p.skip_missing_targets && st->target_seq[tid] < 0) {
        int warned = 0;
        for (size_t k = 0; k < st->n_missing; k++)
            if (strcmp(st->targets[st->missing[k]], st->targets[tid]) == 0) warned = 1;
        if (!warned) {
            st->missing = (int32_t *)realloc(st->missing, sizeof(int32_t) * (st->n_missing + 1));
            st->missing[st->n_missing++] = tid;
            sb_printf(&st->log, "Warning: genomic sequence not found for %s splice junctions not available\n",
                      st->targets[tid]);
        }
        return IPX_OK;
    }
    char a[8], b[8], s[24];
    int rc = genome_fetch(st, tid, i0, i0 + 2, a);
    if (rc) return rc;
    rc = genome_fetch(st, tid, i1 - 2, i1, b);
    if (rc) return rc;
    snprintf(s, sizeof(s), "%s/%s", a, b);
    str_upper(s);
    if (strand == '-') {
        char r[8];
        revcomp(s, r);
        strcpy(s, r);
    }
    if (!is_canonical(s)) str_lower(s);
    strcpy(sites, s);
    return IPX_OK;
}

/* correct_strand, junctions_extractor.cc:431-444 */
static char correct_strand(char strand, char *sites) {
    if (sites[0] && strand == '.') {
        if (is_canonical(sites)) return '+';
        char r[8];
        revcomp(sites, r);
        if (is_canonical(r)) {
            str_upper(r);
            strcpy(sites, r);
            return '-';
        }
    }
    return strand;
}

/* junction_filter_check, splice_juncs.cc:52-61 */
static int filter_check(int filter, const char *sites) {
    if (sites[0] == 0) return 1;
    if (filter == IPX_FILTER_CANON) return is_canonical(sites);
    return 1;
}

/* ---------------------------------------------------------------- per-record classification */
/* get_read_category, junctions_extractor.cc:71-95 with get_category_from_tag :59-68 */
static int read_category(uint16_t flag, uint8_t nh_class) {
    int by_tag = (nh_class == IPX_NH_ONE) ? IPX_CAT_UNIQ : (nh_class == IPX_NH_MANY ? IPX_CAT_MULTI : IPX_CAT_UNSURE);
    if (flag & 0x100) return IPX_CAT_MULTI;
    if (flag & 0x800) return IPX_CAT_MULTI;
    if (flag & 0x1) {
        if ((flag & 0x2) && !(flag & 0x4)) return by_tag;
        return IPX_CAT_UNSURE;
    }
    return by_tag;
}

/* get_junction_strand, junctions_extractor.cc:333-371 */
static char read_strand(const ipx_params *p, uint16_t flag, uint8_t xs) {
    if (p->strandness != IPX_STRAND_UN) {
        int reversed = (flag & 0x10) != 0, mate_reversed = (flag & 0x20) != 0;
        int first_strand = ((flag & 0x40) != 0) ^ reversed;
        int second_strand = ((flag & 0x80) != 0) ^ mate_reversed;
        int is_rf = (p->strandness == IPX_STRAND_RF);
        /* bool_strandness = strandness_ - 1 is 0 for RF; "^= !bool_strandness" therefore flips for RF only */
        first_strand ^= is_rf;
        second_strand ^= is_rf;
        if (first_strand != second_strand) return '.';
        return first_strand ? '+' : '-';
    }
    if (xs == IPX_XS_ABSENT) return '.';
    if (xs == IPX_XS_OTHER) return '.';
    return (char)xs;
}

/* ---------------------------------------------------------------- the per-read automaton */
enum { LIMBO = 0, LEFT = 1, INTRON = 2, RIGHT = 3 };

typedef struct {
    ipxo_state *st;
    int32_t tid;
    int64_t pos;
    uint16_t flag;
    uint8_t nh_class;
    char strand;
    int state;
    int64_t tpos, a0, i0, i1, a1;
    uint32_t lmm, rmm, lind, rind, cur;
    uint64_t seq;     /* stream index of the record */
    uint32_t ordinal; /* candidates the automaton has handed to record_junction so far */
} walker;

/* junction_qc, junctions_extractor.cc:384-401 (signed 64-bit comparisons) */
static int junction_qc(const walker *w) {
    const ipx_params *p = &w->st->p;
    if (w->flag & 0x100) return 0;
    if ((p->excludes & IPX_EXCLUDE_MULTI) && w->nh_class == IPX_NH_MANY) return 0;
    int64_t ilen = w->i1 - w->i0;
    if (ilen < (int64_t)p->min_intron_length || ilen > (int64_t)p->max_intron_length) return 0;
    int64_t left = (w->i0 - w->a0) - (int64_t)(uint32_t)(w->lmm + w->lind);
    int64_t right = (w->a1 - w->i1) - (int64_t)(uint32_t)(w->rmm + w->rind);
    if (left < (int64_t)p->min_anchor_length || right < (int64_t)p->min_anchor_length) return 0;
    return 1;
}

/* record_junction -> add_junction -> create/update_junction, junctions_extractor.cc:446-530 */
static int record_junction(walker *w) {
    ipxo_state *st = w->st;
    const uint32_t ordinal = w->ordinal++;
    if (!junction_qc(w)) return IPX_OK;
    st->n_events++;
    uint32_t ks = (uint32_t)w->i0, ke = (uint32_t)w->i1;
    int cat = read_category(w->flag, w->nh_class);
    uint32_t offset = (uint32_t)(w->i0 - w->pos);
    ojunc *j = tab_find(st, w->tid, ks, ke);
    if (j == NULL) {
        char sites[8];
        int rc = get_splice_sites(st, w->tid, w->strand, w->i0, w->i1, sites);
        if (rc) return rc;
        char strand = correct_strand(w->strand, sites);
        if (!filter_check(st->p.junction_filter, sites)) return IPX_OK;
        j = tab_insert(st, w->tid, ks, ke);
        j->anchor_start = (uint32_t)w->a0;
        j->anchor_end = (uint32_t)w->a1;
        j->strand = strand;
        strcpy(j->sites, sites);
        j->ijunc = st->ijunc++;
        j->creation = (w->seq << 20) | (ordinal & 0xFFFFFu); /* ipx_table.creation_seq */
        count_read(j, cat, offset);
    } else {
        if (w->a0 < (int64_t)j->anchor_start) j->anchor_start = (uint32_t)w->a0;
        if (w->a1 > (int64_t)j->anchor_end) j->anchor_end = (uint32_t)w->a1;
        count_read(j, cat, offset);
    }
    return IPX_OK;
}

/* enter_limbo, junctions_extractor.cc:153-158 */
static int enter_limbo(walker *w) {
    int rc = IPX_OK;
    if (w->state == RIGHT) rc = record_junction(w);
    w->state = LIMBO;
    return rc;
}

/* ReadJunctionExtractor::parse_read and the enter_* handlers, junctions_extractor.cc:208-318 */
static int walk_read(walker *w, const uint32_t *cigar, uint32_t n_cigar, uint32_t max_indel) {
    int rc;
    for (uint32_t k = 0; k < n_cigar; k++) {
        uint32_t op = cigar[k] & 15, len = cigar[k] >> 4;
        switch (op) {
            case 3: /* N, enter_intron :208-229 */
                if (w->state == RIGHT) {
                    if ((rc = record_junction(w))) return rc;
                    w->a0 = w->i1;
                    w->lmm = w->rmm; w->rmm = 0;
                    w->lind = w->rind; w->rind = 0;
                    w->state = LEFT;
                }
                if (w->state == LEFT) {
                    w->i0 = w->tpos;
                    w->i1 = w->tpos + len;
                    w->state = INTRON;
                }
                w->cur = 0;
                w->tpos += len;
                break;
            case 0: case 7: case 8: /* M = X, enter_aligned :232-254 */
                if (w->state == LIMBO) {
                    w->a0 = w->tpos;
                    w->a1 = w->tpos + len;
                    w->lmm = w->rmm = 0;
                    w->state = LEFT;
                } else if (w->state == LEFT || w->state == RIGHT) {
                    w->a1 = w->tpos + len;
                } else {
                    w->i1 = w->tpos;
                    w->a1 = w->tpos + len;
                    w->state = RIGHT;
                }
                if (op == 8) {
                    if (w->state == LEFT) w->lmm += len;
                    else if (w->state == RIGHT) w->rmm += len;
                }
                w->cur = 0;
                w->tpos += len;
                break;
            case 2: case 1: /* D :257-265, I :268-275 */
                if (w->state == LEFT) w->lind += len;
                else if (w->state == RIGHT) w->rind += len;
                w->cur += len;
                if (w->cur > max_indel)
                    if ((rc = enter_limbo(w))) return rc;
                if (op == 2) w->tpos += len;
                break;
            case 4: case 5: /* S, H :306-311 */
                if ((rc = enter_limbo(w))) return rc;
                break;
            default: { /* :312-313 */
                char c = (op < 10) ? "MIDNSHP=XB"[op] : '?';
                snprintf(w->st->error_text, sizeof(w->st->error_text), "Unknown cigar operation '%c'", c);
                return IPX_ERR_CIGAR;
            }
        }
    }
    if (w->state == RIGHT) return record_junction(w); /* finish :278-282 */
    return IPX_OK;
}

/* process_alignment for every record of the batch, junctions_extractor.cc:533-549 */
int ipxo_push_batch(ipxo_state *st, const ipx_batch *b) {
    if (st->status != IPX_OK) return IPX_OK; /* the reference would already have exited */
    for (uint64_t r = 0; r < b->n_records; r++) {
        uint32_t c0 = b->cigar_off[r], c1 = b->cigar_off[r + 1];
        uint32_t n_cigar = c1 - c0;
        int32_t tid = b->tid[r];
        uint16_t flag = b->flag[r];
        if (n_cigar <= 1 || tid < 0 || (flag & (0x200 | 0x400))) continue;
        st->cigar_live += n_cigar;
        walker w;
        memset(&w, 0, sizeof(w));
        w.st = st;
        w.tid = tid;
        w.pos = b->pos[r];
        w.flag = flag;
        w.nh_class = b->nh_class[r];
        w.strand = read_strand(&st->p, flag, b->xs[r]);
        w.seq = b->seq_base + r;
        w.state = LIMBO;
        w.tpos = w.pos;
        w.a0 = w.i0 = w.i1 = w.a1 = -1;
        int rc = walk_read(&w, b->cigar + c0, n_cigar, st->p.max_anchor_indel_size);
        if (rc != IPX_OK) {
            st->status = rc;
            st->error_seq = b->seq_base + r;
            st->error_tid = tid;
            st->err.n = 0;
            sb_printf(&st->err, "%s", st->error_text);
            return IPX_OK; /* deferred: reported by ipxo_flush like the device path */
        }
    }
    return IPX_OK;
}

/* ---------------------------------------------------------------- table construction */
static ipxo_state *g_sort_state;

/* JunctionKey operator<, junctions.hh:77-100: chromosome name as a byte string, then start, end */
static int cmp_key(const void *pa, const void *pb) {
    const ojunc *a = *(const ojunc *const *)pa, *b = *(const ojunc *const *)pb;
    if (a->chrom != b->chrom) {
        int c = strcmp(g_sort_state->targets[a->chrom], g_sort_state->targets[b->chrom]);
        if (c) return c;
    }
    if (a->intron_start != b->intron_start) return a->intron_start < b->intron_start ? -1 : 1;
    if (a->intron_end != b->intron_end) return a->intron_end < b->intron_end ? -1 : 1;
    return 0;
}

static int cmp_u16(const void *a, const void *b) {
    uint16_t x = *(const uint16_t *)a, y = *(const uint16_t *)b;
    return x < y ? -1 : (x > y);
}

/* Junction::calculate_confidence, junctions.cc:54-70; offsets must be sorted */
static float confidence_of(const uint16_t *off, size_t n) {
    float summation = 0.0f;
    for (size_t i = 0; i < n;) {
        uint16_t pos = off[i];
        unsigned cnt = 0;
        for (; i < n && off[i] == pos; i++) cnt++;
        float p = (float)cnt / n;
        if (p != 0.0f) summation += p * log2f(p);
    }
    return (summation == 0.0f) ? 0.0f : -summation;
}

typedef struct {
    ipx_table t;
    int32_t *chrom;
    uint32_t *i0, *i1, *a0, *a1, *ijunc, *hist_cnt;
    uint64_t *creation;
    char *strand, *sites;
    uint64_t *counts, *hist_off;
    float *conf;
    int32_t *missing;
} otable;

static otable *table_alloc(size_t n, size_t n_hist) {
    otable *o = (otable *)calloc(1, sizeof(otable));
    size_t m = n ? n : 1;
    o->chrom = (int32_t *)calloc(m, sizeof(int32_t));
    o->i0 = (uint32_t *)calloc(m, 4);
    o->i1 = (uint32_t *)calloc(m, 4);
    o->a0 = (uint32_t *)calloc(m, 4);
    o->a1 = (uint32_t *)calloc(m, 4);
    o->ijunc = (uint32_t *)calloc(m, 4);
    o->creation = (uint64_t *)calloc(m, 8);
    o->strand = (char *)calloc(m, 1);
    o->sites = (char *)calloc(m, 8);
    o->counts = (uint64_t *)calloc(m * 3, 8);
    o->conf = (float *)calloc(m, 4);
    o->hist_off = (uint64_t *)calloc(m + 1, 8);
    o->hist_cnt = (uint32_t *)calloc(n_hist ? n_hist : 1, 4);
    o->t.n_junctions = n;
    o->t.chrom = o->chrom;
    o->t.intron_start = o->i0;
    o->t.intron_end = o->i1;
    o->t.anchor_start = o->a0;
    o->t.anchor_end = o->a1;
    o->t.strand = o->strand;
    o->t.splice_sites = o->sites;
    o->t.read_counts = o->counts;
    o->t.ijunc = o->ijunc;
    o->t.creation_seq = o->creation;
    o->t.confidence = o->conf;
    o->t.hist_off = o->hist_off;
    o->t.hist_cnt = o->hist_cnt;
    o->t.n_hist = n_hist;
    return o;
}

void ipxo_table_free(ipx_table *t) {
    if (!t) return;
    otable *o = (otable *)t;
    free(o->chrom); free(o->i0); free(o->i1); free(o->a0); free(o->a1); free(o->ijunc); free(o->creation);
    free(o->strand); free(o->sites); free(o->counts); free(o->conf); free(o->hist_off); free(o->hist_cnt);
    free(o->missing);
    free(o);
}

static ojunc **sorted_junctions(ipxo_state *st, size_t *n_out) {
    ojunc **v = (ojunc **)malloc(sizeof(ojunc *) * (st->used ? st->used : 1));
    size_t n = 0;
    for (size_t i = 0; i < st->cap; i++)
        if (st->tab[i].used) v[n++] = &st->tab[i];
    g_sort_state = st;
    qsort(v, n, sizeof(ojunc *), cmp_key);
    *n_out = n;
    return v;
}

/* get_junctions() in map order + clear(), junctions.hh:232-245 */
int ipxo_flush(ipxo_state *st, ipx_table **out) {
    size_t n;
    ojunc **v = sorted_junctions(st, &n);
    size_t n_hist = 0;
    for (size_t i = 0; i < n; i++) {
        qsort(v[i]->offsets, v[i]->n_off, sizeof(uint16_t), cmp_u16);
        for (size_t k = 0; k < v[i]->n_off; k++)
            if (k == 0 || v[i]->offsets[k] != v[i]->offsets[k - 1]) n_hist++;
    }
    otable *o = table_alloc(n, n_hist);
    size_t h = 0;
    for (size_t i = 0; i < n; i++) {
        ojunc *j = v[i];
        o->chrom[i] = j->chrom;
        o->i0[i] = j->intron_start;
        o->i1[i] = j->intron_end;
        o->a0[i] = j->anchor_start;
        o->a1[i] = j->anchor_end;
        o->strand[i] = j->strand;
        memcpy(o->sites + 8 * i, j->sites, 8);
        o->ijunc[i] = j->ijunc;
        o->creation[i] = j->creation;
        for (int c = 0; c < 3; c++) o->counts[3 * i + c] = j->counts[c];
        /* lazy_get_confidence, junctions.cc:44-51 */
        o->conf[i] = (j->counts[0] + j->counts[1] + j->counts[2] == 0) ? 0.0f : confidence_of(j->offsets, j->n_off);
        o->hist_off[i] = h;
        for (size_t k = 0; k < j->n_off;) {
            size_t e = k;
            while (e < j->n_off && j->offsets[e] == j->offsets[k]) e++;
            o->hist_cnt[h++] = (uint32_t)(e - k);
            k = e;
        }
    }
    o->hist_off[n] = h;
    o->t.status = st->status;
    o->t.first_error_seq = st->error_seq;
    o->t.error_tid = st->error_tid;
    memcpy(o->t.error_text, st->error_text, sizeof(o->t.error_text));
    o->missing = (int32_t *)malloc(sizeof(int32_t) * (st->n_missing ? st->n_missing : 1));
    memcpy(o->missing, st->missing, sizeof(int32_t) * st->n_missing);
    o->t.missing_targets = o->missing;
    o->t.n_missing_targets = st->n_missing;
    free(v);
    tab_clear(st);
    *out = &o->t;
    return IPX_OK;
}

/* ---------------------------------------------------------------- merge fold */
static void describe(ipxo_state *st, int32_t chrom, uint32_t s, uint32_t e, char strand, const char *sites, char *out,
                     size_t n) {
    /* Junction::get_description, junctions.cc:72-74 */
    snprintf(out, n, "%s:%d-%d[%c/%s]", st->targets[chrom], (int)s, (int)e, strand, sites);
}

/* process_junction + Junction::merge over all rows, intronProspectorMerge.cc:179-188, junctions.cc:76-105;
 * then sort_by_anchors + renumber_junctions, intronProspectorMerge.cc:201-223 */
int ipxo_merge_rows_fold(ipxo_state *st, const ipx_merge_rows *rows, ipx_table **out) {
    for (uint64_t r = 0; r < rows->n_rows; r++) {
        const char *rsites = rows->site_strings[rows->sites[r]];
        char rstrand = rows->strand[r];
        ojunc *j = tab_find(st, rows->chrom[r], rows->intron_start[r], rows->intron_end[r]);
        if (j == NULL) {
            j = tab_insert(st, rows->chrom[r], rows->intron_start[r], rows->intron_end[r]);
            j->anchor_start = rows->anchor_start[r];
            j->anchor_end = rows->anchor_end[r];
            j->strand = rstrand;
            snprintf(j->sites, sizeof(j->sites), "%s", rsites);
            for (int c = 0; c < 3; c++) j->counts[c] = rows->counts[3 * r + c];
            j->confidence = rows->confidence[r];
            continue;
        }
        int update_strand = 0;
        if (rstrand != '.' && j->strand == '.') {
            update_strand = 1;
        } else if (rstrand == '.' && j->strand != '.') {
            /* keep */
        } else if (rstrand != j->strand) {
            char d1[128], d2[128];
            describe(st, rows->chrom[r], rows->intron_start[r], rows->intron_end[r], rstrand, rsites, d1, sizeof(d1));
            describe(st, j->chrom, j->intron_start, j->intron_end, j->strand, j->sites, d2, sizeof(d2));
            sb_printf(&st->log, "intron records have different strand:%s with %s\n", d1, d2);
            continue;
        } else if (strcmp(rsites, j->sites) != 0) {
            char d1[128], d2[128];
            describe(st, rows->chrom[r], rows->intron_start[r], rows->intron_end[r], rstrand, rsites, d1, sizeof(d1));
            describe(st, j->chrom, j->intron_start, j->intron_end, j->strand, j->sites, d2, sizeof(d2));
            st->err.n = 0;
            sb_printf(&st->err, "intron records have different splice sites:%s with %s", d1, d2);
            return IPX_ERR_MERGE;
        }
        if (rows->anchor_start[r] < j->anchor_start) j->anchor_start = rows->anchor_start[r];
        if (rows->anchor_end[r] > j->anchor_end) j->anchor_end = rows->anchor_end[r];
        for (int c = 0; c < 3; c++) j->counts[c] += rows->counts[3 * r + c];
        /* std::max(j1.confidence, confidence): returns the first argument unless it is less than the second */
        j->confidence = (rows->confidence[r] < j->confidence) ? j->confidence : rows->confidence[r];
        if (update_strand) {
            j->strand = rstrand;
            snprintf(j->sites, sizeof(j->sites), "%s", rsites);
        }
    }
    size_t n;
    ojunc **v = sorted_junctions(st, &n);
    otable *o = table_alloc(n, 0);
    for (size_t i = 0; i < n; i++) {
        ojunc *j = v[i];
        o->chrom[i] = j->chrom;
        o->i0[i] = j->intron_start;
        o->i1[i] = j->intron_end;
        o->a0[i] = j->anchor_start;
        o->a1[i] = j->anchor_end;
        o->strand[i] = j->strand;
        memcpy(o->sites + 8 * i, j->sites, 8);
        for (int c = 0; c < 3; c++) o->counts[3 * i + c] = j->counts[c];
        o->conf[i] = j->confidence;
    }
    /* ijunc = position in anchor order.  The reference uses std::sort, whose order among rows with equal
     * (chrom, anchor_start, anchor_end, strand) is unspecified; here ties keep key order. */
    size_t *order = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) order[i] = i;
    /* insertion-free stable merge sort on indices */
    {
        size_t *tmp = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
        for (size_t w = 1; w < n; w *= 2) {
            for (size_t lo = 0; lo < n; lo += 2 * w) {
                size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
                size_t a = lo, b = mid, k = lo;
                while (a < mid && b < hi) {
                    size_t x = order[a], y = order[b];
                    int lt; /* is y strictly before x? */
                    if (o->chrom[x] != o->chrom[y]) lt = 0; /* key order already groups chromosomes by name */
                    else if (o->a0[y] != o->a0[x]) lt = o->a0[y] < o->a0[x];
                    else if (o->a1[y] != o->a1[x]) lt = o->a1[y] < o->a1[x];
                    else lt = (unsigned char)o->strand[y] < (unsigned char)o->strand[x];
                    tmp[k++] = lt ? order[b++] : order[a++];
                }
                while (a < mid) tmp[k++] = order[a++];
                while (b < hi) tmp[k++] = order[b++];
            }
            memcpy(order, tmp, sizeof(size_t) * n);
        }
        free(tmp);
    }
    for (size_t i = 0; i < n; i++) o->ijunc[order[i]] = (uint32_t)i;
    free(order);
    free(v);
    tab_clear(st);
    *out = &o->t;
    return IPX_OK;
}

/* ---------------------------------------------------------------- formatters */
static const char *bed_color(const char *sites) {
    /* getBedColor, junctions.cc:167-186 (the second GT/AG test is unreachable, so GC/AG is "unknown") */
    if (sites[0] == 0) return "64,128,64";
    if (strcmp(sites, "GT/AG") == 0) return "0,128,0";
    if (strcmp(sites, "AT/AC") == 0 || strcmp(sites, "AT/AG") == 0 || strcmp(sites, "GT/AC") == 0) return "0,0,128";
    return "128,0,0";
}

typedef struct {
    const ipx_table *t;
    const char *const *names;
    int by_anchor;
} sortctx;
static sortctx g_ctx;

static int cmp_rows(const void *pa, const void *pb) {
    uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
    const ipx_table *t = g_ctx.t;
    if (t->chrom[a] != t->chrom[b]) {
        int c = strcmp(g_ctx.names[t->chrom[a]], g_ctx.names[t->chrom[b]]);
        if (c) return c;
    }
    const uint32_t *s = g_ctx.by_anchor ? t->anchor_start : t->intron_start;
    const uint32_t *e = g_ctx.by_anchor ? t->anchor_end : t->intron_end;
    if (s[a] != s[b]) return s[a] < s[b] ? -1 : 1;
    if (e[a] != e[b]) return e[a] < e[b] ? -1 : 1;
    if (t->strand[a] != t->strand[b]) return (unsigned char)t->strand[a] < (unsigned char)t->strand[b] ? -1 : 1;
    return a < b ? -1 : (a > b); /* ties: key order (the reference's std::sort leaves this unspecified) */
}

int ipxo_format_tsv_header(ipx_text *out) {
    /* print_junction_call_header, junctions.cc:247-252 */
    const char *h = "chrom\tintron_start\tintron_end\tstrand\tuniq_mapped_count\tmulti_mapped_count\t"
                    "unsure_mapped_count\tmax_left_overhang\tmax_right_overhang\tconfidence\tsplice_sites\tname\n";
    out->data = strdup(h);
    out->len = strlen(h);
    return IPX_OK;
}

/* print_anchor_bed / print_intron_bed / print_intron_call_tsv, junctions.cc:188-268 */
int ipxo_format(const ipx_table *t, const char *const *names, int format, float min_confidence, uint64_t begin,
                uint64_t end, ipx_text *out) {
    if (end > t->n_junctions || begin > end) return IPX_ERR_ARG;
    uint64_t n = end - begin;
    uint64_t *idx = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
    for (uint64_t i = 0; i < n; i++) idx[i] = begin + i;
    g_ctx.t = t;
    g_ctx.names = names;
    g_ctx.by_anchor = (format == IPX_FMT_JUNCTION_BED);
    qsort(idx, n, sizeof(uint64_t), cmp_rows);
    sbuf b = {0, 0, 0};
    sb_reserve(&b, 16);
    b.s[0] = 0;
    for (uint64_t k = 0; k < n; k++) {
        uint64_t i = idx[k];
        if (!(t->confidence[i] >= min_confidence)) continue;
        const char *chrom = names[t->chrom[i]];
        const char *sites = t->splice_sites + 8 * i;
        uint64_t total = t->read_counts[3 * i] + t->read_counts[3 * i + 1] + t->read_counts[3 * i + 2];
        unsigned score = total <= 1000 ? (unsigned)total : 1000;
        uint32_t a0 = t->anchor_start[i], a1 = t->anchor_end[i], i0 = t->intron_start[i], i1 = t->intron_end[i];
        if (format == IPX_FMT_JUNCTION_BED) {
            sb_printf(&b, "%s\t%u\t%u\tsj%u", chrom, a0, a1, t->ijunc[i]);
            if (sites[0]) sb_printf(&b, "_%s", sites);
            sb_printf(&b, "\t%u\t%c\t%u\t%u\t%s\t2\t%u,%u\t0,%u\n", score, t->strand[i], a0, a1, bed_color(sites),
                      (uint32_t)(i0 - a0), (uint32_t)(a1 - i1), (uint32_t)(i1 - a0));
        } else if (format == IPX_FMT_INTRON_BED9 || format == IPX_FMT_INTRON_BED6) {
            sb_printf(&b, "%s\t%u\t%u\tsj%u", chrom, i0, i1, t->ijunc[i]);
            if (sites[0]) sb_printf(&b, "_%s", sites);
            sb_printf(&b, "\t%u\t%c", score, t->strand[i]);
            if (format == IPX_FMT_INTRON_BED9) sb_printf(&b, "\t%u\t%u\t%s", i0, i1, bed_color(sites));
            sb_printf(&b, "\n");
        } else if (format == IPX_FMT_INTRON_TSV) {
            sb_printf(&b, "%s\t%u\t%u\t%c\t%llu\t%llu\t%llu\t%u\t%u\t%.3g\t%s\tsj%u\n", chrom, i0, i1, t->strand[i],
                      (unsigned long long)t->read_counts[3 * i], (unsigned long long)t->read_counts[3 * i + 1],
                      (unsigned long long)t->read_counts[3 * i + 2], (uint32_t)(i0 - a0), (uint32_t)(a1 - i1),
                      (double)t->confidence[i], sites, t->ijunc[i]);
        } else {
            free(idx);
            free(b.s);
            return IPX_ERR_ARG;
        }
    }
    free(idx);
    out->data = b.s;
    out->len = b.n;
    return IPX_OK;
}

void ipxo_text_free(ipx_text *t) {
    if (t && t->data) free(t->data);
    if (t) {
        t->data = NULL;
        t->len = 0;
    }
}
