/* record_source.cc -- see record_source.h */
#include "record_source.h"
#include <string.h>

namespace ipx {

static size_t aux_size(const uint8_t *t, const uint8_t *end) {
    if (t >= end) return 0;
    switch (*t) {
        case 'A': case 'c': case 'C': return 2;
        case 's': case 'S': return 3;
        case 'i': case 'I': case 'f': return 5;
        case 'd': return 9;
        case 'Z': case 'H': {
            const uint8_t *z = (const uint8_t *)memchr(t + 1, 0, end - (t + 1));
            return z ? (size_t)(z - t) + 1 : 0;
        }
        case 'B': {
            if (end - t < 6) return 0;
            uint32_t n;
            memcpy(&n, t + 2, 4);
            size_t es = (t[1] == 'c' || t[1] == 'C') ? 1 : ((t[1] == 's' || t[1] == 'S') ? 2 : 4);
            return 6 + (size_t)n * es;
        }
        default: return 0;
    }
}


void scan_tags(const bam1_t *b, uint8_t *nh_class, uint8_t *xs) {
    scan_tag_range(bam_get_aux(b), b->data + b->l_data, nh_class, xs);
}

const uint8_t *find_aux(const uint8_t *a, const uint8_t *end, char t0, char t1) {
    while (a + 2 < end) {
        const uint8_t *t = a + 2;
        const size_t vs = aux_size(t, end);
        if (vs == 0 || t + vs > end) return nullptr;
        if (a[0] == (uint8_t)t0 && a[1] == (uint8_t)t1) return t;
        a = t + vs;
    }
    return nullptr;
}

void scan_tag_range(const uint8_t *a, const uint8_t *end, uint8_t *nh_class, uint8_t *xs) {
    *nh_class = IPX_NH_NONE;
    *xs = IPX_XS_ABSENT;
    bool seen_nh = false, seen_xs = false;
    while (a + 2 < end && !(seen_nh && seen_xs)) {
        const uint8_t *t = a + 2;
        size_t vs = aux_size(t, end);
        if (vs == 0 || t + vs > end) break;
        if (a[0] == 'N' && a[1] == 'H' && !seen_nh) {
            seen_nh = true;
            int64_t v = 0;
            switch (*t) {
                case 'c': v = (int8_t)t[1]; break;
                case 'C': v = t[1]; break;
                case 's': { int16_t x; memcpy(&x, t + 1, 2); v = x; break; }
                case 'S': { uint16_t x; memcpy(&x, t + 1, 2); v = x; break; }
                case 'i': { int32_t x; memcpy(&x, t + 1, 4); v = x; break; }
                case 'I': { uint32_t x; memcpy(&x, t + 1, 4); v = x; break; }
                default: v = 0;
            }
            /* the reference stores bam_aux2i's result in an int */
            int iv = (int)v;
            *nh_class = iv <= 0 ? IPX_NH_NONE : (iv == 1 ? IPX_NH_ONE : IPX_NH_MANY);
        } else if (a[0] == 'X' && a[1] == 'S' && !seen_xs) {
            seen_xs = true;
            *xs = (*t == 'A' && t[1] != 0 && t[1] != IPX_XS_OTHER) ? t[1] : IPX_XS_OTHER;
        }
        a = t + vs;
    }
}

void ColumnarBatch::add(const bam1_t *b, int32_t mapped_tid) {
    uint8_t nh, x;
    scan_tags(b, &nh, &x);
    if (run_tid.empty() || run_tid.back() != mapped_tid) {
        run_tid.push_back(mapped_tid);
        run_start.push_back((uint64_t)tid.size());
    }
    if (b->core.n_cigar > 0xFFFFu) compact_ok = false;
    n_cigar_ops.push_back((uint16_t)b->core.n_cigar);
    tid.push_back(mapped_tid);
    pos.push_back((int32_t)b->core.pos);
    flag.push_back(b->core.flag);
    nh_class.push_back(nh);
    xs.push_back(x);
    const uint32_t *c = bam_get_cigar(b);
    cigar.insert(cigar.end(), c, c + b->core.n_cigar);
    cigar_off.push_back((uint32_t)cigar.size());
}

void ColumnarBatch::add_raw(const uint8_t *rec, uint32_t len, int32_t mapped_tid) {
    int32_t p, l_seq;
    uint16_t n_cigar, fl;
    memcpy(&p, rec + 4, 4);
    memcpy(&n_cigar, rec + 12, 2);
    memcpy(&fl, rec + 14, 2);
    memcpy(&l_seq, rec + 16, 4);
    const uint8_t *cig = rec + 32 + rec[8];
    /* BAM is little-endian, and so is every host this runs on; the words may sit at any address in the buffer */
    bool has_intron = false;
    if (n_cigar <= 16) {
        for (unsigned k = 0; k < n_cigar; k++) { /* short reads: cheaper than growing and copying */
            uint32_t w;
            memcpy(&w, cig + 4 * k, 4);
            has_intron |= (w & 15u) == 3u;
            cigar.push_back(w);
        }
    } else {
        const size_t at = cigar.size();
        cigar.resize(at + n_cigar);
        memcpy(cigar.data() + at, cig, 4 * (size_t)n_cigar);
        has_intron = true; /* not worth a second pass: look at the tags */
    }
    cigar_off.push_back((uint32_t)cigar.size());
    /* NH and XS only ever matter for a read that yields a junction candidate, and that takes an N operation
     * (ReadJunctionExtractor, junctions_extractor.cc:208-229): most reads of an RNA-Seq file have none, and their
     * tags -- a dozen per record in aligner output -- need not be walked */
    uint8_t nh = IPX_NH_NONE, x = IPX_XS_ABSENT;
    if (has_intron) {
        const uint8_t *aux = cig + 4 * (size_t)n_cigar + (size_t)((l_seq + 1) / 2) + (size_t)l_seq;
        scan_tag_range(aux, rec + len, &nh, &x);
    }
    if (run_tid.empty() || run_tid.back() != mapped_tid) {
        run_tid.push_back(mapped_tid);
        run_start.push_back((uint64_t)tid.size());
    }
    n_cigar_ops.push_back(n_cigar);
    tid.push_back(mapped_tid);
    pos.push_back(p);
    flag.push_back(fl);
    nh_class.push_back(nh);
    xs.push_back(x);
}

ipx_batch ColumnarBatch::view() const {
    ipx_batch v;
    memset(&v, 0, sizeof(v));
    v.n_records = tid.size();
    v.n_cigar = cigar.size();
    v.seq_base = seq_base;
    v.tid = tid.data();
    v.pos = pos.data();
    v.flag = flag.data();
    v.nh_class = nh_class.data();
    v.xs = xs.data();
    v.cigar_off = cigar_off.data();
    v.cigar = cigar.data();
    return v;
}

ipx_batch_compact ColumnarBatch::view_compact() const {
    ipx_batch_compact v;
    memset(&v, 0, sizeof(v));
    v.n_records = tid.size();
    v.n_cigar = cigar.size();
    v.seq_base = seq_base;
    v.pos = pos.data();
    v.flag = flag.data();
    v.nh_class = nh_class.data();
    v.xs = xs.data();
    v.n_cigar_ops = n_cigar_ops.data();
    v.cigar = cigar.data();
    v.n_tid_runs = run_tid.size();
    v.run_tid = run_tid.data();
    v.run_start = run_start.data();
    return v;
}

}  // namespace ipx
