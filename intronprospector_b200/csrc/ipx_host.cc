/* ipx_host.cc -- host finalizer: float32 confidence and the text formats.
 *
 * These are J-sized (junction-count-sized) steps that must reproduce the reference byte for byte:
 *   confidence   Junction::calculate_confidence (junctions.cc:54-70) uses float32 arithmetic and the
 *                host libm's log2f; a device log2f differs in the last ulp often enough to flip a
 *                "%.3g" rounding or the >= min_confidence gate, so the device only produces the ordered
 *                run-length histogram of read offsets and the float is computed here, with the same
 *                libm the reference links.  This file is compiled with -ffp-contract=off.
 *   sorting      JunctionVector::sort_by_anchors / sort_by_introns (junctions.hh:205-222) call std::sort,
 *                which is not stable; ties between different introns with equal anchors keep whatever
 *                order libstdc++'s introsort gives on key-ordered input.  Calling the same std::sort on
 *                the same initial order with an equivalent comparator reproduces it exactly.
 *   text         Junction::print_* (junctions.cc:188-268).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ipx.h"
#include "ipx_host.h"

namespace ipx {

/* calculate_confidence (junctions.cc:54-70) from run lengths in ascending-offset order */
static float confidence_from_runs(const uint32_t *cnt, uint64_t n_runs) {
    size_t total = 0;
    for (uint64_t i = 0; i < n_runs; i++) total += cnt[i];
    float summation = 0.0f;
    for (uint64_t i = 0; i < n_runs; i++) {
        float p = float(cnt[i]) / total;
        if (p != 0.0f) {
            summation += p * log2f(p);
        }
    }
    return (summation == 0.0f) ? 0.0f : -summation;
}

void compute_confidences(uint64_t n, const uint64_t *read_counts, const uint64_t *hist_off, const uint32_t *hist_cnt,
                         float *out) {
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            /* lazy_get_confidence (junctions.cc:44-51) */
            uint64_t total = read_counts[3 * i] + read_counts[3 * i + 1] + read_counts[3 * i + 2];
            out[i] = (total == 0) ? 0.0f : confidence_from_runs(hist_cnt + hist_off[i], hist_off[i + 1] - hist_off[i]);
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = (n < 65536 || hw < 2) ? 1u : std::min(hw, 32u);
    if (nt == 1) {
        work(0, n);
        return;
    }
    /* split by histogram volume so that threads get equal work */
    std::vector<std::thread> th;
    uint64_t total_runs = hist_off[n];
    uint64_t lo = 0;
    for (unsigned k = 0; k < nt; k++) {
        uint64_t target = total_runs * (k + 1) / nt;
        uint64_t hi = (k + 1 == nt) ? n : (uint64_t)(std::upper_bound(hist_off + lo, hist_off + n, target) - hist_off);
        if (hi > n) hi = n;
        if (hi < lo) hi = lo;
        th.emplace_back(work, lo, hi);
        lo = hi;
    }
    for (auto &t : th) t.join();
}

/* getBedColor (junctions.cc:167-186); the second "GT/AG" test there is unreachable */
static const char *bed_color(const char *sites) {
    if (sites[0] == 0) return "64,128,64";
    if (strcmp(sites, "GT/AG") == 0) return "0,128,0";
    if (strcmp(sites, "AT/AC") == 0 || strcmp(sites, "AT/AG") == 0 || strcmp(sites, "GT/AC") == 0) return "0,0,128";
    return "128,0,0";
}

static inline void put_u64(std::string &s, uint64_t v) {
    char tmp[24];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) s.push_back(tmp[--n]);
}

void sort_rows(const ipx_table *t, const char *const *names, bool by_anchor, uint64_t begin, uint64_t end,
               std::vector<uint64_t> &idx) {
    idx.resize(end - begin);
    for (uint64_t i = begin; i < end; i++) idx[i - begin] = i;
    const uint32_t *s = by_anchor ? t->anchor_start : t->intron_start;
    const uint32_t *e = by_anchor ? t->anchor_end : t->intron_end;
    /* junctions_lt / introns_lt (junctions.cc:108-159) */
    auto lt = [&](uint64_t a, uint64_t b) {
        if (t->chrom[a] != t->chrom[b]) {
            int c = strcmp(names[t->chrom[a]], names[t->chrom[b]]);
            if (c < 0) return true;
            if (c > 0) return false;
        }
        if (s[a] < s[b]) return true;
        if (s[a] > s[b]) return false;
        if (e[a] < e[b]) return true;
        if (e[a] > e[b]) return false;
        return t->strand[a] < t->strand[b];
    };
    std::sort(idx.begin(), idx.end(), lt);
}

int format_rows(const ipx_table *t, const char *const *names, int format, float min_confidence, uint64_t begin,
                uint64_t end, std::string &out) {
    if (begin > end || end > t->n_junctions) return IPX_ERR_ARG;
    if (format < IPX_FMT_JUNCTION_BED || format > IPX_FMT_INTRON_TSV) return IPX_ERR_ARG;
    std::vector<uint64_t> idx;
    sort_rows(t, names, format == IPX_FMT_JUNCTION_BED, begin, end, idx);
    out.clear();
    out.reserve(idx.size() * 96);
    char conf[64];
    for (uint64_t i : idx) {
        if (!(t->confidence[i] >= min_confidence)) continue;
        const char *chrom = names[t->chrom[i]];
        const char *sites = (t->site_strings && t->site_ids) ? t->site_strings[t->site_ids[i]] : t->splice_sites + 8 * i;
        const uint64_t total = t->read_counts[3 * i] + t->read_counts[3 * i + 1] + t->read_counts[3 * i + 2];
        const uint64_t score = total <= 1000 ? total : 1000; /* read_count_to_bed_score (junctions.cc:162-164) */
        const uint32_t a0 = t->anchor_start[i], a1 = t->anchor_end[i], i0 = t->intron_start[i], i1 = t->intron_end[i];
        if (format == IPX_FMT_JUNCTION_BED) {
            out += chrom; out += '\t'; put_u64(out, a0); out += '\t'; put_u64(out, a1); out += "\tsj"; put_u64(out, t->ijunc[i]);
            if (sites[0]) { out += '_'; out += sites; }
            out += '\t'; put_u64(out, score); out += '\t'; out += t->strand[i];
            out += '\t'; put_u64(out, a0); out += '\t'; put_u64(out, a1); out += '\t'; out += bed_color(sites);
            out += "\t2\t"; put_u64(out, (uint32_t)(i0 - a0)); out += ','; put_u64(out, (uint32_t)(a1 - i1));
            out += "\t0,"; put_u64(out, (uint32_t)(i1 - a0)); out += '\n';
        } else if (format == IPX_FMT_INTRON_TSV) {
            out += chrom; out += '\t'; put_u64(out, i0); out += '\t'; put_u64(out, i1); out += '\t'; out += t->strand[i];
            out += '\t'; put_u64(out, t->read_counts[3 * i]); out += '\t'; put_u64(out, t->read_counts[3 * i + 1]);
            out += '\t'; put_u64(out, t->read_counts[3 * i + 2]);
            out += '\t'; put_u64(out, (uint32_t)(i0 - a0)); out += '\t'; put_u64(out, (uint32_t)(a1 - i1));
            /* ostream << float with precision(3) (junctions.cc:259) prints like printf("%.3g") */
            snprintf(conf, sizeof(conf), "%.3g", (double)t->confidence[i]);
            out += '\t'; out += conf; out += '\t'; out += sites; out += "\tsj"; put_u64(out, t->ijunc[i]); out += '\n';
        } else {
            out += chrom; out += '\t'; put_u64(out, i0); out += '\t'; put_u64(out, i1); out += "\tsj"; put_u64(out, t->ijunc[i]);
            if (sites[0]) { out += '_'; out += sites; }
            out += '\t'; put_u64(out, score); out += '\t'; out += t->strand[i];
            if (format == IPX_FMT_INTRON_BED9) {
                out += '\t'; put_u64(out, i0); out += '\t'; put_u64(out, i1); out += '\t'; out += bed_color(sites);
            }
            out += '\n';
        }
    }
    return IPX_OK;
}

const char *tsv_header() {
    /* print_junction_call_header (junctions.cc:247-252) */
    return "chrom\tintron_start\tintron_end\tstrand\tuniq_mapped_count\tmulti_mapped_count\tunsure_mapped_count\t"
           "max_left_overhang\tmax_right_overhang\tconfidence\tsplice_sites\tname\n";
}

}  // namespace ipx
