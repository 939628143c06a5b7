/* synth.cc -- see synth.h.  Deterministic for a given spec (independent of the thread count). */
#include "synth.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) { next(); next(); }
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
    uint32_t geometric(double mean) {
        double u = uniform();
        if (u < 1e-12) u = 1e-12;
        uint32_t k = (uint32_t)(log(u) / log(1.0 - 1.0 / mean)) + 1;
        return k;
    }
};

struct Gene {
    int32_t contig;
    char strand;
    int32_t start;                 /* genomic start of first exon */
    std::vector<int32_t> exon_len; /* n_introns + 1 */
    std::vector<int32_t> intron_len;
    std::vector<uint8_t> motif;    /* per intron: 0 GT/AG, 1 GC/AG, 2 AT/AC, 3 non-canonical */
    int32_t tx_len;
    int32_t span;
    uint64_t n_reads;
};

struct Out {
    std::vector<int32_t> tid, pos;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> nh, xs;
    std::vector<uint32_t> ncig; /* per record */
    std::vector<uint32_t> cigar;
};

inline uint32_t cg(uint32_t len, uint32_t op) { return (len << 4) | op; }
enum { M = 0, I = 1, D = 2, N = 3, S = 4, H = 5, EQ = 7, X = 8 };

}  // namespace

struct ipx_synth_data {
    ipx_synth_spec spec;
    std::vector<Gene> genes;
    std::vector<std::string> contig_names;
    std::vector<uint64_t> contig_len;
    Out out;
    std::vector<uint32_t> cigar_off;
};

static void build_model(ipx_synth_data *d) {
    const ipx_synth_spec &sp = d->spec;
    Rng rng(sp.model_seed ^ 0xA5A5A5A5ull);
    const int nc = sp.n_contigs;
    const uint64_t lc = sp.genome_len / (uint64_t)nc;
    d->contig_names.resize(nc);
    d->contig_len.assign(nc, lc);
    for (int i = 0; i < nc; i++) d->contig_names[i] = "chr" + std::to_string(i + 1);
    const int per = std::max(1, sp.n_genes / nc);
    const int64_t slot = (int64_t)(lc / (uint64_t)per);
    const bool longreads = sp.kind == IPX_SYNTH_LONG;
    d->genes.clear();
    for (int c = 0; c < nc; c++) {
        for (int g = 0; g < per; g++) {
            Gene ge;
            ge.contig = c;
            ge.strand = (rng.next() & 1) ? '+' : '-';
            const int64_t budget = slot - 2000;
            int n_introns = 1 + (int)rng.below(16);
            int64_t used = 0;
            ge.tx_len = 0;
            for (int k = 0; k <= n_introns; k++) {
                int32_t el = longreads ? (int32_t)(200 + rng.below(2800)) : (int32_t)(60 + rng.below(240));
                if (used + el > budget) break;
                ge.exon_len.push_back(el);
                used += el;
                ge.tx_len += el;
                if (k == n_introns) break;
                /* log-uniform 70 .. 100000 */
                double u = rng.uniform();
                int32_t il = (int32_t)(70.0 * exp(u * log(100000.0 / 70.0)));
                if (used + il + 400 > budget) break;
                ge.intron_len.push_back(il);
                used += il;
                uint32_t m = rng.below(1000);
                ge.motif.push_back(m < 980 ? 0 : (m < 990 ? 1 : (m < 995 ? 2 : 3)));
            }
            while (ge.exon_len.size() > ge.intron_len.size() + 1) {
                ge.tx_len -= ge.exon_len.back();
                ge.exon_len.pop_back();
            }
            while (ge.intron_len.size() + 1 > ge.exon_len.size()) {
                ge.intron_len.pop_back();
                ge.motif.pop_back();
            }
            if (ge.exon_len.empty()) {
                ge.exon_len.push_back(300);
                ge.tx_len = 300;
            }
            ge.span = 0;
            for (int32_t e : ge.exon_len) ge.span += e;
            for (int32_t il : ge.intron_len) ge.span += il;
            ge.start = (int32_t)((int64_t)g * slot + 1000);
            ge.n_reads = 0;
            d->genes.push_back(ge);
        }
    }
    /* expression: log-normal weights -> read counts that sum to n_records exactly */
    std::vector<double> w(d->genes.size());
    double tot = 0;
    for (size_t i = 0; i < w.size(); i++) {
        w[i] = exp(1.5 * rng.normal());
        tot += w[i];
    }
    double acc = 0;
    uint64_t given = 0;
    for (size_t i = 0; i < w.size(); i++) {
        acc += w[i];
        uint64_t upto = (uint64_t)((acc / tot) * (double)sp.n_records + 0.5);
        if (i + 1 == w.size()) upto = sp.n_records;
        if (upto < given) upto = given;
        d->genes[i].n_reads = upto - given;
        given = upto;
    }
}

/* genomic position of transcript coordinate t */
static inline int32_t tx_to_genome(const Gene &g, int32_t t, size_t *exon_idx, int32_t *within) {
    int32_t gp = g.start;
    for (size_t e = 0; e < g.exon_len.size(); e++) {
        if (t < g.exon_len[e] || e + 1 == g.exon_len.size()) {
            *exon_idx = e;
            *within = t;
            return gp + t;
        }
        t -= g.exon_len[e];
        gp += g.exon_len[e] + g.intron_len[e];
    }
    *exon_idx = 0;
    *within = 0;
    return g.start;
}

struct Desc {
    int32_t pos;
    int32_t t;      /* transcript coordinate, -1 = unspliced read placed at pos */
    uint32_t salt;
};

static void emit_short(const ipx_synth_spec &sp, const Gene &g, const Desc &ds, Rng &rng, Out &o) {
    const uint32_t RL = 150;
    uint16_t flag;
    const uint32_t f4 = rng.below(4);
    if (sp.kind == IPX_SYNTH_SHORT_STRANDED) {
        flag = (g.strand == '+') ? ((f4 & 1) ? 83 : 163) : ((f4 & 1) ? 99 : 147);
    } else {
        static const uint16_t fl[4] = {99, 147, 83, 163};
        flag = fl[f4];
    }
    const uint32_t r = rng.below(1000);
    if (r < 20) flag |= 0x100; else if (r < 25) flag |= 0x400;
    const uint32_t nhr = rng.below(100);
    const uint8_t nh = nhr < 90 ? 1 : 2;
    const size_t c0 = o.cigar.size();
    bool spliced = false;
    int32_t pos = ds.pos;
    if (ds.t < 0) {
        o.cigar.push_back(cg(RL, M));
    } else {
        size_t e;
        int32_t within;
        pos = tx_to_genome(g, ds.t, &e, &within);
        uint32_t remaining = RL;
        while (remaining > 0) {
            uint32_t avail = (uint32_t)(g.exon_len[e] - within);
            uint32_t take = (e + 1 == g.exon_len.size()) ? remaining : std::min(avail, remaining);
            o.cigar.push_back(cg(take, M));
            remaining -= take;
            if (remaining > 0) {
                o.cigar.push_back(cg((uint32_t)g.intron_len[e], N));
                spliced = true;
                e++;
                within = 0;
            }
        }
        const uint32_t v = rng.below(100);
        if (v < 2) {
            /* soft clip at the start */
            uint32_t first = o.cigar[c0] >> 4;
            uint32_t k = 1 + rng.below(20);
            if (first > k + 1) {
                o.cigar[c0] = cg(first - k, M);
                o.cigar.insert(o.cigar.begin() + (long)c0, cg(k, S));
                pos += (int32_t)k;
            }
        } else if (v < 3) {
            /* small indel inside the last aligned segment */
            uint32_t last = o.cigar.back() >> 4;
            if (last >= 30) {
                uint32_t k = 1 + rng.below(3), a = 10 + rng.below(last - 20);
                o.cigar.back() = cg(a, M);
                o.cigar.push_back(cg(k, (rng.next() & 1) ? I : D));
                o.cigar.push_back(cg(last - a, M));
            }
        }
    }
    o.tid.push_back(g.contig);
    o.pos.push_back(pos);
    o.flag.push_back(flag);
    o.nh.push_back(nh);
    o.xs.push_back(spliced ? (uint8_t)g.strand : 0);
    o.ncig.push_back((uint32_t)(o.cigar.size() - c0));
}

static void emit_long(const Gene &g, const Desc &ds, Rng &rng, Out &o) {
    uint16_t flag = (rng.next() & 1) ? 16 : 0;
    if (rng.below(100) < 5) flag |= 0x800;
    const size_t c0 = o.cigar.size();
    /* log-normal read length around 10 kb */
    double L = exp(log(10000.0) + 0.5 * rng.normal());
    uint32_t remaining = (uint32_t)std::min(std::max(L, 300.0), 60000.0);
    size_t e;
    int32_t within;
    int32_t pos = tx_to_genome(g, ds.t, &e, &within);
    const uint32_t clip5 = rng.below(100);
    if (clip5 < 30) o.cigar.push_back(cg(1 + rng.below(80), S)); else if (clip5 < 40) o.cigar.push_back(cg(1 + rng.below(80), H));
    bool first_op = true;
    while (remaining > 0 && e < g.exon_len.size()) {
        uint32_t avail = (uint32_t)(g.exon_len[e] - within);
        uint32_t seg = std::min(avail, remaining);
        /* one op per ~12 reference bases inside the exon segment */
        uint32_t left = seg;
        while (left > 0) {
            uint32_t run = std::min(left, rng.geometric(12.0));
            o.cigar.push_back(cg(run, EQ));
            left -= run;
            first_op = false;
            if (left == 0) break;
            const uint32_t k = rng.below(4);
            if (k < 2) {
                o.cigar.push_back(cg(1, X));
                left -= 1;
            } else if (k == 2) {
                o.cigar.push_back(cg(rng.geometric(2.0), I));
            } else {
                uint32_t dl = std::min(left, rng.geometric(2.0));
                /* keep an aligned base after the deletion inside this exon */
                if (dl >= left) dl = left > 1 ? left - 1 : 0;
                if (dl) {
                    o.cigar.push_back(cg(dl, D));
                    left -= dl;
                }
            }
        }
        remaining -= seg;
        if (seg == avail && e + 1 < g.exon_len.size() && remaining > 0) {
            o.cigar.push_back(cg((uint32_t)g.intron_len[e], N));
            e++;
            within = 0;
        } else {
            break;
        }
    }
    (void)first_op;
    const uint32_t clip3 = rng.below(100);
    if (clip3 < 30) o.cigar.push_back(cg(1 + rng.below(80), S)); else if (clip3 < 40) o.cigar.push_back(cg(1 + rng.below(80), H));
    o.tid.push_back(g.contig);
    o.pos.push_back(pos);
    o.flag.push_back(flag);
    o.nh.push_back(0);
    o.xs.push_back(0);
    o.ncig.push_back((uint32_t)(o.cigar.size() - c0));
}

static void generate_gene(const ipx_synth_spec &sp, const Gene &g, size_t gene_index, Out &o) {
    if (g.n_reads == 0) return;
    Rng rng(sp.seed * 1000003ull + gene_index * 7919ull + 17ull);
    std::vector<Desc> ds((size_t)g.n_reads);
    const bool longreads = sp.kind == IPX_SYNTH_LONG;
    for (size_t i = 0; i < ds.size(); i++) {
        Desc d;
        d.salt = (uint32_t)rng.next();
        if (!longreads && rng.below(100) < 70) {
            d.t = -1;
            d.pos = g.start + (int32_t)rng.below((uint32_t)std::max(1, g.span - 150));
        } else {
            int32_t room = longreads ? std::max(1, g.tx_len - 200) : std::max(1, g.tx_len - 150);
            d.t = (int32_t)rng.below((uint32_t)room);
            size_t e;
            int32_t w;
            d.pos = tx_to_genome(g, d.t, &e, &w);
        }
        ds[i] = d;
    }
    std::sort(ds.begin(), ds.end(), [](const Desc &a, const Desc &b) { return a.pos != b.pos ? a.pos < b.pos : a.salt < b.salt; });
    for (const Desc &d : ds) {
        Rng r2(((uint64_t)d.salt << 20) ^ sp.seed ^ (gene_index * 0x9E37ull));
        if (longreads) emit_long(g, d, r2, o); else emit_short(sp, g, d, r2, o);
    }
    /* soft clips move pos forward by a few bases: restore order inside the gene */
    (void)0;
}

extern "C" ipx_synth_data *ipx_synth_generate(const ipx_synth_spec *spec) {
    if (!spec || spec->n_contigs <= 0 || spec->n_genes <= 0) return nullptr;
    ipx_synth_data *d = new ipx_synth_data();
    d->spec = *spec;
    build_model(d);
    unsigned nt = spec->threads > 0 ? (unsigned)spec->threads : std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = std::min<unsigned>(nt, 64);
    /* contiguous gene groups of roughly equal read counts */
    const size_t ng = d->genes.size();
    const size_t chunks = std::min<size_t>(ng, (size_t)nt * 8);
    std::vector<size_t> bounds(chunks + 1, ng);
    {
        uint64_t acc = 0;
        size_t c = 0;
        bounds[0] = 0;
        for (size_t i = 0; i < ng && c + 1 < chunks; i++) {
            acc += d->genes[i].n_reads;
            if (acc >= (spec->n_records * (c + 1)) / chunks) bounds[++c] = i + 1;
        }
        for (size_t k = c + 1; k <= chunks; k++) bounds[k] = ng;
    }
    std::vector<Out> outs(chunks);
    std::vector<std::thread> th;
    std::vector<size_t> next_chunk(1, 0);
    std::atomic<size_t> cursor(0);
    auto work = [&]() {
        for (;;) {
            size_t c = cursor.fetch_add(1);
            if (c >= chunks) break;
            Out &o = outs[c];
            for (size_t gi = bounds[c]; gi < bounds[c + 1]; gi++) {
                size_t before = o.tid.size();
                generate_gene(d->spec, d->genes[gi], gi, o);
                /* soft clips shift pos by up to 20: re-sort this gene's records by pos (stable) */
                size_t n = o.tid.size() - before;
                bool sorted = true;
                for (size_t k = before + 1; k < o.tid.size(); k++)
                    if (o.pos[k] < o.pos[k - 1]) { sorted = false; break; }
                if (!sorted && n > 1) {
                    std::vector<uint32_t> idx(n), coff(n + 1);
                    size_t cbase = o.cigar.size();
                    for (size_t k = 0; k < n; k++) cbase -= o.ncig[before + k];
                    coff[0] = 0;
                    for (size_t k = 0; k < n; k++) { idx[k] = (uint32_t)k; coff[k + 1] = coff[k] + o.ncig[before + k]; }
                    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return o.pos[before + a] < o.pos[before + b]; });
                    std::vector<int32_t> tp(n), tt(n);
                    std::vector<uint16_t> tf(n);
                    std::vector<uint8_t> tn(n), tx(n);
                    std::vector<uint32_t> tc(n), cig(o.cigar.begin() + (long)cbase, o.cigar.end());
                    size_t w = cbase;
                    for (size_t k = 0; k < n; k++) {
                        uint32_t s = idx[k];
                        tt[k] = o.tid[before + s]; tp[k] = o.pos[before + s]; tf[k] = o.flag[before + s];
                        tn[k] = o.nh[before + s]; tx[k] = o.xs[before + s]; tc[k] = o.ncig[before + s];
                        memcpy(&o.cigar[w], &cig[coff[s]], (size_t)tc[k] * 4);
                        w += tc[k];
                    }
                    for (size_t k = 0; k < n; k++) {
                        o.tid[before + k] = tt[k]; o.pos[before + k] = tp[k]; o.flag[before + k] = tf[k];
                        o.nh[before + k] = tn[k]; o.xs[before + k] = tx[k]; o.ncig[before + k] = tc[k];
                    }
                }
            }
        }
    };
    for (unsigned t = 0; t < nt; t++) th.emplace_back(work);
    for (auto &t : th) t.join();
    /* concatenate in gene order */
    uint64_t n = 0, nc = 0;
    for (const Out &o : outs) { n += o.tid.size(); nc += o.cigar.size(); }
    Out &f = d->out;
    f.tid.resize(n); f.pos.resize(n); f.flag.resize(n); f.nh.resize(n); f.xs.resize(n); f.cigar.resize(nc);
    d->cigar_off.resize(n + 1);
    uint64_t r = 0, c = 0;
    for (Out &o : outs) {
        const size_t k = o.tid.size();
        if (k) {
            memcpy(&f.tid[r], o.tid.data(), k * 4); memcpy(&f.pos[r], o.pos.data(), k * 4); memcpy(&f.flag[r], o.flag.data(), k * 2);
            memcpy(&f.nh[r], o.nh.data(), k); memcpy(&f.xs[r], o.xs.data(), k);
            if (!o.cigar.empty()) memcpy(&f.cigar[c], o.cigar.data(), o.cigar.size() * 4);
            for (size_t i = 0; i < k; i++) {
                d->cigar_off[r + i] = (uint32_t)c;
                c += o.ncig[i];
            }
            r += k;
        }
        o = Out();
    }
    d->cigar_off[n] = (uint32_t)c;
    return d;
}

extern "C" void ipx_synth_free(ipx_synth_data *d) { delete d; }
extern "C" uint64_t ipx_synth_n_records(const ipx_synth_data *d) { return d->out.tid.size(); }
extern "C" uint64_t ipx_synth_n_cigar(const ipx_synth_data *d) { return d->out.cigar.size(); }
extern "C" int32_t ipx_synth_n_contigs(const ipx_synth_data *d) { return (int32_t)d->contig_names.size(); }
extern "C" const char *ipx_synth_contig_name(const ipx_synth_data *d, int32_t i) { return d->contig_names[i].c_str(); }
extern "C" uint64_t ipx_synth_contig_len(const ipx_synth_data *d, int32_t i) { return d->contig_len[i]; }

extern "C" void ipx_synth_copy(const ipx_synth_data *d, int32_t *tid, int32_t *pos, uint16_t *flag, uint8_t *nh_class, uint8_t *xs,
                               uint32_t *cigar_off, uint32_t *cigar) {
    const Out &o = d->out;
    const size_t n = o.tid.size();
    if (tid) memcpy(tid, o.tid.data(), n * 4);
    if (pos) memcpy(pos, o.pos.data(), n * 4);
    if (flag) memcpy(flag, o.flag.data(), n * 2);
    if (nh_class) memcpy(nh_class, o.nh.data(), n);
    if (xs) memcpy(xs, o.xs.data(), n);
    if (cigar_off) memcpy(cigar_off, d->cigar_off.data(), (n + 1) * 4);
    if (cigar) memcpy(cigar, o.cigar.data(), o.cigar.size() * 4);
}

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

extern "C" int ipx_synth_write_bam(const ipx_synth_data *d, const char *path, uint64_t begin, uint64_t end, int raw) {
    if (!raw) return -1; /* BGZF output goes through hts_tool view-bam */
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    std::string hdr_text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (size_t i = 0; i < d->contig_names.size(); i++)
        hdr_text += "@SQ\tSN:" + d->contig_names[i] + "\tLN:" + std::to_string(d->contig_len[i]) + "\n";
    std::string b;
    b.append("BAM\1", 4);
    int32_t l_text = (int32_t)hdr_text.size();
    b.append((const char *)&l_text, 4);
    b += hdr_text;
    int32_t n_ref = (int32_t)d->contig_names.size();
    b.append((const char *)&n_ref, 4);
    for (int32_t i = 0; i < n_ref; i++) {
        int32_t l_name = (int32_t)d->contig_names[i].size() + 1;
        b.append((const char *)&l_name, 4);
        b.append(d->contig_names[i].c_str(), l_name);
        int32_t l_ref = (int32_t)d->contig_len[i];
        b.append((const char *)&l_ref, 4);
    }
    const Out &o = d->out;
    if (end > o.tid.size()) end = o.tid.size();
    for (uint64_t r = begin; r < end; r++) {
        const uint32_t c0 = d->cigar_off[r], nc = d->cigar_off[r + 1] - c0;
        int64_t rlen = 0;
        for (uint32_t k = 0; k < nc; k++) {
            uint32_t op = o.cigar[c0 + k] & 15;
            if (op == M || op == D || op == N || op == EQ || op == X) rlen += o.cigar[c0 + k] >> 4;
        }
        if (rlen == 0) rlen = 1;
        uint8_t aux[16];
        int la = 0;
        if (o.nh[r]) { aux[la++] = 'N'; aux[la++] = 'H'; aux[la++] = 'C'; aux[la++] = (o.nh[r] == 1) ? 1 : 3; }
        if (o.xs[r]) { aux[la++] = 'X'; aux[la++] = 'S'; aux[la++] = 'A'; aux[la++] = o.xs[r]; }
        const uint32_t l_qname = 2;
        int32_t x[8];
        x[0] = o.tid[r];
        x[1] = o.pos[r];
        x[2] = (int32_t)(((uint32_t)reg2bin(o.pos[r], o.pos[r] + rlen) << 16) | (255u << 8) | l_qname);
        x[3] = (int32_t)(((uint32_t)o.flag[r] << 16) | nc);
        x[4] = 0;
        x[5] = -1;
        x[6] = -1;
        x[7] = 0;
        int32_t block = 32 + (int32_t)l_qname + (int32_t)nc * 4 + la;
        b.append((const char *)&block, 4);
        b.append((const char *)x, 32);
        b.append("r\0", 2);
        b.append((const char *)&o.cigar[c0], (size_t)nc * 4);
        b.append((const char *)aux, la);
        if (b.size() > (1u << 24)) {
            if (fwrite(b.data(), 1, b.size(), fp) != b.size()) { fclose(fp); return -1; }
            b.clear();
        }
    }
    if (!b.empty() && fwrite(b.data(), 1, b.size(), fp) != b.size()) { fclose(fp); return -1; }
    return fclose(fp) == 0 ? 0 : -1;
}

/* ---- the same records as an aligner writes them: BGZF-compressed BAM with read names, bases, qualities, mate fields and
 * a STAR-style tag set (NH HI AS nM, XS:A on spliced reads, jM / jI arrays) -- what the command-line tools read ---- */
static std::string bam_header_bytes(const ipx_synth_data *d) {
    std::string hdr_text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (size_t i = 0; i < d->contig_names.size(); i++)
        hdr_text += "@SQ\tSN:" + d->contig_names[i] + "\tLN:" + std::to_string(d->contig_len[i]) + "\n";
    hdr_text += "@PG\tID:synth\tPN:ipx_synth\tVN:1\tCL:seeded generator (SURVEY.md 8d)\n";
    std::string b;
    b.append("BAM\1", 4);
    int32_t l_text = (int32_t)hdr_text.size();
    b.append((const char *)&l_text, 4);
    b += hdr_text;
    int32_t n_ref = (int32_t)d->contig_names.size();
    b.append((const char *)&n_ref, 4);
    for (int32_t i = 0; i < n_ref; i++) {
        int32_t l_name = (int32_t)d->contig_names[i].size() + 1;
        b.append((const char *)&l_name, 4);
        b.append(d->contig_names[i].c_str(), l_name);
        int32_t l_ref = (int32_t)d->contig_len[i];
        b.append((const char *)&l_ref, 4);
    }
    return b;
}

static void append_full_record(const ipx_synth_data *d, uint64_t r, std::string &b) {
    const Out &o = d->out;
    const uint32_t c0 = d->cigar_off[r], nc = d->cigar_off[r + 1] - c0;
    int64_t rlen = 0, qlen = 0;
    bool spliced = false;
    for (uint32_t k = 0; k < nc; k++) {
        const uint32_t op = o.cigar[c0 + k] & 15, len = o.cigar[c0 + k] >> 4;
        if (op == M || op == D || op == N || op == EQ || op == X) rlen += len;
        if (op == M || op == I || op == S || op == EQ || op == X) qlen += len;
        spliced |= op == N;
    }
    if (rlen == 0) rlen = 1;
    if (qlen > 400) qlen = 400; /* long reads: keep the file a sensible size; nothing on the path reads SEQ */
    Rng rng(0xA11CEull ^ (r * 0x9E3779B97F4A7C15ull));
    char qname[40];
    const int lq = snprintf(qname, sizeof(qname), "SYN%llu.%llu", (unsigned long long)(d->spec.seed % 100000), (unsigned long long)(r >> 1)) + 1;
    uint8_t aux[96];
    int la = 0;
    auto tag_C = [&](char a, char bch, uint8_t v) { aux[la++] = a; aux[la++] = bch; aux[la++] = 'C'; aux[la++] = v; };
    if (o.nh[r]) tag_C('N', 'H', (o.nh[r] == 1) ? 1 : 3);
    tag_C('H', 'I', 1);
    aux[la++] = 'A'; aux[la++] = 'S'; aux[la++] = 'S';
    const uint16_t as = (uint16_t)(2 * qlen - 2 - rng.below(6));
    memcpy(aux + la, &as, 2); la += 2;
    tag_C('n', 'M', (uint8_t)rng.below(3));
    if (o.xs[r]) { aux[la++] = 'X'; aux[la++] = 'S'; aux[la++] = 'A'; aux[la++] = o.xs[r]; }
    /* jM:B:c / jI:B:i as STAR writes them with --outSAMattributes All */
    aux[la++] = 'j'; aux[la++] = 'M'; aux[la++] = 'B'; aux[la++] = 'c';
    int32_t one = 1;
    memcpy(aux + la, &one, 4); la += 4;
    aux[la++] = spliced ? 1 : (uint8_t)-1;
    aux[la++] = 'j'; aux[la++] = 'I'; aux[la++] = 'B'; aux[la++] = 'i';
    int32_t two = spliced ? 2 : 1;
    memcpy(aux + la, &two, 4); la += 4;
    int32_t ji0 = spliced ? o.pos[r] + 50 : -1, ji1 = o.pos[r] + 150;
    memcpy(aux + la, &ji0, 4); la += 4;
    if (spliced) { memcpy(aux + la, &ji1, 4); la += 4; }
    int32_t x[8];
    x[0] = o.tid[r];
    x[1] = o.pos[r];
    x[2] = (int32_t)(((uint32_t)reg2bin(o.pos[r], o.pos[r] + rlen) << 16) | (255u << 8) | (uint32_t)lq);
    x[3] = (int32_t)(((uint32_t)o.flag[r] << 16) | nc);
    x[4] = (int32_t)qlen;
    x[5] = (o.flag[r] & 1) ? o.tid[r] : -1;
    x[6] = (o.flag[r] & 1) ? o.pos[r] + (int32_t)rng.below(300) - 150 : -1;
    if (x[6] < -1) x[6] = 0;
    x[7] = (o.flag[r] & 1) ? (int32_t)rng.below(500) - 250 : 0;
    const size_t seq_bytes = (size_t)((qlen + 1) / 2);
    int32_t block = 32 + lq + (int32_t)nc * 4 + (int32_t)seq_bytes + (int32_t)qlen + la;
    b.append((const char *)&block, 4);
    b.append((const char *)x, 32);
    b.append(qname, lq);
    b.append((const char *)&o.cigar[c0], (size_t)nc * 4);
    const size_t at = b.size();
    b.resize(at + seq_bytes + (size_t)qlen);
    uint8_t *sq = (uint8_t *)&b[at];
    static const uint8_t nib[4] = {1, 2, 4, 8};
    for (size_t k = 0; k < seq_bytes; k += 8) {
        uint64_t v = rng.next();
        for (size_t q = k; q < k + 8 && q < seq_bytes; q++, v >>= 4) sq[q] = (uint8_t)((nib[v & 3] << 4) | nib[(v >> 2) & 3]);
    }
    uint8_t *ql = sq + seq_bytes;
    for (int64_t k = 0; k < qlen; k += 8) {
        uint64_t v = rng.next();
        for (int64_t q = k; q < k + 8 && q < qlen; q++, v >>= 8) ql[q] = (uint8_t)(((v & 0xFF) < 200) ? 37 : (2 + (v & 0x1F)));
    }
    b.append((const char *)aux, la);
}

#include <zlib.h>
/* `in` as BGZF blocks appended to `out` */
static bool bgzf_compress(const std::string &in, std::string &out, int level) {
    const size_t BLOCK = 0xff00;
    std::vector<uint8_t> tmp(0x10000 + 1024);
    for (size_t at = 0; at < in.size() || (at == 0 && in.empty());) {
        const size_t n = std::min(BLOCK, in.size() - at);
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
        zs.next_in = (Bytef *)(in.data() + at);
        zs.avail_in = (uInt)n;
        zs.next_out = tmp.data();
        zs.avail_out = (uInt)tmp.size();
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); return false; }
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        const uint8_t head[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
        out.append((const char *)head, 12);
        const uint8_t sub[4] = {'B', 'C', 2, 0};
        out.append((const char *)sub, 4);
        const uint16_t bsize = (uint16_t)(clen + 25);
        out.append((const char *)&bsize, 2);
        out.append((const char *)tmp.data(), clen);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)(in.data() + at), (uInt)n);
        const uint32_t isize = (uint32_t)n;
        out.append((const char *)&crc, 4);
        out.append((const char *)&isize, 4);
        at += n;
        if (n == 0) break;
    }
    return true;
}

extern "C" int ipx_synth_write_bam_full(const ipx_synth_data *d, const char *path, uint64_t begin, uint64_t end, int threads, int level) {
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    if (end > d->out.tid.size()) end = d->out.tid.size();
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if (threads > 64) threads = 64;
    bool ok = true;
    {
        std::string z;
        ok = bgzf_compress(bam_header_bytes(d), z, level) && fwrite(z.data(), 1, z.size(), fp) == z.size();
    }
    const uint64_t SLICE = 1u << 16; /* records a thread turns into BGZF blocks at a time */
    for (uint64_t r0 = begin; r0 < end && ok; r0 += SLICE * (uint64_t)threads) {
        std::vector<std::string> zs((size_t)threads);
        std::vector<char> good((size_t)threads, 1);
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) {
            pool.emplace_back([&, t]() {
                const uint64_t lo = r0 + SLICE * (uint64_t)t, hi = std::min(end, lo + SLICE);
                if (lo >= hi) return;
                std::string raw;
                raw.reserve((size_t)(hi - lo) * 400);
                for (uint64_t r = lo; r < hi; r++) append_full_record(d, r, raw);
                good[(size_t)t] = bgzf_compress(raw, zs[(size_t)t], level) ? 1 : 0;
            });
        }
        for (auto &th : pool) th.join();
        for (int t = 0; t < threads && ok; t++)
            ok = good[(size_t)t] && (zs[(size_t)t].empty() || fwrite(zs[(size_t)t].data(), 1, zs[(size_t)t].size(), fp) == zs[(size_t)t].size());
    }
    static const uint8_t eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (ok) ok = fwrite(eof_block, 1, 28, fp) == 28;
    return (fclose(fp) == 0 && ok) ? 0 : -1;
}

/* deterministic background base for (contig, pos) */
static inline char bg_base(uint64_t seed, int contig, int64_t p) {
    uint64_t z = seed ^ ((uint64_t)contig << 48) ^ (uint64_t)p;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return "ACGT"[z & 3];
}

static void plant(std::string &s, int64_t p, char a, char b) {
    if (p >= 0 && p + 1 < (int64_t)s.size()) { s[p] = a; s[p + 1] = b; }
}

extern "C" int ipx_synth_write_genome(const ipx_synth_data *d, const char *fasta_path) {
    FILE *fa = fopen(fasta_path, "wb");
    if (!fa) return -1;
    std::string fai_path = std::string(fasta_path) + ".fai";
    FILE *fai = fopen(fai_path.c_str(), "w");
    if (!fai) { fclose(fa); return -1; }
    int64_t off = 0;
    const int W = 60;
    /* the contigs' text is made by a pool of threads (a few at a time: each holds its sequence twice), written in order */
    const size_t n_contigs = d->contig_names.size();
    size_t wave = std::thread::hardware_concurrency();
    if (wave < 1) wave = 1;
    if (wave > 8) wave = 8;
    auto make_text = [&](size_t c, std::string &buf) {
        const int64_t len = (int64_t)d->contig_len[c];
        std::string s((size_t)len, 'A');
        for (int64_t p = 0; p < len; p++) s[p] = bg_base(d->spec.model_seed, (int)c, p);
        for (const Gene &g : d->genes) {
            if (g.contig != (int)c) continue;
            int64_t gp = g.start;
            for (size_t e = 0; e + 1 < g.exon_len.size(); e++) {
                const int64_t i0 = gp + g.exon_len[e], i1 = i0 + g.intron_len[e];
                const uint8_t m = g.motif[e];
                if (m < 3) {
                    static const char *don[3] = {"GT", "GC", "AT"}, *acc[3] = {"AG", "AG", "AC"};
                    if (g.strand == '+') {
                        plant(s, i0, don[m][0], don[m][1]);
                        plant(s, i1 - 2, acc[m][0], acc[m][1]);
                    } else {
                        /* reverse complement of donor/acceptor */
                        auto rc = [](char ch) { return ch == 'A' ? 'T' : ch == 'T' ? 'A' : ch == 'G' ? 'C' : 'G'; };
                        plant(s, i0, rc(acc[m][1]), rc(acc[m][0]));
                        plant(s, i1 - 2, rc(don[m][1]), rc(don[m][0]));
                    }
                } else {
                    plant(s, i0, 'T', 'T');
                    plant(s, i1 - 2, 'C', 'C');
                }
                gp = i1;
            }
        }
        /* soft-masked stretches and N runs */
        Rng rng(d->spec.model_seed * 31 + c);
        const int64_t n_feat = len / 20000 + 1;
        for (int64_t k = 0; k < n_feat; k++) {
            int64_t p = (int64_t)(rng.uniform() * (double)len), l = 50 + (int64_t)rng.below(400);
            const bool ns = rng.below(10) == 0;
            for (int64_t q = p; q < std::min(len, p + l); q++) s[q] = ns ? 'N' : (char)tolower(s[q]);
        }
        buf.clear();
        buf.reserve((size_t)(len + len / W + 2));
        for (int64_t p = 0; p < len; p += W) {
            buf.append(s, (size_t)p, (size_t)std::min<int64_t>(W, len - p));
            buf.push_back('\n');
        }
    };
    bool ok = true;
    for (size_t c0 = 0; c0 < n_contigs && ok; c0 += wave) {
        const size_t c1 = std::min(n_contigs, c0 + wave);
        std::vector<std::string> text(c1 - c0);
        std::vector<std::thread> pool;
        for (size_t c = c0; c < c1; c++) pool.emplace_back([&, c]() { make_text(c, text[c - c0]); });
        for (auto &th : pool) th.join();
        for (size_t c = c0; c < c1 && ok; c++) {
            const int nl = fprintf(fa, ">%s\n", d->contig_names[c].c_str());
            off += nl;
            fprintf(fai, "%s\t%lld\t%lld\t%d\t%d\n", d->contig_names[c].c_str(), (long long)d->contig_len[c], (long long)off, W, W + 1);
            ok = fwrite(text[c - c0].data(), 1, text[c - c0].size(), fa) == text[c - c0].size();
            off += (int64_t)text[c - c0].size();
        }
    }
    if (!ok) { fclose(fai); fclose(fa); return -1; }
    fclose(fai);
    return fclose(fa) == 0 ? 0 : -1;
}
