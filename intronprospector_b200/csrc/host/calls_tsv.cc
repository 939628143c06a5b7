/* calls_tsv.cc -- see calls_tsv.h.  Reading goes through zlib's gzread, which also passes plain files through;
 * the reference picks gzip by the ".gz" suffix (AutoGzip.hh:9-11), which gives the same bytes for both kinds. */
#include "calls_tsv.h"
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <atomic>
#include <memory>
#include <stdexcept>
#include <thread>

namespace ipx {

CallsRows::CallsRows() {
    memset(&rows, 0, sizeof(rows));
    site_strings.push_back("");
    site_index.emplace("", 0u);
}

void CallsRows::bind() {
    chrom_ptrs.resize(chrom_names.size());
    for (size_t i = 0; i < chrom_names.size(); i++) chrom_ptrs[i] = chrom_names[i].c_str();
    site_ptrs.resize(site_strings.size());
    for (size_t i = 0; i < site_strings.size(); i++) site_ptrs[i] = site_strings[i].c_str();
    rows.n_rows = chrom.size();
    rows.chrom = chrom.data();
    rows.intron_start = i0.data();
    rows.intron_end = i1.data();
    rows.anchor_start = a0.data();
    rows.anchor_end = a1.data();
    rows.strand = strand.data();
    rows.sites = sites.data();
    rows.counts = counts.data();
    rows.confidence = conf.data();
    rows.site_strings = site_ptrs.data();
    rows.n_site_strings = (uint32_t)site_ptrs.size();
}

namespace {

struct LineReader {
    gzFile gz;
    std::vector<char> buf;
    size_t beg = 0, end = 0;
    bool eof = false;
    explicit LineReader(gzFile g) : gz(g), buf(1 << 20) {}
    /* std::getline semantics: false only when nothing at all could be read */
    bool getline(std::string &line) {
        line.clear();
        bool any = false;
        for (;;) {
            if (beg == end) {
                if (eof) return any;
                int got = gzread(gz, buf.data(), (unsigned)buf.size());
                if (got <= 0) {
                    eof = true;
                    return any;
                }
                beg = 0;
                end = (size_t)got;
            }
            any = true;
            const char *p = buf.data() + beg;
            const char *nl = (const char *)memchr(p, '\n', end - beg);
            if (nl) {
                line.append(p, nl - p);
                beg += (size_t)(nl - p) + 1;
                return true;
            }
            line.append(p, end - beg);
            beg = end;
        }
    }
};

void split_tabs(const std::string &line, std::vector<std::string> &words) {
    /* Tsv::split_line (tsv.hh:27-38) */
    words.clear();
    size_t prev = 0, cur = line.find('\t');
    while (cur != std::string::npos) {
        words.push_back(line.substr(prev, cur - prev));
        prev = cur + 1;
        cur = line.find('\t', prev);
    }
    words.push_back(line.substr(prev));
}

uint32_t to_uint(const std::string &s) {
    /* string_to_uint (type_ops.hh:19-26) */
    char *end;
    uint32_t n = (uint32_t)strtoul(s.c_str(), &end, 10);
    if (end == s.c_str() || *end != '\0') throw std::runtime_error("invalid integer \"" + s + "\"");
    return n;
}

float to_float(const std::string &s) {
    /* string_to_float (type_ops.hh:37-44) */
    char *end;
    float n = strtof(s.c_str(), &end);
    if (end == s.c_str() || *end != '\0') throw std::runtime_error("invalid float \"" + s + "\"");
    return n;
}

bool canonical(const std::string &s) {
    /* is_canonical (splice_juncs.cc:25-38): any letter case */
    if (s.size() != 5) return false;
    char u[6];
    for (int i = 0; i < 5; i++) u[i] = (s[i] >= 'a' && s[i] <= 'z') ? (char)(s[i] - 32) : s[i];
    u[5] = 0;
    return !strcmp(u, "GT/AG") || !strcmp(u, "GC/AG") || !strcmp(u, "AT/AC");
}

}  // namespace

void load_calls_tsv(const std::string &path, int junction_filter, CallsRows &out) {
    gzFile gz = gzopen(path.c_str(), "rb");
    if (!gz) throw std::runtime_error(std::string("Failed to open ") + (path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0 ? "gzip file: " : "file: ") + path);
    gzbuffer(gz, 1 << 20);
    LineReader rd(gz);
    std::string line;
    std::vector<std::string> words;
    try {
        if (!rd.getline(line)) throw std::runtime_error("empty TSV file: " + path);
        split_tabs(line, words);
        const int ncol = (int)words.size();
        std::unordered_map<std::string, int> col;
        for (int i = 0; i < ncol; i++) col[words[i]] = i; /* later duplicates win, as with map::operator[] */
        auto need = [&](const char *name) -> int {
            auto it = col.find(name);
            return it == col.end() ? -1 : it->second;
        };
        const int c_chrom = need("chrom"), c_start = need("intron_start"), c_end = need("intron_end"), c_strand = need("strand"),
                  c_uniq = need("uniq_mapped_count"), c_multi = need("multi_mapped_count"), c_unsure = need("unsure_mapped_count"),
                  c_left = need("max_left_overhang"), c_right = need("max_right_overhang"), c_conf = need("confidence"),
                  c_sites = need("splice_sites");
        auto get = [&](int c, const char *name) -> const std::string & {
            if (c < 0) throw std::runtime_error(std::string("column '") + name + "' not found: " + path);
            return words[c];
        };
        while (rd.getline(line)) {
            split_tabs(line, words);
            if ((int)words.size() != ncol) {
                const std::string n = std::to_string(ncol);
                throw std::runtime_error("expected " + n + " columns, got " + n + ":" + path); /* sic, tsv.hh:71 */
            }
            const std::string &sites = get(c_sites, "splice_sites");
            /* junction_filter_check (splice_juncs.cc:52-61) */
            if (!sites.empty() && junction_filter == IPX_FILTER_CANON && !canonical(sites)) continue;
            const uint32_t start = to_uint(get(c_start, "intron_start"));
            const uint32_t end = to_uint(get(c_end, "intron_end"));
            const std::string &chrom = get(c_chrom, "chrom");
            const uint32_t a0 = start - to_uint(get(c_left, "max_left_overhang"));
            const uint32_t a1 = end + to_uint(get(c_right, "max_right_overhang"));
            const std::string &strand = get(c_strand, "strand");
            const uint32_t uniq = to_uint(get(c_uniq, "uniq_mapped_count"));
            const uint32_t multi = to_uint(get(c_multi, "multi_mapped_count"));
            const uint32_t unsure = to_uint(get(c_unsure, "unsure_mapped_count"));
            const float conf = to_float(get(c_conf, "confidence"));
            auto ci = out.chrom_index.find(chrom);
            if (ci == out.chrom_index.end()) {
                ci = out.chrom_index.emplace(chrom, (int32_t)out.chrom_names.size()).first;
                out.chrom_names.push_back(chrom);
            }
            auto si = out.site_index.find(sites);
            if (si == out.site_index.end()) {
                si = out.site_index.emplace(sites, (uint32_t)out.site_strings.size()).first;
                out.site_strings.push_back(sites);
            }
            out.chrom.push_back(ci->second);
            out.i0.push_back(start);
            out.i1.push_back(end);
            out.a0.push_back(a0);
            out.a1.push_back(a1);
            out.strand.push_back(strand.empty() ? '\0' : strand[0]);
            out.sites.push_back(si->second);
            out.counts.push_back(uniq);
            out.counts.push_back(multi);
            out.counts.push_back(unsure);
            out.conf.push_back(conf);
        }
    } catch (...) {
        gzclose(gz);
        throw;
    }
    gzclose(gz);
}

namespace {

/* append the rows of one file, parsed into dictionaries of its own, translating them into the common dictionaries:
 * names new to `out` get their numbers in the order the file first used them, as if it had been parsed into `out` */
void append_part(const CallsRows &part, CallsRows &out) {
    std::vector<int32_t> chrom_map(part.chrom_names.size());
    for (size_t i = 0; i < part.chrom_names.size(); i++) {
        auto it = out.chrom_index.find(part.chrom_names[i]);
        if (it == out.chrom_index.end()) {
            it = out.chrom_index.emplace(part.chrom_names[i], (int32_t)out.chrom_names.size()).first;
            out.chrom_names.push_back(part.chrom_names[i]);
        }
        chrom_map[i] = it->second;
    }
    std::vector<uint32_t> site_map(part.site_strings.size());
    for (size_t i = 0; i < part.site_strings.size(); i++) {
        auto it = out.site_index.find(part.site_strings[i]);
        if (it == out.site_index.end()) {
            it = out.site_index.emplace(part.site_strings[i], (uint32_t)out.site_strings.size()).first;
            out.site_strings.push_back(part.site_strings[i]);
        }
        site_map[i] = it->second;
    }
    const size_t n0 = out.chrom.size(), n = part.chrom.size();
    out.chrom.resize(n0 + n);
    out.sites.resize(n0 + n);
    for (size_t i = 0; i < n; i++) {
        out.chrom[n0 + i] = chrom_map[part.chrom[i]];
        out.sites[n0 + i] = site_map[part.sites[i]];
    }
    out.i0.insert(out.i0.end(), part.i0.begin(), part.i0.end());
    out.i1.insert(out.i1.end(), part.i1.begin(), part.i1.end());
    out.a0.insert(out.a0.end(), part.a0.begin(), part.a0.end());
    out.a1.insert(out.a1.end(), part.a1.begin(), part.a1.end());
    out.strand.insert(out.strand.end(), part.strand.begin(), part.strand.end());
    out.counts.insert(out.counts.end(), part.counts.begin(), part.counts.end());
    out.conf.insert(out.conf.end(), part.conf.begin(), part.conf.end());
}

}  // namespace

void load_calls_tsvs(const std::vector<std::string> &paths, int junction_filter, CallsRows &out) {
    unsigned threads = std::thread::hardware_concurrency();
    if (threads > 16u) threads = 16u;
    if (const char *e = getenv("IPX_PARSE_THREADS")) threads = (unsigned)atoi(e);
    if (threads > paths.size()) threads = (unsigned)paths.size();
    if (threads <= 1u) {
        for (const std::string &f : paths) load_calls_tsv(f, junction_filter, out);
        return;
    }
    struct Part {
        CallsRows rows;
        std::string error;
        bool failed = false;
    };
    std::vector<std::unique_ptr<Part>> parts(paths.size());
    for (auto &p : parts) p.reset(new Part());
    std::atomic<size_t> next(0), first_bad(paths.size());
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= paths.size() || i > first_bad.load()) break; /* nothing after the first bad file is used */
            try {
                load_calls_tsv(paths[i], junction_filter, parts[i]->rows);
            } catch (const std::exception &e) {
                parts[i]->error = e.what();
                parts[i]->failed = true;
                size_t seen = first_bad.load();
                while (i < seen && !first_bad.compare_exchange_weak(seen, i)) {
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
    for (size_t i = 0; i < paths.size(); i++) {
        append_part(parts[i]->rows, out);
        if (parts[i]->failed) throw std::runtime_error(parts[i]->error);
        parts[i].reset(); /* give the file's rows back before the next one is appended */
    }
}

}  // namespace ipx
