/* calls_tsv.cc -- see calls_tsv.h.  Reading goes through zlib's gzread, which also passes plain files through;
 * the reference picks gzip by the ".gz" suffix (AutoGzip.hh:9-11), which gives the same bytes for both kinds. */
#include "calls_tsv.h"
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <algorithm>
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

/* what the header says about the columns read_junction asks for (-1: absent, which only matters once a row is read) */
struct Header {
    int ncol = 0;
    int c_chrom = -1, c_start = -1, c_end = -1, c_strand = -1, c_uniq = -1, c_multi = -1, c_unsure = -1, c_left = -1, c_right = -1,
        c_conf = -1, c_sites = -1;
    bool complete() const {
        return c_chrom >= 0 && c_start >= 0 && c_end >= 0 && c_strand >= 0 && c_uniq >= 0 && c_multi >= 0 && c_unsure >= 0 && c_left >= 0 &&
               c_right >= 0 && c_conf >= 0 && c_sites >= 0;
    }
};

static Header read_header(const std::string &line) {
    std::vector<std::string> words;
    split_tabs(line, words);
    Header h;
    h.ncol = (int)words.size();
    std::unordered_map<std::string, int> col;
    for (int i = 0; i < h.ncol; i++) col[words[i]] = i; /* later duplicates win, as with map::operator[] */
    auto need = [&](const char *name) -> int {
        auto it = col.find(name);
        return it == col.end() ? -1 : it->second;
    };
    h.c_chrom = need("chrom"); h.c_start = need("intron_start"); h.c_end = need("intron_end"); h.c_strand = need("strand");
    h.c_uniq = need("uniq_mapped_count"); h.c_multi = need("multi_mapped_count"); h.c_unsure = need("unsure_mapped_count");
    h.c_left = need("max_left_overhang"); h.c_right = need("max_right_overhang"); h.c_conf = need("confidence");
    h.c_sites = need("splice_sites");
    return h;
}

static int32_t chrom_id(CallsRows &out, const std::string &chrom) {
    auto ci = out.chrom_index.find(chrom);
    if (ci == out.chrom_index.end()) {
        ci = out.chrom_index.emplace(chrom, (int32_t)out.chrom_names.size()).first;
        out.chrom_names.push_back(chrom);
    }
    return ci->second;
}

static uint32_t site_id(CallsRows &out, const std::string &sites) {
    auto si = out.site_index.find(sites);
    if (si == out.site_index.end()) {
        si = out.site_index.emplace(sites, (uint32_t)out.site_strings.size()).first;
        out.site_strings.push_back(sites);
    }
    return si->second;
}

/* one row: Tsv::next_row's column check, junction_filter_check, read_junction -- in the reference's order, with its errors */
static void add_line(const std::string &line, const Header &h, const std::string &path, int junction_filter, std::vector<std::string> &words,
                     CallsRows &out) {
    auto get = [&](int c, const char *name) -> const std::string & {
        if (c < 0) throw std::runtime_error(std::string("column '") + name + "' not found: " + path);
        return words[c];
    };
    split_tabs(line, words);
    if ((int)words.size() != h.ncol) {
        const std::string n = std::to_string(h.ncol);
        throw std::runtime_error("expected " + n + " columns, got " + n + ":" + path); /* sic, tsv.hh:71 */
    }
    const std::string &sites = get(h.c_sites, "splice_sites");
    /* junction_filter_check (splice_juncs.cc:52-61) */
    if (!sites.empty() && junction_filter == IPX_FILTER_CANON && !canonical(sites)) return;
    const uint32_t start = to_uint(get(h.c_start, "intron_start"));
    const uint32_t end = to_uint(get(h.c_end, "intron_end"));
    const std::string &chrom = get(h.c_chrom, "chrom");
    const uint32_t a0 = start - to_uint(get(h.c_left, "max_left_overhang"));
    const uint32_t a1 = end + to_uint(get(h.c_right, "max_right_overhang"));
    const std::string &strand = get(h.c_strand, "strand");
    const uint32_t uniq = to_uint(get(h.c_uniq, "uniq_mapped_count"));
    const uint32_t multi = to_uint(get(h.c_multi, "multi_mapped_count"));
    const uint32_t unsure = to_uint(get(h.c_unsure, "unsure_mapped_count"));
    const float conf = to_float(get(h.c_conf, "confidence"));
    out.chrom.push_back(chrom_id(out, chrom));
    out.i0.push_back(start);
    out.i1.push_back(end);
    out.a0.push_back(a0);
    out.a1.push_back(a1);
    out.strand.push_back(strand.empty() ? '\0' : strand[0]);
    out.sites.push_back(site_id(out, sites));
    out.counts.push_back(uniq);
    out.counts.push_back(multi);
    out.counts.push_back(unsure);
    out.conf.push_back(conf);
}

static std::string open_error(const std::string &path) {
    return std::string("Failed to open ") + (path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0 ? "gzip file: " : "file: ") + path;
}

void load_calls_tsv(const std::string &path, int junction_filter, CallsRows &out) {
    gzFile gz = gzopen(path.c_str(), "rb");
    if (!gz) throw std::runtime_error(open_error(path));
    gzbuffer(gz, 1 << 20);
    LineReader rd(gz);
    std::string line;
    std::vector<std::string> words;
    try {
        if (!rd.getline(line)) throw std::runtime_error("empty TSV file: " + path);
        const Header h = read_header(line);
        while (rd.getline(line)) add_line(line, h, path, junction_filter, words, out);
    } catch (...) {
        gzclose(gz);
        throw;
    }
    gzclose(gz);
}

/* The rows of one file through the device tokeniser.  false: not an input for the device (nothing was added to `out`). */
static bool load_calls_tsv_tokenized(const std::string &path, int junction_filter, CallsRows &out, const TsvTokenizer &tok) {
    gzFile gz = gzopen(path.c_str(), "rb");
    if (!gz) throw std::runtime_error(open_error(path));
    gzbuffer(gz, 1 << 20);
    /* the whole file, inflated, in memory the tokeniser likes */
    size_t cap = 32u << 20, n = 0;
    char *bytes = (char *)tok.alloc(cap);
    struct Guard {
        const TsvTokenizer &tok;
        gzFile gz;
        char *&bytes;
        ipx_tsv_rows *rows = nullptr;
        ~Guard() {
            if (gz) gzclose(gz);
            if (rows) tok.rows_free(rows);
            if (bytes) tok.release(bytes);
        }
    } guard{tok, gz, bytes};
    if (!bytes) return false;
    for (;;) {
        if (n == cap) {
            if (cap >= (1ull << 32)) return false; /* beyond what the device entry point takes */
            char *grown = (char *)tok.alloc(cap * 2);
            if (!grown) return false;
            memcpy(grown, bytes, n);
            tok.release(bytes);
            bytes = grown;
            cap *= 2;
        }
        const int got = gzread(gz, bytes + n, (unsigned)std::min<size_t>(cap - n, 1u << 30));
        if (got <= 0) break;
        n += (size_t)got;
    }
    if (n == 0) throw std::runtime_error("empty TSV file: " + path);
    const char *nl = (const char *)memchr(bytes, '\n', n);
    const size_t header_len = nl ? (size_t)(nl - bytes) : n, body = nl ? header_len + 1 : n;
    const Header h = read_header(std::string(bytes, header_len));
    if (!h.complete()) return false; /* the reference's "column not found" comes from the host parser, at the first row */
    ipx_tsv_columns cols;
    cols.n_columns = h.ncol;
    cols.chrom = h.c_chrom; cols.intron_start = h.c_start; cols.intron_end = h.c_end; cols.strand = h.c_strand;
    cols.uniq_mapped_count = h.c_uniq; cols.multi_mapped_count = h.c_multi; cols.unsure_mapped_count = h.c_unsure;
    cols.max_left_overhang = h.c_left; cols.max_right_overhang = h.c_right; cols.confidence = h.c_conf; cols.splice_sites = h.c_sites;
    const char *text = bytes + body;
    const int rc = tok.tokenize(text, n - body, &cols, &guard.rows);
    if (rc == IPX_ERR_ARG) return false;
    if (rc != IPX_OK || !guard.rows) throw std::runtime_error("the device could not tokenise " + path);
    const ipx_tsv_rows &r = *guard.rows;
    const size_t n0 = out.chrom.size();
    out.chrom.reserve(n0 + r.n_rows); out.i0.reserve(n0 + r.n_rows); out.i1.reserve(n0 + r.n_rows); out.a0.reserve(n0 + r.n_rows);
    out.a1.reserve(n0 + r.n_rows); out.strand.reserve(n0 + r.n_rows); out.sites.reserve(n0 + r.n_rows);
    out.counts.reserve(3 * (n0 + r.n_rows)); out.conf.reserve(n0 + r.n_rows);
    std::vector<std::string> words;
    std::string line;
    /* consecutive rows mostly share their chromosome and their splice sites: remember the last of each */
    const char *last_chrom = nullptr;
    uint32_t last_chrom_len = 0;
    int32_t last_chrom_id = -1;
    uint64_t last_s8 = 0;
    uint32_t last_s_len = 0xFFFFFFFFu, last_site_id = 0;
    bool last_site_kept = true;
    for (uint64_t i = 0; i < r.n_rows; i++) {
        if (r.flags[i]) {
            line.assign(text + r.line_off[i], r.line_len[i]);
            add_line(line, h, path, junction_filter, words, out); /* (throws what the reference throws for this row) */
            continue;
        }
        const uint64_t s8 = r.sites8[i];
        const uint32_t s_len = r.sites_len[i];
        if (s8 != last_s8 || s_len != last_s_len) {
            char sb[9];
            for (uint32_t k = 0; k < 8; k++) sb[k] = (char)((s8 >> (8 * k)) & 0xFF);
            const std::string sites(sb, s_len);
            last_site_kept = sites.empty() || junction_filter != IPX_FILTER_CANON || canonical(sites);
            /* (a string the filter drops gets no number, as in the host parser) */
            last_site_id = last_site_kept ? site_id(out, sites) : 0u;
            last_s8 = s8;
            last_s_len = s_len;
        }
        if (!last_site_kept) continue;
        const char *cp = text + r.chrom_off[i];
        const uint32_t cl = r.chrom_len[i];
        if (last_chrom == nullptr || cl != last_chrom_len || memcmp(cp, last_chrom, cl) != 0) {
            last_chrom_id = chrom_id(out, std::string(cp, cl));
            last_chrom = cp;
            last_chrom_len = cl;
        }
        out.chrom.push_back(last_chrom_id);
        out.i0.push_back(r.intron_start[i]);
        out.i1.push_back(r.intron_end[i]);
        out.a0.push_back(r.anchor_start[i]);
        out.a1.push_back(r.anchor_end[i]);
        out.strand.push_back(r.strand[i]);
        out.sites.push_back(last_site_id);
        out.counts.push_back(r.counts[3 * i]);
        out.counts.push_back(r.counts[3 * i + 1]);
        out.counts.push_back(r.counts[3 * i + 2]);
        out.conf.push_back(r.confidence[i]);
    }
    return true;
}

static void load_calls_tsvs_impl(const std::vector<std::string> &paths, int junction_filter, CallsRows &out, const TsvTokenizer *tok);

void load_calls_tsvs(const std::vector<std::string> &paths, int junction_filter, CallsRows &out) {
    load_calls_tsvs_impl(paths, junction_filter, out, nullptr);
}
void load_calls_tsvs(const std::vector<std::string> &paths, int junction_filter, CallsRows &out, const TsvTokenizer &tok) {
    load_calls_tsvs_impl(paths, junction_filter, out, &tok);
}

static void load_calls_tsvs_impl(const std::vector<std::string> &paths, int junction_filter, CallsRows &out, const TsvTokenizer *tok) {
    auto load_one = [&](const std::string &path, CallsRows &rows) {
        if (tok && (!tok->ready || tok->ready()) && load_calls_tsv_tokenized(path, junction_filter, rows, *tok)) return;
        load_calls_tsv(path, junction_filter, rows);
    };
    unsigned threads = std::thread::hardware_concurrency();
    if (threads > 16u) threads = 16u;
    if (const char *e = getenv("IPX_PARSE_THREADS")) threads = (unsigned)atoi(e);
    if (threads > paths.size()) threads = (unsigned)paths.size();
    if (threads <= 1u) {
        for (const std::string &f : paths) load_one(f, out);
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
                load_one(paths[i], parts[i]->rows);
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
    /* The files' rows into the common columns: the dictionaries are merged file by file (names new to `out` get their
     * numbers in the order the files first used them, as if everything had been parsed into `out`), which also gives
     * every file its place in the columns; the rows are then copied, and renumbered, by the pool.  Everything up to and
     * including the first bad file is used. */
    size_t n_use = paths.size();
    for (size_t i = 0; i < paths.size(); i++)
        if (parts[i]->failed) {
            n_use = i + 1;
            break;
        }
    std::vector<std::vector<int32_t>> chrom_map(n_use);
    std::vector<std::vector<uint32_t>> site_map(n_use);
    std::vector<size_t> at(n_use + 1, out.chrom.size());
    for (size_t i = 0; i < n_use; i++) {
        const CallsRows &part = parts[i]->rows;
        chrom_map[i].resize(part.chrom_names.size());
        for (size_t k = 0; k < part.chrom_names.size(); k++) chrom_map[i][k] = chrom_id(out, part.chrom_names[k]);
        site_map[i].resize(part.site_strings.size());
        for (size_t k = 0; k < part.site_strings.size(); k++) site_map[i][k] = site_id(out, part.site_strings[k]);
        at[i + 1] = at[i] + part.chrom.size();
    }
    const size_t total = at[n_use];
    out.chrom.resize(total); out.sites.resize(total); out.i0.resize(total); out.i1.resize(total); out.a0.resize(total);
    out.a1.resize(total); out.strand.resize(total); out.conf.resize(total); out.counts.resize(3 * total);
    std::atomic<size_t> next_copy(0);
    auto copy = [&]() {
        for (size_t i; (i = next_copy.fetch_add(1)) < n_use;) {
            const CallsRows &part = parts[i]->rows;
            const size_t n0 = at[i], n = part.chrom.size();
            for (size_t r = 0; r < n; r++) {
                out.chrom[n0 + r] = chrom_map[i][part.chrom[r]];
                out.sites[n0 + r] = site_map[i][part.sites[r]];
            }
            std::copy(part.i0.begin(), part.i0.end(), out.i0.begin() + n0);
            std::copy(part.i1.begin(), part.i1.end(), out.i1.begin() + n0);
            std::copy(part.a0.begin(), part.a0.end(), out.a0.begin() + n0);
            std::copy(part.a1.begin(), part.a1.end(), out.a1.begin() + n0);
            std::copy(part.strand.begin(), part.strand.end(), out.strand.begin() + n0);
            std::copy(part.conf.begin(), part.conf.end(), out.conf.begin() + n0);
            std::copy(part.counts.begin(), part.counts.end(), out.counts.begin() + 3 * n0);
            parts[i].reset(); /* give the file's rows back */
        }
    };
    std::string error = n_use && parts[n_use - 1]->failed ? parts[n_use - 1]->error : std::string();
    const bool failed = n_use && parts[n_use - 1]->failed;
    pool.clear();
    for (unsigned t = 0; t < threads && t < n_use; t++) pool.emplace_back(copy);
    for (auto &t : pool) t.join();
    if (failed) throw std::runtime_error(error);
}

}  // namespace ipx
