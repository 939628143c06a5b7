/* intronProspectorMerge.cc -- drop-in command line tool: same options, outputs and error text as the
 * reference's intronProspectorMerge (src/intronProspectorMerge.cc), with the per-intron fold running on a
 * B200 behind ipx_merge_rows_fold (include/ipx.h).  Host side only: options, TSV loading, text output.
 */
#include <getopt.h>
#include <stdio.h>
#include <time.h>
#include <algorithm>
#include <atomic>
#include <memory>
#include <thread>
#include <vector>

#include "calls_tsv.h"
#include "cli_common.h"
#include "ipx.h"

using namespace std;

static const char *usage_msg =
    "intronProspectorMerge [options] [intron_calls_tsv ...]\n"
    "\n"
    "Merge intron call TSVs produced by intronProspector --intron-calls.  The fold runs on an NVIDIA B200 GPU.\n"
    "\n"
    "Options:\n"
    "  -h, --help                      print this message\n"
    "  -v, --version                   print version\n"
    "  -i, --intron-calls-files=FILE   file with one input TSV path per line\n"
    "  -j, --junction-bed=FILE         junction BED12 output\n"
    "  -n, --intron-bed=FILE           intron BED9 output\n"
    "  -b, --intron-bed6=FILE          intron BED6 output\n"
    "  -c, --intron-calls=FILE         merged intron calls TSV output\n"
    "  -f, --sj-filter=all|canon       splice junction filter\n"
    "Input and output files ending in .gz are gzip compressed.\n";

static void usage() {
    cerr << usage_msg;
    exit(0);
}

/* IPX_TRACE=1: wall-clock marks of the tool's phases on stderr */
struct PhaseTrace {
    bool on = getenv("IPX_TRACE") != NULL;
    double t0 = now();
    static double now() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    void mark(const char *what) {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[merge] %-32s %9.1f ms\n", what, t - t0);
        t0 = t;
    }
};

struct Options {
    vector<string> input_calls_tsvs;
    string intron_calls_files, junction_bed, intron_bed, intron_bed6, intron_call_tsv;
    int junction_filter = IPX_FILTER_NULL;
};

static void parse_options(int argc, char *argv[], Options &o) {
    /* intronProspectorMerge.cc:90-131 */
    struct option long_options[] = {{"help", no_argument, NULL, 'h'},
                                    {"version", no_argument, NULL, 'v'},
                                    {"intron-calls-files", required_argument, NULL, 'i'},
                                    {"junction-bed", required_argument, NULL, 'j'},
                                    {"intron-bed", required_argument, NULL, 'n'},
                                    {"intron-bed6", required_argument, NULL, 'b'},
                                    {"intron-calls", required_argument, NULL, 'c'},
                                    {"sj-filter", required_argument, NULL, 'f'},
                                    {NULL, 0, NULL, 0}};
    const char *short_options = "hvi:j:n:b:c:f:";
    int c;
    while ((c = getopt_long(argc, argv, short_options, long_options, NULL)) != -1) {
        switch (c) {
            case 'i': o.intron_calls_files = optarg; break;
            case 'j': o.junction_bed = optarg; break;
            case 'n': o.intron_bed = optarg; break;
            case 'b': o.intron_bed6 = optarg; break;
            case 'c': o.intron_call_tsv = optarg; break;
            case 'f': {
                string spec(optarg);
                if (spec == "all") o.junction_filter = IPX_FILTER_ALL;
                else if (spec == "canon") o.junction_filter = IPX_FILTER_CANON;
                else throw invalid_argument("invalid splice junction filter '" + spec + "', expected 'all' or 'canon'");
                break;
            }
            case 'h': usage();
            case 'v':
                cerr << IPX_PACKAGE_NAME << " " << IPX_PACKAGE_VERSION << " " << IPX_PACKAGE_URL << endl;
                exit(0);
            case '?':
            default: cmd_error("bug: option not handled");
        }
    }
    while (optind < argc) o.input_calls_tsvs.push_back(string(argv[optind++]));
    if (o.intron_calls_files.size() > 0) {
        /* read_input_call_list (intronProspectorMerge.cc:133-139) */
        gzFile gz = gzopen(o.intron_calls_files.c_str(), "rb");
        if (!gz) throw runtime_error("Failed to open file: " + o.intron_calls_files);
        string all;
        char buf[65536];
        int got;
        while ((got = gzread(gz, buf, sizeof(buf))) > 0) all.append(buf, got);
        gzclose(gz);
        size_t pos = 0;
        while (pos < all.size()) {
            size_t nl = all.find('\n', pos);
            if (nl == string::npos) nl = all.size();
            o.input_calls_tsvs.push_back(all.substr(pos, nl - pos));
            pos = nl + 1;
        }
    }
    if (o.input_calls_tsvs.size() == 0)
        cmd_error("a least one input call TSV required on the command list or with the --intron_call_files (-i) option");
}

/* One key range of the merge on one device: the rows whose chromosomes fall into the range, in input order. */
struct Part {
    int device = 0;
    vector<int32_t> chrom;
    vector<uint32_t> i0, i1, a0, a1, sites, counts;
    vector<char> strand;
    vector<float> conf;
    vector<uint64_t> row_ids;
    ipx_merge_rows rows;
    /* results */
    ipx_table *table = nullptr;
    int rc = IPX_OK;
    string error, log;
    vector<uint64_t> log_rows;
    uint64_t fail_row = ~0ull;
};

/* Chromosome ranges with about the same number of rows each, contiguous in the order the outputs are sorted by
 * (chromosome name): part_of[chromosome index]. */
static vector<int> balanced_ranges(const ipx::CallsRows &rows, int n_parts) {
    const size_t nc = rows.chrom_names.size();
    vector<uint64_t> per_chrom(nc, 0);
    for (int32_t c : rows.chrom) per_chrom[c]++;
    vector<uint32_t> order(nc);
    for (size_t i = 0; i < nc; i++) order[i] = (uint32_t)i;
    sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return rows.chrom_names[a] < rows.chrom_names[b]; });
    vector<int> part_of(nc, 0);
    const uint64_t total = rows.chrom.size();
    uint64_t seen = 0;
    for (uint32_t c : order) {
        /* the part in which the chromosome's middle row falls */
        const uint64_t mid = seen + per_chrom[c] / 2;
        part_of[c] = total ? (int)min<uint64_t>((uint64_t)n_parts - 1, mid * (uint64_t)n_parts / total) : 0;
        seen += per_chrom[c];
    }
    return part_of;
}

static void fold_part(Part *p, const vector<const char *> *names) {
    PhaseTrace tr;
    ipx_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.device = p->device;
    ipx_handle *h = nullptr;
    if (ipx_create(&prm, &h) != IPX_OK) {
        p->rc = IPX_ERR_CUDA;
        p->error = "no usable CUDA device: the merge fold needs an NVIDIA B200 (sm_100a) GPU";
        return;
    }
    tr.mark("  handle");
    p->rc = ipx_set_targets(h, (int32_t)names->size(), names->data());
    tr.mark("  targets");
    if (p->rc == IPX_OK) {
        ipx_text log;
        uint64_t *log_rows = nullptr, n_log_rows = 0;
        p->rc = ipx_merge_rows_fold_part(h, &p->rows, p->row_ids.data(), &p->table, &log, &log_rows, &n_log_rows, &p->fail_row);
        if (log.data) p->log.assign(log.data, log.len);
        ipx_text_free(&log);
        if (log_rows) p->log_rows.assign(log_rows, log_rows + n_log_rows);
        ipx_free(log_rows);
    }
    if (p->rc != IPX_OK) p->error = ipx_last_error(h);
    tr.mark("  fold call");
    ipx_destroy(h);
    tr.mark("  handle released");
}

/* The rows are tokenised on the device (ipx_tsv_tokenize); IPX_TSV_HOST=1 keeps the host parser (tests compare the two).
 * Bringing the device up takes a second or so of a tool's run: a thread does it from the start, and the files that come
 * by before it is done are parsed by the host's threads (the rows are the same either way). */
static int tokenizer_device = 0;
static std::atomic<bool> device_up{false};
static std::atomic<int> files_on_device{0};
static int tokenize_on_device(const char *bytes, uint64_t n, const ipx_tsv_columns *cols, ipx_tsv_rows **out) {
    files_on_device++;
    return ipx_tsv_tokenize(tokenizer_device, bytes, n, cols, out);
}
static bool device_is_up() { return device_up.load(); }
static void bring_device_up() {
    PhaseTrace tr;
    ipx_tsv_columns c;
    memset(&c, 0, sizeof(c));
    c.n_columns = 1;
    ipx_tsv_rows *r = nullptr;
    if (ipx_tsv_tokenize(tokenizer_device, "", 0, &c, &r) == IPX_OK) device_up.store(true);
    ipx_tsv_rows_free(r);
    tr.mark("(device up)");
}

/* intron_prospector_merge (intronProspectorMerge.cc:207-238) */
static void merge(const Options &opts) {
    PhaseTrace tr;
    ipx::CallsRows rows;
    string parse_error;
    if (const char *d = getenv("IPX_DEVICE")) tokenizer_device = atoi(d);
    thread warm_up(bring_device_up);
    struct Join {
        thread &t;
        ~Join() {
            if (t.joinable()) t.join();
        }
    } join_warm_up{warm_up};
    try {
        if (getenv("IPX_TSV_HOST")) {
            ipx::load_calls_tsvs(opts.input_calls_tsvs, opts.junction_filter, rows);
        } else {
            /* IPX_TSV_DEVICE=1: wait for the device first, so that every file goes through it (tests) */
            if (getenv("IPX_TSV_DEVICE")) warm_up.join();
            const ipx::TsvTokenizer tok{malloc, free, tokenize_on_device, ipx_tsv_rows_free, device_is_up};
            ipx::load_calls_tsvs(opts.input_calls_tsvs, opts.junction_filter, rows, tok);
        }
    } catch (const std::exception &e) {
        parse_error = e.what(); /* rows read before the bad one are still folded: their stderr lines come first */
    }
    if (rows.chrom_names.empty()) rows.chrom_names.push_back("");
    rows.bind();
    if (tr.on) fprintf(stderr, "[merge] files tokenised on the device: %d of %zu\n", files_on_device.load(), opts.input_calls_tsvs.size());
    tr.mark("rows read");
    /* Key ranges over the devices of this host (BASELINE config 5): every device folds the rows of a contiguous range of
     * chromosomes.  A key lives in one range, so the partial tables are simply concatenated in key order.
     * IPX_MERGE_DEVICES=n sets the number of devices (default: one per 32 M rows, at most all that are visible);
     * IPX_MERGE_PARTS=n asks for n ranges whatever the number of devices (range k runs on device k mod devices: tests);
     * IPX_DEVICE picks the device of a single-device run. */
    const int n_dev = max(1, ipx_device_count()); /* (without a device ipx_create fails below: there is no other path) */
    /* one device per 32 M rows: bringing a second device up costs a second or two (2.1 s each for two B200s side by side),
     * the fold of 15 M rows 50 ms */
    int n_parts = (int)min<uint64_t>((uint64_t)n_dev, (rows.chrom.size() + (32u << 20) - 1) / (32u << 20));
    if (n_parts < 1) n_parts = 1;
    if (const char *e = getenv("IPX_MERGE_DEVICES")) n_parts = max(1, min(atoi(e), n_dev));
    if (const char *e = getenv("IPX_MERGE_PARTS")) n_parts = max(1, min(atoi(e), 64));
    n_parts = (int)min<size_t>((size_t)n_parts, rows.chrom_names.size());
    vector<unique_ptr<Part>> parts;
    if (n_parts <= 1) {
        parts.emplace_back(new Part());
        Part &p = *parts[0];
        if (const char *d = getenv("IPX_DEVICE")) p.device = atoi(d);
        p.rows = rows.rows;
        p.row_ids.resize(rows.chrom.size());
        for (size_t i = 0; i < p.row_ids.size(); i++) p.row_ids[i] = i;
        fold_part(&p, &rows.chrom_ptrs);
    } else {
        const vector<int> part_of = balanced_ranges(rows, n_parts);
        for (int k = 0; k < n_parts; k++) {
            parts.emplace_back(new Part());
            parts[k]->device = k % n_dev;
        }
        /* the rows of every part, gathered by a thread per part, which then folds them on its device */
        vector<thread> threads;
        for (int k = 0; k < n_parts; k++)
            threads.emplace_back([&, k]() {
                Part &p = *parts[k];
                const size_t n = rows.chrom.size();
                size_t mine = 0;
                for (size_t i = 0; i < n; i++) mine += part_of[rows.chrom[i]] == k;
                p.chrom.reserve(mine); p.i0.reserve(mine); p.i1.reserve(mine); p.a0.reserve(mine); p.a1.reserve(mine);
                p.sites.reserve(mine); p.counts.reserve(3 * mine); p.strand.reserve(mine); p.conf.reserve(mine); p.row_ids.reserve(mine);
                for (size_t i = 0; i < n; i++) {
                    if (part_of[rows.chrom[i]] != k) continue;
                    p.chrom.push_back(rows.chrom[i]); p.i0.push_back(rows.i0[i]); p.i1.push_back(rows.i1[i]);
                    p.a0.push_back(rows.a0[i]); p.a1.push_back(rows.a1[i]); p.sites.push_back(rows.sites[i]);
                    p.counts.push_back(rows.counts[3 * i]); p.counts.push_back(rows.counts[3 * i + 1]); p.counts.push_back(rows.counts[3 * i + 2]);
                    p.strand.push_back(rows.strand[i]); p.conf.push_back(rows.conf[i]); p.row_ids.push_back(i);
                }
                p.rows = rows.rows; /* the dictionaries */
                p.rows.n_rows = p.chrom.size();
                p.rows.chrom = p.chrom.data(); p.rows.intron_start = p.i0.data(); p.rows.intron_end = p.i1.data();
                p.rows.anchor_start = p.a0.data(); p.rows.anchor_end = p.a1.data(); p.rows.strand = p.strand.data();
                p.rows.sites = p.sites.data(); p.rows.counts = p.counts.data(); p.rows.confidence = p.conf.data();
                fold_part(&p, &rows.chrom_ptrs);
            });
        for (auto &t : threads) t.join();
    }
    tr.mark("fold");
    struct Tables {
        vector<unique_ptr<Part>> &parts;
        ~Tables() {
            for (auto &p : parts) ipx_table_free(p->table);
        }
    } guard{parts};
    /* what the reference does in input order: strand-conflict lines on stderr as their rows come by, the fatal
     * disagreement at the first row that has one */
    uint64_t fail_row = ~0ull;
    string fail_text;
    for (auto &p : parts) {
        if (p->rc != IPX_OK && p->rc != IPX_ERR_MERGE) throw runtime_error(p->error);
        if (p->rc == IPX_ERR_MERGE && p->fail_row < fail_row) {
            fail_row = p->fail_row;
            fail_text = p->error;
        }
    }
    vector<pair<uint64_t, string>> lines;
    for (auto &p : parts) {
        size_t pos = 0, k = 0;
        while (pos < p->log.size()) {
            size_t nl = p->log.find('\n', pos);
            if (nl == string::npos) nl = p->log.size() - 1;
            const uint64_t row = k < p->log_rows.size() ? p->log_rows[k] : 0;
            if (row < fail_row) lines.emplace_back(row, p->log.substr(pos, nl + 1 - pos));
            pos = nl + 1;
            k++;
        }
    }
    stable_sort(lines.begin(), lines.end(), [](const pair<uint64_t, string> &a, const pair<uint64_t, string> &b) { return a.first < b.first; });
    for (auto &l : lines) cerr << l.second;
    if (fail_row != ~0ull) throw runtime_error(fail_text);
    if (!parse_error.empty()) throw runtime_error(parse_error);
    /* the ranges are contiguous in the order of the outputs: sjN = position in anchor order over all of them */
    vector<Part *> in_order;
    for (auto &p : parts) in_order.push_back(p.get());
    if (n_parts > 1) {
        /* by the name of a part's first chromosome (parts without rows are empty tables) */
        vector<pair<string, Part *>> keyed;
        for (Part *p : in_order)
            if (p->table && p->table->n_junctions) keyed.emplace_back(rows.chrom_names[p->table->chrom[0]], p);
        sort(keyed.begin(), keyed.end(), [](const pair<string, Part *> &a, const pair<string, Part *> &b) { return a.first < b.first; });
        in_order.clear();
        for (auto &k : keyed) in_order.push_back(k.second);
        uint32_t base = 0;
        for (Part *p : in_order) {
            vector<uint32_t> ij(p->table->ijunc, p->table->ijunc + p->table->n_junctions);
            for (auto &x : ij) x += base;
            ipx_table_set_ijunc(p->table, ij.data());
            base += (uint32_t)p->table->n_junctions;
        }
    }
    /* The output files are independent of each other: they are opened in the reference's order (so that a file that
     * cannot be opened stops the run where it stops the reference's: the files before it written, none after it), then
     * formatted and written by a thread each. */
    struct Job {
        const string *fn;
        int fmt;
        bool header;
        unique_ptr<OutFile> out;
        string error;
    } jobs[4] = {{&opts.junction_bed, IPX_FMT_JUNCTION_BED, false, nullptr, ""}, {&opts.intron_bed, IPX_FMT_INTRON_BED9, false, nullptr, ""},
                 {&opts.intron_bed6, IPX_FMT_INTRON_BED6, false, nullptr, ""}, {&opts.intron_call_tsv, IPX_FMT_INTRON_TSV, true, nullptr, ""}};
    string open_error;
    int n_jobs = 0;
    for (Job &j : jobs) {
        if (*j.fn != "") {
            try {
                j.out.reset(new OutFile(*j.fn));
            } catch (const std::exception &e) {
                open_error = e.what();
                break;
            }
        }
        n_jobs++;
    }
    auto emit = [&](Job &j) {
        if (!j.out) return;
        ipx_text txt;
        if (j.header) {
            ipx_format_tsv_header(&txt);
            j.out->write(txt);
            ipx_text_free(&txt);
        }
        for (Part *p : in_order) {
            if (!p->table) continue;
            if (ipx_format(p->table, rows.chrom_ptrs.data(), j.fmt, 0.0f, 0, p->table->n_junctions, &txt) != IPX_OK)
                throw runtime_error("formatting failed");
            j.out->write(txt);
            ipx_text_free(&txt);
        }
        j.out.reset();
    };
    vector<thread> writers;
    for (int k = 0; k < n_jobs; k++)
        writers.emplace_back([&emit, &jobs, k]() {
            try {
                emit(jobs[k]);
            } catch (const std::exception &e) {
                jobs[k].error = e.what();
                if (jobs[k].error.empty()) jobs[k].error = "output failed";
            }
        });
    for (auto &w : writers) w.join();
    for (int k = 0; k < n_jobs; k++)
        if (!jobs[k].error.empty()) throw runtime_error(jobs[k].error);
    if (!open_error.empty()) throw runtime_error(open_error);
    tr.mark("outputs written");
}

int main(int argc, char *argv[]) {
    Options opts;
    try {
        parse_options(argc, argv, opts);
    } catch (const std::exception &ex) {
        cerr << "Error parsing command line: " << ex.what() << endl;
        exit(1);
    }
    try {
        merge(opts);
    } catch (const std::exception &e) {
        cmd_error(e.what());
    }
    return 0;
}
