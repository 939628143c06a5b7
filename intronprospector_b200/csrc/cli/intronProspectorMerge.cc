/* intronProspectorMerge.cc -- drop-in command line tool: same options, outputs and error text as the
 * reference's intronProspectorMerge (src/intronProspectorMerge.cc), with the per-intron fold running on a
 * B200 behind ipx_merge_rows_fold (include/ipx.h).  Host side only: options, TSV loading, text output.
 */
#include <getopt.h>
#include <memory>
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

/* intron_prospector_merge (intronProspectorMerge.cc:207-238) */
static void merge(const Options &opts) {
    ipx::CallsRows rows;
    string parse_error;
    try {
        ipx::load_calls_tsvs(opts.input_calls_tsvs, opts.junction_filter, rows);
    } catch (const std::exception &e) {
        parse_error = e.what(); /* rows read before the bad one are still folded: their stderr lines come first */
    }
    rows.bind();
    ipx_params p;
    memset(&p, 0, sizeof(p));
    if (const char *d = getenv("IPX_DEVICE")) p.device = atoi(d);
    ipx_handle *h = nullptr;
    if (ipx_create(&p, &h) != IPX_OK) throw runtime_error("no usable CUDA device: the merge fold needs an NVIDIA B200 (sm_100a) GPU");
    struct Guard {
        ipx_handle *h;
        ~Guard() { ipx_destroy(h); }
    } guard{h};
    if (rows.chrom_names.empty()) rows.chrom_names.push_back("");
    rows.bind();
    check_ipx(h, ipx_set_targets(h, (int32_t)rows.chrom_ptrs.size(), rows.chrom_ptrs.data()));
    ipx_table *t = nullptr;
    ipx_text log;
    int rc = ipx_merge_rows_fold(h, &rows.rows, &t, &log);
    if (log.data) cerr << string(log.data, log.len);
    ipx_text_free(&log);
    if (rc != IPX_OK) throw runtime_error(ipx_last_error(h));
    if (!parse_error.empty()) {
        ipx_table_free(t);
        throw runtime_error(parse_error);
    }
    auto emit = [&](const string &fn, int fmt, bool header) {
        if (fn == "") return;
        OutFile out(fn);
        ipx_text txt;
        if (header) {
            ipx_format_tsv_header(&txt);
            out.write(txt);
            ipx_text_free(&txt);
        }
        if (ipx_format(t, rows.chrom_ptrs.data(), fmt, 0.0f, 0, t->n_junctions, &txt) != IPX_OK) throw runtime_error("formatting failed");
        out.write(txt);
        ipx_text_free(&txt);
    };
    emit(opts.junction_bed, IPX_FMT_JUNCTION_BED, false);
    emit(opts.intron_bed, IPX_FMT_INTRON_BED9, false);
    emit(opts.intron_bed6, IPX_FMT_INTRON_BED6, false);
    emit(opts.intron_call_tsv, IPX_FMT_INTRON_TSV, true);
    ipx_table_free(t);
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
