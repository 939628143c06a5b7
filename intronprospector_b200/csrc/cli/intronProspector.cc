/* intronProspector.cc -- drop-in command line tool: same options, defaults, run modes, outputs and error
 * text as the reference's intronProspector (src/intronProspector.cc), with junction extraction and
 * aggregation running on a B200 behind the C ABI of include/ipx.h.
 *
 * This file is the host side of the boundary: option parsing (intronProspector.cc:131-300), the three run
 * modes and their record-selection rules (intronProspector.cc:356-401, junctions_extractor.cc:620-683), and
 * writing the formatted text.  It contains no junction arithmetic.
 */
#include <getopt.h>
#include <time.h>
#include <algorithm>
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>

#include "cli_common.h"
#include "genome_loader.h"
#include "htslib/sam.h"
#include "ipx.h"
#include "bam_pipeline.h"
#include "record_source.h"
#include "../ipx_walk.cuh" /* the automaton, host build: per-read annotations below */

using namespace std;

static const char *usage_msg =
    "intronProspector [options] [readaligns ...]\n"
    "\n"
    "Find putative intron junctions in RNA-Seq alignments (SAM/BAM).  With no input file, or '/dev/stdin',\n"
    "reads standard input.  Junction extraction runs on an NVIDIA B200 GPU.\n"
    "\n"
    "Options:\n"
    "  -h, --help                        print this message\n"
    "  -v, --version                     print version\n"
    "  -u, --unsorted                    input is not coordinate sorted (uses more memory)\n"
    "  -a, --min-anchor-length=N         minimum anchor length on each side of an intron [8]\n"
    "  -i, --min-intron-length=N         minimum intron length [70]\n"
    "  -I, --max-intron-length=N         maximum intron length [500000]\n"
    "      --allow-anchor-indels         allow indels in anchors\n"
    "      --max-anchor-indel-size=N     maximum indel size in anchors, with --allow-anchor-indels [36]\n"
    "  -C, --min-confidence-score=X      minimum Shannon-Wiener confidence score to report [1.0]\n"
    "  -s, --strandness=UN|RF|FR         library strandness [UN]\n"
    "  -X, --excludes=multi              exclude a category of reads\n"
    "  -g, --genome-fasta=FASTA          genome (with .fai) to get splice sites; defaults --sj-filter to canon\n"
    "  -f, --sj-filter=all|canon         splice junction filter\n"
    "      --skip-missing-targets        warn rather than fail for targets missing from the genome\n"
    "  -j, --junction-bed=FILE           junction BED12 output\n"
    "  -n, --intron-bed=FILE             intron BED9 output\n"
    "  -b, --intron-bed6=FILE            intron BED6 output\n"
    "  -c, --intron-calls=FILE           intron calls TSV output\n"
    "  -p, --pass-through=FILE           copy alignments to this file\n"
    "  -D, --debug-trace=FILE            per-read trace TSV\n"
    "      --set-XS-strand-tag           set XS:A on pass-through records from the observed introns\n"
    "      --set-TS-strand-tag           set TS:A on pass-through records from the observed introns\n"
    "Output files ending in .gz are gzip compressed.\n";

static void usage() {
    cerr << usage_msg;
    exit(0);
}

/* CmdParser (intronProspector.cc:104-300) */
struct Options {
    vector<string> bam_files;
    uint32_t min_anchor_length = 8;
    uint32_t min_intron_length = 70;
    uint32_t max_intron_length = 500000;
    bool allow_anchor_indels = false;
    bool unsorted = false; /* uninitialised in the reference (intronProspector.cc:131-143); false is what its tests assume */
    uint32_t max_anchor_indel_size = 36;
    float min_confidence_score = 1.0;
    int junction_filter = IPX_FILTER_NULL;
    int strandness = IPX_STRAND_UN;
    unsigned excludes = 0;
    string genome_fa;
    bool skip_missing_targets = false;
    string junction_bed, intron_bed, intron_bed6, intron_call_tsv, bam_pass_through, debug_trace_tsv;
    bool set_XS_strand_tag = false, set_TS_strand_tag = false;
};

static int parse_strandness(const string &s) {
    string su(s);
    transform(su.begin(), su.end(), su.begin(), ::toupper);
    if (su == "UN") return IPX_STRAND_UN;
    if (su == "RF") return IPX_STRAND_RF;
    if (su == "FR") return IPX_STRAND_FR;
    cerr << "Error: invalid strandness value '" + s + "', expected one of 'UN', 'RF', or 'FR' (case insensitive)" << endl;
    exit(1);
}

static unsigned parse_exclude_cat(const string &cat) {
    if (cat == "multi") return IPX_EXCLUDE_MULTI;
    cerr << "Error: invalid exclude category '" << cat << "', expected one of: 'multi'" << endl;
    exit(1);
}

static int parse_filter(const string &spec) {
    if (spec == "all") return IPX_FILTER_ALL;
    if (spec == "canon") return IPX_FILTER_CANON;
    throw invalid_argument("invalid splice junction filter '" + spec + "', expected 'all' or 'canon'");
}

static void parse_options(int argc, char *argv[], Options &o) {
    static const int OPT_SET_XS_STRAND_TAG = 256;
    static const int OPT_SET_TS_STRAND_TAG = 257;
    /* same tables as the reference (intronProspector.cc:159-185), including its quirks: -S takes an argument,
     * --allow-anchor-indels / --max-anchor-indel-size have no short form, long options may be abbreviated */
    struct option long_options[] = {{"help", no_argument, NULL, 'h'},
                                    {"version", no_argument, NULL, 'v'},
                                    {"unsorted", no_argument, NULL, 'u'},
                                    {"min-anchor-length", required_argument, NULL, 'a'},
                                    {"min-intron-length", required_argument, NULL, 'i'},
                                    {"max-intron-length", required_argument, NULL, 'I'},
                                    {"allow-anchor-indels", no_argument, NULL, 'd'},
                                    {"max-anchor-indel-size", required_argument, NULL, 'm'},
                                    {"min-confidence-score", required_argument, NULL, 'C'},
                                    {"strandness", required_argument, NULL, 's'},
                                    {"excludes", required_argument, NULL, 'X'},
                                    {"genome-fasta", required_argument, NULL, 'g'},
                                    {"sj-filter", required_argument, NULL, 'f'},
                                    {"skip-missing-targets", no_argument, NULL, 'S'},
                                    {"junction-bed", required_argument, NULL, 'j'},
                                    {"intron-bed", required_argument, NULL, 'n'},
                                    {"intron-bed6", required_argument, NULL, 'b'},
                                    {"intron-calls", required_argument, NULL, 'c'},
                                    {"pass-through", required_argument, NULL, 'p'},
                                    {"debug-trace", required_argument, NULL, 'D'},
                                    {"set-XS-strand-tag", no_argument, NULL, OPT_SET_XS_STRAND_TAG},
                                    {"set-TS-strand-tag", no_argument, NULL, OPT_SET_TS_STRAND_TAG},
                                    {NULL, 0, NULL, 0}};
    const char *short_options = "hvua:i:I:C:s:X:g:S:j:n:b:c:p:D:f:";
    int c;
    while ((c = getopt_long(argc, argv, short_options, long_options, NULL)) != -1) {
        switch (c) {
            case 'a': o.min_anchor_length = atoi(optarg); break;
            case 'u': o.unsorted = true; break;
            case 'd': o.allow_anchor_indels = true; break;
            case 'm': o.max_anchor_indel_size = cli_to_unsigned(optarg); break;
            case 'i': o.min_intron_length = cli_to_unsigned(optarg); break;
            case 'I': o.max_intron_length = cli_to_unsigned(optarg); break;
            case 'C': o.min_confidence_score = cli_to_float(optarg); break;
            case 's': o.strandness = parse_strandness(optarg); break;
            case 'X': o.excludes |= parse_exclude_cat(optarg); break;
            case 'g': o.genome_fa = string(optarg); break;
            case 'f': o.junction_filter = parse_filter(string(optarg)); break;
            case 'S': o.skip_missing_targets = true; break;
            case 'j': o.junction_bed = optarg; break;
            case 'n': o.intron_bed = optarg; break;
            case 'b': o.intron_bed6 = optarg; break;
            case 'c': o.intron_call_tsv = optarg; break;
            case 'p': o.bam_pass_through = optarg; break;
            case 'D': o.debug_trace_tsv = optarg; break;
            case OPT_SET_XS_STRAND_TAG: o.set_XS_strand_tag = true; break;
            case OPT_SET_TS_STRAND_TAG: o.set_TS_strand_tag = true; break;
            case 'h': usage();
            case 'v':
                cerr << IPX_PACKAGE_NAME << " " << IPX_PACKAGE_VERSION << " " << IPX_PACKAGE_URL << endl;
                exit(0);
            case '?':
            default: cmd_error("bug: option not handled");
        }
    }
    while (optind < argc) o.bam_files.push_back(string(argv[optind++]));
    if (o.bam_files.size() == 0) o.bam_files.push_back("/dev/stdin");
    if ((o.bam_files.size() > 1) and (o.bam_pass_through != "")) cmd_error("can only have one input BAM with --bam-pass-through");
    if ((o.junction_filter == IPX_FILTER_CANON) and (o.genome_fa == "")) cmd_error("--sj-filter=canon requires --genome_fa");
    if ((o.junction_filter == IPX_FILTER_NULL) and (o.genome_fa != "")) o.junction_filter = IPX_FILTER_CANON;
    if (not o.allow_anchor_indels) o.max_anchor_indel_size = 0;
    if ((o.set_XS_strand_tag or o.set_TS_strand_tag) and ((o.bam_pass_through.size() == 0) or (o.genome_fa.size() == 0)))
        cmd_error("--set-XS-strand-tag and --set-TS-strand-tag require --pass-through and --genome-fasta");
}

/* OutputFiles (intronProspector.cc:303-328): the TSV header is written when the file is opened */
struct Outputs {
    unique_ptr<OutFile> junction_bed, intron_bed, intron_bed6, intron_call;
    explicit Outputs(const Options &o) {
        if (o.junction_bed != "") junction_bed.reset(new OutFile(o.junction_bed));
        if (o.intron_bed != "") intron_bed.reset(new OutFile(o.intron_bed));
        if (o.intron_bed6 != "") intron_bed6.reset(new OutFile(o.intron_bed6));
        if (o.intron_call_tsv != "") {
            intron_call.reset(new OutFile(o.intron_call_tsv));
            ipx_text t;
            ipx_format_tsv_header(&t);
            intron_call->write(t);
            ipx_text_free(&t);
        }
    }
};

/* one input file being read (JunctionsExtractor::open / close, junctions_extractor.cc:571-617) */
struct Input {
    samFile *in = nullptr;
    sam_hdr_t *hdr = nullptr;
    samFile *out = nullptr;
    bam1_t *rec = nullptr;
    string name;
    explicit Input(const string &bam) : name(bam) {
        in = sam_open(bam.c_str(), "r");
        if (in == NULL) throw runtime_error("Error opening BAM/SAM/CRAM file: " + bam);
        hdr = sam_hdr_read(in);
        if (hdr == NULL) throw runtime_error("Error opening BAM/SAM/CRAM file: " + bam);
        rec = bam_init1();
    }
    void open_pass_through(const string &path) {
        const htsFormat *fmt = hts_get_format(in);
        if (!((fmt->format == sam) || (fmt->format == bam)))
            throw invalid_argument("Error: pass-through is only implemented for SAM or BAM files: " + path);
        out = hts_open_format(path.c_str(), "w", fmt);
        if (out == NULL) throw runtime_error("Error opening BAM/SAM/CRAM file: " + path);
        if (sam_hdr_write(out, hdr) < 0) throw runtime_error("Error writing SAM header: " + path);
    }
    /* BAM records can be packed where they lie in the reader's buffer, without building a bam1_t, when nobody needs
     * the record itself afterwards (no pass-through, no per-read side output).  IPX_NO_RAW_PACK=1 turns that off. */
    bool raw = false;
    void enable_raw() { raw = hts_get_format(in)->format == bam && getenv("IPX_NO_RAW_PACK") == NULL; }
    bool peek_raw(const uint8_t **rec_bytes, uint32_t *len) { return raw && ipx_bam_peek(in, rec_bytes, len) == 1; }
    void consume_raw() { ipx_bam_consume(in); }
    /* read_align (junctions_extractor.cc:620-634) without the one-record read-ahead, which the caller keeps */
    bool read() {
        int stat = sam_read1(in, hdr, rec);
        if (stat < -1) throw runtime_error("Error reader BAM record: " + name);
        return stat >= 0;
    }
    void pass(const bam1_t *b) {
        if (out != NULL && sam_write1(out, hdr, b) < 0) throw runtime_error("Error writing BAM record to pass-through file");
    }
    ~Input() {
        if (out) sam_close(out);
        if (rec) bam_destroy1(rec);
        if (hdr) sam_hdr_destroy(hdr);
        if (in) sam_close(in);
    }
};

/* Per-read side outputs: --debug-trace rows and --set-XS/TS-strand-tag (create_junction / update_junction trace lines
 * junctions_extractor.cc:446-485, orientCnt :526-528, update_strand_tag :98-111).  Both depend on the junction a read's
 * event lands in -- its creator-fixed strand and motif, and whether the event came before the junction existed -- which
 * is known once the device table is flushed.  So with either option the records are kept in memory in stream order,
 * and after the flush every kept read is walked once more on the host with the same automaton the kernels run
 * (ipx_walk.cuh) and looked up in the table.  The junction tables themselves still come from the device. */
class Annotator {
    struct Kept {
        bam1_t *rec;
        uint64_t seq;
        int32_t tid;      /* engine target index */
        bool processed;   /* went through process_alignment */
        bool pass;        /* is written to the pass-through file */
    };
    vector<Kept> kept_;
    const Options &opts_;
    unique_ptr<OutFile> trace_;

    struct Key {
        int32_t tid;
        uint32_t i0, i1;
        bool operator==(const Key &o) const { return tid == o.tid && i0 == o.i0 && i1 == o.i1; }
    };
    struct KeyHash {
        size_t operator()(const Key &k) const {
            uint64_t x = ((uint64_t)(uint32_t)k.tid << 40) ^ ((uint64_t)k.i0 << 20) ^ k.i1;
            x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
            return (size_t)x;
        }
    };

    public:
    explicit Annotator(const Options &o) : opts_(o) {
        if (o.debug_trace_tsv != "") trace_.reset(new OutFile(o.debug_trace_tsv));
    }
    ~Annotator() {
        for (auto &k : kept_) bam_destroy1(k.rec);
    }
    bool active() const { return trace_ || opts_.set_XS_strand_tag || opts_.set_TS_strand_tag; }
    /* JunctionsExtractor::open writes the header every time an input is opened (junctions_extractor.cc:579-584) */
    void input_opened() {
        if (!trace_) return;
        const string h = "chrom\tintron_start\tintron_end\tqname\tflag\tstrand\tsplice_sites\tanchor_start\tanchor_end\n";
        trace_->write(h.data(), h.size());
    }
    void keep(const bam1_t *b, uint64_t seq, int32_t tid, bool processed, bool pass) {
        Kept k{bam_copy1(bam_init1(), b), seq, tid, processed, pass};
        kept_.push_back(k);
    }
    void mark_pass_last() { if (!kept_.empty()) kept_.back().pass = true; }

    /* after the flush: trace rows, strand tags, and the delayed pass-through writes, in stream order */
    void finish(const ipx_table *t, const vector<string> &targets, Input *in) {
        std::unordered_map<Key, uint64_t, KeyHash> rows;
        rows.reserve(t->n_junctions * 2 + 16);
        for (uint64_t i = 0; i < t->n_junctions; i++) rows.emplace(Key{t->chrom[i], t->intron_start[i], t->intron_end[i]}, i);
        ipx::ExtractParams xp;
        xp.min_anchor_length = opts_.min_anchor_length;
        xp.max_anchor_indel_size = opts_.max_anchor_indel_size;
        xp.min_intron_length = opts_.min_intron_length;
        xp.max_intron_length = opts_.max_intron_length;
        xp.excludes = opts_.excludes;
        xp.strandness = opts_.strandness;
        string line;
        for (auto &k : kept_) {
            bam1_t *b = k.rec;
            int orient = 0;
            if (k.processed && ipx::record_is_live(b->core.n_cigar, b->core.tid, b->core.flag)) {
                uint8_t nh, xs;
                ipx::scan_tags(b, &nh, &xs);
                const char strand = ipx::read_strand(xp.strandness, b->core.flag, xs);
                const bool may_emit = ipx::read_may_emit(xp, b->core.flag, nh);
                auto sink = [&](const ipx::Candidate &c) {
                    if (!may_emit || !ipx::junction_qc(xp, c)) return;
                    auto it = rows.find(Key{k.tid, (uint32_t)c.intron_start, (uint32_t)c.intron_end});
                    if (it == rows.end()) return; /* the junction never came to exist */
                    const uint64_t r = it->second;
                    const uint64_t ev = (k.seq << 20) | (c.ordinal & 0xFFFFFu);
                    if (!t->creation_seq || ev < t->creation_seq[r]) return; /* before it existed: add_junction returned NULL */
                    const char *sites = t->splice_sites + 8 * r;
                    if (trace_) {
                        line.clear();
                        line += targets[k.tid]; line += '\t';
                        line += to_string((long long)c.intron_start); line += '\t';
                        line += to_string((long long)c.intron_end); line += '\t';
                        line += bam_get_qname(b); line += '\t';
                        line += to_string((unsigned)b->core.flag); line += '\t';
                        if (ev == t->creation_seq[r]) { /* create_junction: the corrected strand, then an empty column */
                            line += t->strand[r]; line += '\t';
                        } else {                        /* update_junction: the read's strand, then the junction's */
                            line += strand; line += '\t';
                            line += t->strand[r];
                        }
                        line += '\t';
                        line += sites; line += '\t';
                        line += to_string((long long)c.anchor_start); line += '\t';
                        line += to_string((long long)c.anchor_end); line += '\n';
                        trace_->write(line.data(), line.size());
                    }
                    if (sites[0] >= 'A' && sites[0] <= 'Z') orient += (t->strand[r] == '+') ? 1 : -1; /* Junction::is_canonical */
                };
                ipx::WalkState w;
                w.init(b->core.pos);
                const uint32_t *cig = bam_get_cigar(b);
                uint32_t bad = 0;
                for (uint32_t i = 0; i < b->core.n_cigar && !bad; i++) bad = ipx::walk_op(w, cig[i], xp.max_anchor_indel_size, sink);
                if (!bad) ipx::walk_finish(w, sink);
            }
            if (orient != 0) {
                if (opts_.set_XS_strand_tag) update_strand_tag("XS", orient, b);
                if (opts_.set_TS_strand_tag) update_strand_tag("TS", orient, b);
            }
            if (k.pass && in) in->pass(b);
            bam_destroy1(b);
        }
        kept_.clear();
    }

    private:
    static void update_strand_tag(const char *tag, int orient, bam1_t *b) {
        const uint8_t strand = (orient > 0) ? '+' : '-';
        uint8_t *existing = bam_aux_get(b, tag);
        if (existing != NULL && bam_aux_del(b, existing) != 0) throw runtime_error("Error deleting read tag");
        if (bam_aux_append(b, tag, 'A', 1, &strand) != 0) throw runtime_error("Error adding read tag");
    }
};

/* IPX_TRACE=1: wall-clock phases of the run on stderr */
struct PhaseTrace {
    bool on = getenv("IPX_TRACE") != NULL;
    double t0 = now();
    static double now() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + ts.tv_nsec * 1e-9;
    }
    void mark(const char *what) {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[ipx cli] %-32s %8.3f s\n", what, t - t0);
        t0 = t;
    }
};
static PhaseTrace g_trace;

/* One batch of packed records in pinned host memory (ipx_batch_compact: 10 bytes per record + 4 per CIGAR word).
 * Two of them take turns: while the copy of one to the device is in flight, the other is being filled. */
struct PinnedBatch {
    size_t cap_rec = 0, cap_cig = 0, n = 0, n_cig = 0;
    int32_t *pos = nullptr;
    uint16_t *flag = nullptr, *n_ops = nullptr;
    uint8_t *nh = nullptr, *xs = nullptr;
    uint32_t *cigar = nullptr;
    vector<int32_t> run_tid;
    vector<uint64_t> run_start;
    vector<uint32_t> ops32; /* the operation counts once more, 32 bits wide: only used when one exceeds 65535 */
    bool compact_ok = true;
    uint64_t seq_base = 0;
    bool pinned = true; /* false: ordinary memory (the batch that fills while the device is still coming up) */
    void *take(size_t bytes) const { return pinned ? ipx_host_alloc(bytes) : malloc(bytes ? bytes : 1); }
    void give(void *p) const {
        if (pinned) ipx_host_free(p); else free(p);
    }
    void allocate(size_t records, size_t cigar_words) {
        cap_rec = records;
        cap_cig = cigar_words;
        pos = (int32_t *)take(records * 4);
        flag = (uint16_t *)take(records * 2);
        n_ops = (uint16_t *)take(records * 2);
        nh = (uint8_t *)take(records);
        xs = (uint8_t *)take(records);
        cigar = (uint32_t *)take(cigar_words * 4);
        if (!pos || !flag || !n_ops || !nh || !xs || !cigar) throw runtime_error("cannot allocate pinned host memory for the record batches");
    }
    void release() {
        give(pos); give(flag); give(n_ops); give(nh); give(xs); give(cigar);
        pos = nullptr; flag = n_ops = nullptr; nh = xs = nullptr; cigar = nullptr;
    }
    void clear(uint64_t next_seq_base) {
        n = n_cig = 0;
        run_tid.clear(); run_start.clear(); ops32.clear();
        compact_ok = true;
        seq_base = next_seq_base;
    }
    bool room_for(size_t records, size_t cigar_words) const { return n + records <= cap_rec && n_cig + cigar_words <= cap_cig; }
    void begin_run(int32_t tid) {
        if (run_tid.empty() || run_tid.back() != tid) {
            run_tid.push_back(tid);
            run_start.push_back((uint64_t)n);
        }
    }
    void add_fields(int32_t tid, int32_t p, uint16_t fl, uint8_t nh_class, uint8_t x, const uint32_t *cig, uint32_t n_cigar) {
        begin_run(tid);
        pos[n] = p; flag[n] = fl; nh[n] = nh_class; xs[n] = x;
        n_ops[n] = (uint16_t)n_cigar;
        ops32.push_back(n_cigar);
        if (n_cigar > 0xFFFFu) compact_ok = false;
        memcpy(cigar + n_cig, cig, 4 * (size_t)n_cigar);
        n_cig += n_cigar;
        n++;
    }
};

class Driver {
    const Options &opts_;
    ipx_handle *h_ = nullptr;
    vector<string> targets_;            /* targets_, appended per file (save_targets, junctions_extractor.cc:564-568) */
    vector<int32_t> canon_tid_;         /* first index carrying the same name: the reference keys junctions by name */
    map<string, int32_t> name_first_;
    /* Bringing the device up (CUDA start-up, the handle, pinning the two batches) takes most of a second: a thread does
     * it while the first records are already being read and packed into batches_[2], an ordinary-memory batch.  Whatever
     * needs the handle waits for that thread (dev()). */
    PinnedBatch batches_[3];
    int cur_ = 2;
    thread init_;
    bool ready_ = false, targets_dirty_ = false;
    string init_error_;
    uint64_t next_seq_ = 0;
    uint64_t last_seq_ = 0; /* stream index and engine target of the record process() saw last */
    int32_t last_tid_ = -1;
    /* stream segments (first record index, target) of processed records, for attributing deferred errors */
    vector<pair<uint64_t, int32_t>> segments_;
    static const size_t BATCH_RECORDS = 1u << 22;
    static const size_t BATCH_CIGAR = 1u << 24;

    PinnedBatch &batch() { return batches_[cur_]; }
    int32_t map_tid(int32_t tid) const { return (tid >= 0 && (size_t)tid < canon_tid_.size()) ? canon_tid_[tid] : tid; }
    void note(int32_t tid, uint64_t n_records) {
        if (segments_.empty() || segments_.back().second != tid) segments_.emplace_back(next_seq_, tid);
        last_seq_ = next_seq_ + n_records - 1;
        last_tid_ = tid;
        next_seq_ += n_records;
    }

    void dev() {
        if (ready_) return;
        if (init_.joinable()) init_.join();
        ready_ = true;
        if (!init_error_.empty()) throw runtime_error(init_error_);
        g_trace.mark("(waited for the device)");
    }
    void sync_targets() {
        if (!targets_dirty_) return;
        targets_dirty_ = false;
        vector<const char *> ptrs(targets_.size());
        for (size_t i = 0; i < targets_.size(); i++) ptrs[i] = targets_[i].c_str();
        check_ipx(h_, ipx_set_targets(h_, (int32_t)ptrs.size(), ptrs.data()));
    }

    public:
    explicit Driver(const Options &o) : opts_(o) {
        ipx_params p;
        memset(&p, 0, sizeof(p));
        p.min_anchor_length = o.min_anchor_length;
        p.max_anchor_indel_size = o.max_anchor_indel_size;
        p.min_intron_length = o.min_intron_length;
        p.max_intron_length = o.max_intron_length;
        p.junction_filter = o.junction_filter;
        p.strandness = o.strandness;
        p.excludes = o.excludes;
        p.skip_missing_targets = o.skip_missing_targets ? 1 : 0;
        p.device = 0;
        if (const char *d = getenv("IPX_DEVICE")) p.device = atoi(d);
        batches_[2].pinned = false;
        batches_[2].allocate(BATCH_RECORDS, BATCH_CIGAR);
        batches_[2].clear(0);
        init_ = thread([this, p]() {
            try {
                if (ipx_create(&p, &h_) != IPX_OK)
                    throw runtime_error("no usable CUDA device: junction extraction needs an NVIDIA B200 (sm_100a) GPU");
                for (int k = 0; k < 2; k++) {
                    batches_[k].allocate(BATCH_RECORDS, BATCH_CIGAR);
                    batches_[k].clear(0);
                }
            } catch (const std::exception &e) {
                init_error_ = e.what();
            }
        });
    }
    ~Driver() {
        if (init_.joinable()) init_.join();
        ipx_destroy(h_);
        for (auto &b : batches_)
            if (b.pos) b.release();
    }

    void set_genome(const ipx::LoadedGenome &g) {
        dev();
        check_ipx(h_, ipx_set_genome(h_, (int32_t)g.names.size(), g.name_ptrs.data(), g.lengths.data(), g.seq_ptrs.data()));
    }

    void add_targets(const sam_hdr_t *hdr) {
        for (int i = 0; i < hdr->n_targets; i++) {
            string nm(hdr->target_name[i]);
            auto it = name_first_.find(nm);
            int32_t idx = (int32_t)targets_.size();
            if (it == name_first_.end()) {
                name_first_.emplace(nm, idx);
                canon_tid_.push_back(idx);
            } else {
                canon_tid_.push_back(it->second);
                /* keep names unique for the engine: this index is never used as a key */
                nm += string("\x01") + to_string(idx);
            }
            targets_.push_back(nm);
        }
        targets_dirty_ = true; /* handed to the engine with the next batch (the device may still be coming up) */
    }
    size_t n_targets() const { return targets_.size(); }
    uint64_t last_seq() const { return last_seq_; }
    uint64_t records_seen() const { return next_seq_; }
    int32_t last_tid() const { return last_tid_; }
    const vector<string> &targets() const { return targets_; }

    /* process_alignment (junctions_extractor.cc:533-549): hand the record to the engine */
    void process(const bam1_t *b) {
        const int32_t tid = map_tid(b->core.tid);
        if (!batch().room_for(1, b->core.n_cigar)) push(b->core.n_cigar);
        uint8_t nh, x;
        ipx::scan_tags(b, &nh, &x);
        batch().add_fields(tid, (int32_t)b->core.pos, b->core.flag, nh, x, bam_get_cigar(b), b->core.n_cigar);
        note(tid, 1);
    }
    /* the same for a BAM record in the reader's buffer (Input::peek_raw) */
    void process_raw(const uint8_t *rec, uint32_t len) {
        int32_t tid, p, l_seq;
        uint16_t n_cigar, fl;
        memcpy(&tid, rec, 4);
        memcpy(&p, rec + 4, 4);
        memcpy(&n_cigar, rec + 12, 2);
        memcpy(&fl, rec + 14, 2);
        memcpy(&l_seq, rec + 16, 4);
        tid = map_tid(tid);
        const uint8_t *cig = rec + 32 + rec[8];
        if (!batch().room_for(1, n_cigar)) push(n_cigar);
        /* NH and XS only ever matter for a read that yields a junction candidate, and that takes an N operation
         * (ReadJunctionExtractor, junctions_extractor.cc:208-229): most reads of an RNA-Seq file have none, and their
         * tags -- a dozen per record in aligner output -- need not be walked */
        bool has_intron = n_cigar > 16;
        for (unsigned k = 0; k < n_cigar && !has_intron; k++) {
            uint32_t w;
            memcpy(&w, cig + 4 * k, 4);
            has_intron = (w & 15u) == 3u;
        }
        uint8_t nh = IPX_NH_NONE, x = IPX_XS_ABSENT;
        if (has_intron) ipx::scan_tag_range(cig + 4 * (size_t)n_cigar + (size_t)((l_seq + 1) / 2) + (size_t)l_seq, rec + len, &nh, &x);
        PinnedBatch &b = batch();
        b.begin_run(tid);
        b.pos[b.n] = p; b.flag[b.n] = fl; b.nh[b.n] = nh; b.xs[b.n] = x; b.n_ops[b.n] = n_cigar;
        b.ops32.push_back(n_cigar);
        memcpy(b.cigar + b.n_cig, cig, 4 * (size_t)n_cigar); /* BAM is little-endian, and so is every host this runs on */
        b.n_cig += n_cigar;
        b.n++;
        note(tid, 1);
    }
    /* records [first, first + n) of a chunk the BAM pipeline packed: one run of the same target index */
    void process_run(const ipx::PackedChunk &c, const ipx::TidRun &run, uint32_t first, uint32_t n, uint64_t cigar_first) {
        const int32_t tid = map_tid(run.tid);
        uint32_t done = 0;
        while (done < n) {
            PinnedBatch *b = &batch();
            /* as many records as fit the batch (a single record always fits an empty one: see push) */
            uint32_t take = 0;
            uint64_t words = 0;
            while (done + take < n && b->n + take < b->cap_rec && b->n_cig + words + c.n_ops[first + done + take] <= b->cap_cig) {
                words += c.n_ops[first + done + take];
                take++;
            }
            if (take == 0) {
                push(c.n_ops[first + done]);
                continue;
            }
            b->begin_run(tid);
            memcpy(b->pos + b->n, c.pos.data() + first + done, (size_t)take * 4);
            memcpy(b->flag + b->n, c.flag.data() + first + done, (size_t)take * 2);
            memcpy(b->nh + b->n, c.nh_class.data() + first + done, take);
            memcpy(b->xs + b->n, c.xs.data() + first + done, take);
            for (uint32_t k = 0; k < take; k++) b->n_ops[b->n + k] = (uint16_t)c.n_ops[first + done + k];
            b->ops32.insert(b->ops32.end(), c.n_ops.begin() + first + done, c.n_ops.begin() + first + done + take);
            if (c.max_ops > 0xFFFFu) b->compact_ok = false;
            memcpy(b->cigar + b->n_cig, c.cigar.data() + cigar_first, (size_t)words * 4);
            b->n_cig += words;
            b->n += take;
            cigar_first += words;
            done += take;
            note(tid, take);
        }
    }
    /* hand the batch in hand to the engine (the copy to the device is asynchronous) and continue in the other one;
     * next_cigar: CIGAR words of the record that comes next (a record with more words than a batch holds gets a batch
     * of its own size) */
    void push(size_t next_cigar = 0) {
        dev();
        sync_targets();
        PinnedBatch &b = batch();
        if (b.n != 0) {
            if (b.compact_ok) {
                ipx_batch_compact v;
                memset(&v, 0, sizeof(v));
                v.n_records = b.n; v.n_cigar = b.n_cig; v.seq_base = b.seq_base;
                v.pos = b.pos; v.flag = b.flag; v.nh_class = b.nh; v.xs = b.xs; v.n_cigar_ops = b.n_ops; v.cigar = b.cigar;
                v.n_tid_runs = b.run_tid.size(); v.run_tid = b.run_tid.data(); v.run_start = b.run_start.data();
                check_ipx(h_, ipx_push_batch_compact(h_, &v));
            } else {
                /* a record with more than 65535 operations: the plain batch with 32-bit offsets */
                vector<uint32_t> off(b.n + 1, 0u);
                vector<int32_t> tid(b.n);
                for (size_t i = 0; i < b.n; i++) off[i + 1] = off[i] + b.ops32[i];
                for (size_t r = 0; r < b.run_tid.size(); r++) {
                    const size_t e = r + 1 < b.run_tid.size() ? (size_t)b.run_start[r + 1] : b.n;
                    for (size_t i = (size_t)b.run_start[r]; i < e; i++) tid[i] = b.run_tid[r];
                }
                ipx_batch v;
                memset(&v, 0, sizeof(v));
                v.n_records = b.n; v.n_cigar = b.n_cig; v.seq_base = b.seq_base;
                v.tid = tid.data(); v.pos = b.pos; v.flag = b.flag; v.nh_class = b.nh; v.xs = b.xs; v.cigar_off = off.data(); v.cigar = b.cigar;
                check_ipx(h_, ipx_push_batch(h_, &v));
                check_ipx(h_, ipx_sync_uploads(h_, 0)); /* off / tid go out of scope */
            }
            if (cur_ == 2) {
                /* the batch that filled while the device came up: ordinary memory, not used again */
                check_ipx(h_, ipx_sync_uploads(h_, 0));
                b.release();
                cur_ = 0;
            } else {
                cur_ ^= 1;
                /* the other buffer was handed over two pushes ago: its copy must have run before it is written again */
                check_ipx(h_, ipx_sync_uploads(h_, 1));
            }
        } else if (cur_ == 2) {
            b.release();
            cur_ = 0;
        }
        PinnedBatch &nb = batch();
        nb.clear(next_seq_);
        if (next_cigar > nb.cap_cig) {
            nb.release();
            nb.allocate(nb.cap_rec, next_cigar + next_cigar / 4);
        }
    }
    ipx_table *flush() {
        push();
        g_trace.mark("records read and pushed");
        ipx_table *t = nullptr;
        check_ipx(h_, ipx_flush(h_, &t));
        g_trace.mark("table flushed");
        /* get_splice_sites' warning (junctions_extractor.cc:411-415) */
        for (uint64_t i = 0; i < t->n_missing_targets; i++)
            cerr << "Warning: genomic sequence not found for " << targets_[t->missing_targets[i]]
                 << " splice junctions not available" << endl;
        return t;
    }
    int32_t target_of_seq(uint64_t seq) const {
        int32_t tid = -1;
        for (auto &s : segments_) {
            if (s.first > seq) break;
            tid = s.second;
        }
        return tid;
    }
    ipx_handle *handle() {
        dev();
        sync_targets();
        return h_;
    }
};

/* The BAM pipeline (host/bam_pipeline.h) takes a regular BAM file whenever nobody needs the records themselves (no
 * pass-through, no per-read side output).  IPX_NO_PIPELINE=1 keeps the record-at-a-time reader. */
static bool pipeline_eligible(const Input &in, const string &path) {
    if (getenv("IPX_NO_PIPELINE") != NULL || getenv("IPX_NO_RAW_PACK") != NULL) return false;
    if (hts_get_format(in.in)->format != bam) return false;
    if (path == "-" || path == "/dev/stdin") return false;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    const bool seekable = fseek(f, 0, SEEK_END) == 0;
    fclose(f);
    return seekable;
}

static int pipeline_threads() {
    if (const char *e = getenv("IPX_DECODE_THREADS")) return atoi(e) > 0 ? atoi(e) : 1;
    return 0; /* one per core */
}

/* output_junctions (intronProspector.cc:330-351) for table rows [begin, end) */
static void output_junctions(const ipx_table *t, const vector<string> &targets, uint64_t begin, uint64_t end,
                             float min_confidence, Outputs &out) {
    vector<const char *> names(targets.size());
    for (size_t i = 0; i < targets.size(); i++) names[i] = targets[i].c_str();
    auto emit = [&](unique_ptr<OutFile> &f, int fmt) {
        if (!f) return;
        ipx_text txt;
        int rc = ipx_format(t, names.data(), fmt, min_confidence, begin, end, &txt);
        if (rc != IPX_OK) throw runtime_error("formatting failed");
        f->write(txt);
        ipx_text_free(&txt);
    };
    emit(out.junction_bed, IPX_FMT_JUNCTION_BED);
    emit(out.intron_bed, IPX_FMT_INTRON_BED9);
    emit(out.intron_bed6, IPX_FMT_INTRON_BED6);
    emit(out.intron_call, IPX_FMT_INTRON_TSV);
}

/* extract_junctions_by_target (intronProspector.cc:356-371) with identify_junctions_for_target /
 * process_target_alignment (junctions_extractor.cc:637-664) */
static void extract_by_target(Driver &drv, const Options &opts, Outputs &out, Annotator &ann) {
    Input in(opts.bam_files[0]);
    drv.add_targets(in.hdr);
    ann.input_opened();
    if (opts.bam_pass_through != "") in.open_pass_through(opts.bam_pass_through);
    if (!ann.active() && in.out == NULL) in.enable_raw();
    const int n_targets = (int)drv.n_targets();
    int previous_target = -1;
    bool pending = false, eof = false;
    int stop_target = n_targets; /* targets [0, stop_target) get their output */
    string fatal;
    const bool fast = !ann.active() && in.out == NULL && pipeline_eligible(in, opts.bam_files[0]);
    if (fast) {
        /* The same selection of records as the loops below, on runs of equal target index: target t takes the records
         * of index t; a record of another index ends target t (it is "not coordinate sorted" when its index lies below
         * the target before), and stays in hand for the targets that follow. */
        ipx::BamPipeline pipe;
        if (!pipe.open(opts.bam_files[0].c_str(), ipx_bam_header_bytes(in.in), pipeline_threads()))
            throw runtime_error("Error opening BAM/SAM/CRAM file: " + opts.bam_files[0]);
        int target = 0;
        bool stopped = n_targets == 0, damaged = false;
        ipx::PackedChunk *c;
        while (!stopped && !damaged && (c = pipe.next()) != NULL) {
            for (size_t r = 0; r < c->runs.size() && !stopped; r++) {
                const ipx::TidRun &run = c->runs[r];
                while (!stopped && run.tid != target) {
                    if (run.tid >= 0 && run.tid < previous_target) {
                        fatal = "BAM appears to not be coordinate sorted";
                        stop_target = target;
                        stopped = true;
                        break;
                    }
                    previous_target = target;
                    if (++target >= n_targets) stopped = true;
                }
                if (!stopped) drv.process_run(*c, run, run.first, run.n, run.cigar_first);
            }
            damaged = c->bad && !stopped;
            pipe.release(c);
        }
        if (damaged) throw runtime_error("Error reader BAM record: " + opts.bam_files[0]);
    }
    /* the record in hand: in the reader's buffer (not consumed until it is processed) or in in.rec */
    bool cur_raw = false;
    const uint8_t *cur_bytes = NULL;
    uint32_t cur_len = 0;
    int32_t cur_tid = -1;
    auto fetch = [&]() -> bool {
        if (in.peek_raw(&cur_bytes, &cur_len)) {
            cur_raw = true;
            memcpy(&cur_tid, cur_bytes, 4);
            return true;
        }
        if (!in.read()) return false;
        cur_raw = false;
        cur_tid = in.rec->core.tid;
        return true;
    };
    uint64_t flushed_records = 0; /* records the engine had seen at the last per-target flush */
    for (int target = 0; !fast && target < n_targets && fatal.empty(); target++) {
        for (;;) {
            if (!pending) {
                if (eof || !fetch()) {
                    eof = true;
                    break;
                }
            }
            pending = false;
            if (cur_tid != target) {
                if ((cur_tid >= 0) and (cur_tid < previous_target)) {
                    fatal = "BAM appears to not be coordinate sorted";
                    stop_target = target;
                    break;
                }
                pending = true;
                break;
            }
            if (cur_raw) {
                drv.process_raw(cur_bytes, cur_len);
                in.consume_raw();
                continue;
            }
            const bam1_t *aln = in.rec;
            drv.process(aln);
            if (ann.active()) ann.keep(aln, drv.last_seq(), drv.last_tid(), true, true); /* written after the flush */
            else in.pass(aln);
        }
        if (ann.active() && fatal.empty() && drv.records_seen() != flushed_records) {
            flushed_records = drv.records_seen();
            /* With the per-read side outputs the records are kept until their junctions are known: flush target by target,
             * as the reference does (intronProspector.cc:365-369), so that no more than one target's records are held.
             * Junction keys never cross targets and sjN keeps counting across flushes. */
            ipx_table *tt = drv.flush();
            ann.finish(tt, drv.targets(), in.out != NULL ? &in : NULL);
            if (tt->status != IPX_OK) {
                fatal = tt->error_text;
                stop_target = target;
            } else if (tt->n_junctions) {
                output_junctions(tt, drv.targets(), 0, tt->n_junctions, opts.min_confidence_score, out);
            }
            ipx_table_free(tt);
        }
        previous_target = target;
    }
    ipx_table *t = drv.flush();
    if (ann.active()) ann.finish(t, drv.targets(), in.out != NULL ? &in : NULL);
    if (t->status != IPX_OK) {
        /* the reference throws while processing the failing record: earlier targets are already written */
        int32_t err_tid = t->error_tid >= 0 ? t->error_tid : drv.target_of_seq(t->first_error_seq);
        if (fatal.empty() || err_tid < stop_target) {
            fatal = t->error_text;
            stop_target = err_tid < 0 ? 0 : err_tid;
        }
    }
    /* per target, in header order (intronProspector.cc:365-369) */
    vector<pair<uint64_t, uint64_t>> ranges(n_targets, make_pair((uint64_t)0, (uint64_t)0));
    {
        uint64_t i = 0;
        while (i < t->n_junctions) {
            uint64_t e = i;
            while (e < t->n_junctions && t->chrom[e] == t->chrom[i]) e++;
            if (t->chrom[i] >= 0 && t->chrom[i] < n_targets) ranges[t->chrom[i]] = make_pair(i, e);
            i = e;
        }
    }
    for (int target = 0; target < stop_target && target < n_targets; target++)
        if (ranges[target].second > ranges[target].first)
            output_junctions(t, drv.targets(), ranges[target].first, ranges[target].second, opts.min_confidence_score, out);
    ipx_table_free(t);
    if (!fatal.empty()) throw runtime_error(fatal);
    /* copy_unaligned_reads (junctions_extractor.cc:667-674) */
    if (in.out != NULL) {
        if (pending) in.pass(in.rec);
        while (!eof && in.read()) in.pass(in.rec);
    }
}

/* extract_junctions_unsorted / extract_junctions_multiple (intronProspector.cc:375-401) */
static void extract_whole_files(Driver &drv, const Options &opts, Outputs &out, Annotator &ann) {
    for (size_t ibam = 0; ibam < opts.bam_files.size(); ibam++) {
        Input in(opts.bam_files[ibam]);
        drv.add_targets(in.hdr);
        ann.input_opened();
        /* in --unsorted mode the reference opens the pass-through file but only copy_unaligned_reads writes to it,
         * after every record has been consumed: the file gets the header only */
        if (opts.bam_files.size() == 1 && opts.bam_pass_through != "") in.open_pass_through(opts.bam_pass_through);
        if (!ann.active()) in.enable_raw();
        if (!ann.active() && in.out == NULL && pipeline_eligible(in, opts.bam_files[ibam])) {
            ipx::BamPipeline pipe;
            if (!pipe.open(opts.bam_files[ibam].c_str(), ipx_bam_header_bytes(in.in), pipeline_threads()))
                throw runtime_error("Error opening BAM/SAM/CRAM file: " + opts.bam_files[ibam]);
            bool damaged = false;
            ipx::PackedChunk *c;
            while (!damaged && (c = pipe.next()) != NULL) {
                for (const ipx::TidRun &run : c->runs) drv.process_run(*c, run, run.first, run.n, run.cigar_first);
                damaged = c->bad;
                pipe.release(c);
            }
            if (damaged) throw runtime_error("Error reader BAM record: " + opts.bam_files[ibam]);
            continue;
        }
        for (;;) {
            const uint8_t *bytes;
            uint32_t len;
            if (in.peek_raw(&bytes, &len)) {
                drv.process_raw(bytes, len);
                in.consume_raw();
                continue;
            }
            if (!in.read()) break;
            drv.process(in.rec);
            if (ann.active()) ann.keep(in.rec, drv.last_seq(), drv.last_tid(), true, false);
        }
    }
    ipx_table *t = drv.flush();
    if (ann.active()) ann.finish(t, drv.targets(), NULL);
    if (t->status != IPX_OK) {
        string msg = t->error_text;
        ipx_table_free(t);
        throw runtime_error(msg);
    }
    output_junctions(t, drv.targets(), 0, t->n_junctions, opts.min_confidence_score, out);
    ipx_table_free(t);
}

static void extract_junctions(const Options &opts) {
    Driver drv(opts); /* (the device comes up while the genome is read) */
    ipx::LoadedGenome genome;
    if (opts.genome_fa.size() > 0) ipx::load_genome(opts.genome_fa, genome);
    g_trace.mark("genome loaded");
    Annotator ann(opts);
    if (opts.genome_fa.size() > 0) drv.set_genome(genome);
    Outputs out(opts);
    if (opts.bam_files.size() > 1 || opts.unsorted) {
        extract_whole_files(drv, opts, out, ann);
    } else {
        extract_by_target(drv, opts, out, ann);
    }
    g_trace.mark("outputs written");
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
        extract_junctions(opts);
    } catch (const std::exception &e) {
        cmd_error(e.what());
    }
    return 0;
}
