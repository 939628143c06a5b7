/* ipx_io_api.cc -- C ABI of the host I/O library, see include/ipx_io.h */
#include "ipx_io.h"
#include <string.h>
#include <stdexcept>
#include <string>
#include "calls_tsv.h"
#include "genome_loader.h"
#include "htslib/faidx.h"
#include "htslib/sam.h"
#include "bam_pipeline.h"
#include "record_source.h"

static thread_local std::string g_err;

struct ipxio_reader {
    samFile *fp = nullptr;
    sam_hdr_t *hdr = nullptr;
    bam1_t *rec = nullptr;
    bool pending = false;
};
struct ipxio_genome {
    ipx::LoadedGenome g;
};
struct ipxio_calls {
    ipx::CallsRows rows;
};

extern "C" const char *ipxio_last_error(void) { return g_err.c_str(); }

extern "C" int ipxio_open(const char *path, ipxio_reader **out) {
    *out = nullptr;
    samFile *fp = sam_open(path, "r");
    if (!fp) {
        g_err = std::string("Error opening BAM/SAM/CRAM file: ") + path;
        return IPX_ERR_ARG;
    }
    sam_hdr_t *hdr = sam_hdr_read(fp);
    if (!hdr) {
        sam_close(fp);
        g_err = std::string("Error reading header: ") + path;
        return IPX_ERR_ARG;
    }
    ipxio_reader *r = new ipxio_reader();
    r->fp = fp;
    r->hdr = hdr;
    r->rec = bam_init1();
    *out = r;
    return IPX_OK;
}

extern "C" void ipxio_close(ipxio_reader *r) {
    if (!r) return;
    bam_destroy1(r->rec);
    sam_hdr_destroy(r->hdr);
    sam_close(r->fp);
    delete r;
}

extern "C" int32_t ipxio_n_targets(const ipxio_reader *r) { return r->hdr->n_targets; }
extern "C" const char *ipxio_target_name(const ipxio_reader *r, int32_t i) { return r->hdr->target_name[i]; }

extern "C" int64_t ipxio_read_batch(ipxio_reader *r, uint64_t max_records, uint64_t max_cigar, int32_t *tid, int32_t *pos,
                                    uint16_t *flag, uint8_t *nh_class, uint8_t *xs, uint32_t *cigar_off, uint32_t *cigar,
                                    uint64_t *n_cigar_out) {
    uint64_t n = 0, nc = 0;
    cigar_off[0] = 0;
    while (n < max_records) {
        if (!r->pending) {
            int rc = sam_read1(r->fp, r->hdr, r->rec);
            if (rc == -1) break;
            if (rc < -1) {
                g_err = "Error reader BAM record";
                return -1;
            }
        }
        const bam1_t *b = r->rec;
        if (nc + b->core.n_cigar > max_cigar) {
            if (n == 0) {
                g_err = "record has more CIGAR operations than the batch can hold";
                return -1;
            }
            r->pending = true;
            break;
        }
        r->pending = false;
        tid[n] = b->core.tid;
        pos[n] = (int32_t)b->core.pos;
        flag[n] = b->core.flag;
        ipx::scan_tags(b, &nh_class[n], &xs[n]);
        memcpy(cigar + nc, bam_get_cigar(b), (size_t)b->core.n_cigar * 4);
        nc += b->core.n_cigar;
        n++;
        cigar_off[n] = (uint32_t)nc;
    }
    if (n_cigar_out) *n_cigar_out = nc;
    return (int64_t)n;
}

/* The whole file through the BAM pipeline (what the command-line tool does with a BAM file): the same columns as
 * ipxio_read_batch over all records, into caller buffers sized max_records / max_cigar.  Returns the number of records,
 * -1 on a damaged file (after the records before the damage), -2 when the buffers are too small. */
extern "C" int64_t ipxio_pipeline_read(const char *path, int32_t threads, uint64_t max_records, uint64_t max_cigar, int32_t *tid,
                                       int32_t *pos, uint16_t *flag, uint8_t *nh_class, uint8_t *xs, uint32_t *cigar_off, uint32_t *cigar,
                                       uint64_t *n_cigar_out) {
    samFile *fp = sam_open(path, "r");
    if (!fp) {
        g_err = std::string("Error opening BAM/SAM/CRAM file: ") + path;
        return -1;
    }
    sam_hdr_t *hdr = sam_hdr_read(fp);
    if (!hdr || hts_get_format(fp)->format != bam) {
        if (hdr) sam_hdr_destroy(hdr);
        sam_close(fp);
        g_err = std::string("not a BAM file: ") + path;
        return -1;
    }
    const uint64_t skip = ipx_bam_header_bytes(fp);
    sam_hdr_destroy(hdr);
    sam_close(fp);
    ipx::BamPipeline pipe;
    if (!pipe.open(path, skip, threads)) {
        g_err = std::string("Error opening BAM/SAM/CRAM file: ") + path;
        return -1;
    }
    uint64_t n = 0, nc = 0;
    cigar_off[0] = 0;
    bool damaged = false, full = false;
    ipx::PackedChunk *c;
    while (!damaged && !full && (c = pipe.next()) != nullptr) {
        if (n + c->size() > max_records || nc + c->cigar.size() > max_cigar) {
            full = true;
        } else {
            for (const ipx::TidRun &run : c->runs)
                for (uint32_t i = 0; i < run.n; i++) tid[n + run.first + i] = run.tid;
            for (size_t i = 0; i < c->size(); i++) {
                pos[n + i] = c->pos[i];
                flag[n + i] = c->flag[i];
                nh_class[n + i] = c->nh_class[i];
                xs[n + i] = c->xs[i];
                cigar_off[n + i + 1] = cigar_off[n + i] + c->n_ops[i];
            }
            memcpy(cigar + nc, c->cigar.data(), c->cigar.size() * 4);
            n += c->size();
            nc += c->cigar.size();
            damaged = c->bad;
        }
        pipe.release(c);
    }
    if (n_cigar_out) *n_cigar_out = nc;
    if (full) {
        g_err = "the buffers are too small for the file";
        return -2;
    }
    if (damaged) {
        g_err = "Error reader BAM record";
        return -1;
    }
    return (int64_t)n;
}

extern "C" int ipxio_genome_load(const char *fasta, ipxio_genome **out) {
    *out = nullptr;
    ipxio_genome *g = new ipxio_genome();
    try {
        ipx::load_genome(fasta, g->g);
    } catch (const std::exception &e) {
        g_err = e.what();
        delete g;
        return IPX_ERR_GENOME;
    }
    *out = g;
    return IPX_OK;
}

extern "C" void ipxio_genome_free(ipxio_genome *g) { delete g; }
extern "C" int32_t ipxio_genome_n_seqs(const ipxio_genome *g) { return (int32_t)g->g.names.size(); }
extern "C" const char *const *ipxio_genome_names(const ipxio_genome *g) { return g->g.name_ptrs.data(); }
extern "C" const uint64_t *ipxio_genome_lengths(const ipxio_genome *g) { return g->g.lengths.data(); }
extern "C" const char *const *ipxio_genome_seqs(const ipxio_genome *g) { return g->g.seq_ptrs.data(); }
extern "C" int ipxio_faidx_build(const char *fasta) { return fai_build(fasta) < 0 ? IPX_ERR_ARG : IPX_OK; }

extern "C" int ipxio_calls_load(const char *const *paths, int32_t n_paths, int32_t junction_filter, ipxio_calls **out) {
    ipxio_calls *c = new ipxio_calls();
    *out = c;
    int rc = IPX_OK;
    try {
        std::vector<std::string> files;
        for (int32_t i = 0; i < n_paths; i++) files.push_back(paths[i]);
        ipx::load_calls_tsvs(files, junction_filter, c->rows);
    } catch (const std::exception &e) {
        g_err = e.what();
        rc = IPX_ERR_ARG;
    }
    c->rows.bind();
    return rc;
}

extern "C" int ipxio_calls_load_tokenized(const char *const *paths, int32_t n_paths, int32_t junction_filter, ipxio_tokenize_fn tokenize,
                                          ipxio_rows_free_fn rows_free, ipxio_calls **out) {
    if (!tokenize || !rows_free) return IPX_ERR_ARG;
    ipxio_calls *c = new ipxio_calls();
    *out = c;
    int rc = IPX_OK;
    try {
        std::vector<std::string> files;
        for (int32_t i = 0; i < n_paths; i++) files.push_back(paths[i]);
        const ipx::TsvTokenizer tok{malloc, free, tokenize, rows_free, nullptr};
        ipx::load_calls_tsvs(files, junction_filter, c->rows, tok);
    } catch (const std::exception &e) {
        g_err = e.what();
        rc = IPX_ERR_ARG;
    }
    c->rows.bind();
    return rc;
}

extern "C" void ipxio_calls_free(ipxio_calls *c) { delete c; }
extern "C" const ipx_merge_rows *ipxio_calls_rows(const ipxio_calls *c) { return &c->rows.rows; }
extern "C" int32_t ipxio_calls_n_chroms(const ipxio_calls *c) { return (int32_t)c->rows.chrom_names.size(); }
extern "C" const char *const *ipxio_calls_chroms(const ipxio_calls *c) { return c->rows.chrom_ptrs.data(); }
