/* calls_tsv.h -- intron-calls TSV files -> merge rows (host side of ipx_merge_rows_fold). */
#ifndef IPX_CALLS_TSV_H
#define IPX_CALLS_TSV_H
#include <stdint.h>
#include <string>
#include <unordered_map>
#include <vector>
#include "ipx.h"

namespace ipx {

struct CallsRows {
    std::vector<int32_t> chrom;
    std::vector<uint32_t> i0, i1, a0, a1, sites, counts;
    std::vector<char> strand;
    std::vector<float> conf;
    std::vector<std::string> chrom_names, site_strings;
    std::vector<const char *> chrom_ptrs, site_ptrs;
    std::unordered_map<std::string, int32_t> chrom_index;
    std::unordered_map<std::string, uint32_t> site_index;
    ipx_merge_rows rows;
    CallsRows();
    void bind();
};

/* process_calls_tsv (intronProspectorMerge.cc:190-199) without the fold: appends the rows of one file.
 * Throws std::runtime_error with the reference's messages (tsv.hh:41,71,88; type_ops.hh:23,32,41). */
void load_calls_tsv(const std::string &path, int junction_filter, CallsRows &out);

/* The same for a list of files, in list order, with the files parsed by a pool of threads (IPX_PARSE_THREADS, default:
 * the host's cores, at most 16).  Rows, chromosome and splice-site dictionaries come out exactly as from calling
 * load_calls_tsv file by file; so does a failure: the rows up to the first bad line of the first bad file are kept and
 * its error is thrown. */
void load_calls_tsvs(const std::vector<std::string> &paths, int junction_filter, CallsRows &out);

/* The same with the rows tokenised on the device (ipx_tsv_tokenize of include/ipx.h, handed in as functions so that this
 * file needs no CUDA): a file is read (and inflated) whole, its header is looked at here, the rows come back as columns;
 * what the device flags as not being in the plain form is parsed here, line by line, with the rules above -- so rows,
 * dictionaries and errors are those of load_calls_tsvs.  tokenize() returning IPX_ERR_ARG means "not an input for the
 * device" (the file is then parsed here); any other failure is an error. */
struct TsvTokenizer {
    void *(*alloc)(size_t);  /* memory for a file's bytes (pinned memory crosses the link faster) */
    void (*release)(void *);
    int (*tokenize)(const char *bytes, uint64_t n_bytes, const ipx_tsv_columns *columns, ipx_tsv_rows **out);
    void (*rows_free)(ipx_tsv_rows *);
    bool (*ready)(); /* optional: false while the device is not up yet -- files that come by meanwhile are parsed here */
};
void load_calls_tsvs(const std::vector<std::string> &paths, int junction_filter, CallsRows &out, const TsvTokenizer &tok);

}  // namespace ipx
#endif
