/* ipx_oracle_abi.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Presents the CPU oracle (ipx_oracle.c) under the symbol names of the C ABI in include/ipx.h, so that the
 * product's command line tools can be LINKED, by oracle/Makefile and for tests only, against the oracle
 * instead of libipx.so.  The resulting binaries (the .oracle files under oracle/_ref) let the host-side logic of the CLIs
 * (option handling, run modes, record selection, text output) be checked against the reference's goldens on
 * a machine without a GPU, and give parity tests a second, independent producer of whole output files.
 * Nothing in intronprospector_b200/ references this file; the product libraries never load it.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include "ipx_oracle.h"

struct ipx_handle {
    ipxo_state *st;
    size_t log_reported; /* how much of the oracle's stderr text has been handed out as merge conflict log */
};

const char *ipx_version(void) { return "ipx CPU oracle (test only)"; }
int ipx_abi_version(void) { return IPX_ABI_VERSION; }
int ipx_device_count(void) { return 0; }

int ipx_create(const ipx_params *params, ipx_handle **out) {
    ipx_handle *h = (ipx_handle *)calloc(1, sizeof(*h));
    h->st = ipxo_create(params);
    *out = h;
    return IPX_OK;
}

void ipx_destroy(ipx_handle *h) {
    if (!h) return;
    ipxo_destroy(h->st);
    free(h);
}

const char *ipx_last_error(const ipx_handle *h) { return ipxo_last_error(h->st); }
void *ipx_stream(ipx_handle *h) { (void)h; return NULL; }
int ipx_set_targets(ipx_handle *h, int32_t n, const char *const *names) { return ipxo_set_targets(h->st, n, names); }
int ipx_set_genome(ipx_handle *h, int32_t n, const char *const *names, const uint64_t *lengths, const char *const *seqs) {
    return ipxo_set_genome(h->st, n, names, lengths, seqs);
}
int ipx_push_batch(ipx_handle *h, const ipx_batch *b) { return ipxo_push_batch(h->st, b); }
int ipx_push_batch_device(ipx_handle *h, const ipx_batch *b) { (void)h; (void)b; return IPX_ERR_CUDA; }
int ipx_finalize_device(ipx_handle *h) { (void)h; return IPX_OK; }
/* ipx_batch_compact -> ipx_batch on the host (the product expands it on the device) */
int ipx_push_batch_compact(ipx_handle *h, const ipx_batch_compact *c) {
    if (!h || !c) return IPX_ERR_ARG;
    if (c->n_records == 0) return IPX_OK;
    const uint64_t n = c->n_records;
    int32_t *tid = (int32_t *)malloc(n * sizeof(int32_t));
    uint32_t *off = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    uint64_t run = 0;
    off[0] = 0;
    for (uint64_t i = 0; i < n; i++) {
        while (run + 1 < c->n_tid_runs && c->run_start[run + 1] <= i) run++;
        tid[i] = c->run_tid[run];
        off[i + 1] = off[i] + c->n_cigar_ops[i];
    }
    ipx_batch b;
    b.n_records = n; b.n_cigar = c->n_cigar; b.seq_base = c->seq_base;
    b.tid = tid; b.pos = c->pos; b.flag = c->flag; b.nh_class = c->nh_class; b.xs = c->xs; b.cigar_off = off; b.cigar = c->cigar;
    int rc = ipxo_push_batch(h->st, &b);
    free(tid);
    free(off);
    return rc;
}

int ipx_flush(ipx_handle *h, ipx_table **out) { return ipxo_flush(h->st, out); }
void ipx_table_free(ipx_table *t) { ipxo_table_free(t); }
int ipx_clear(ipx_handle *h) { ipx_table *t = NULL; ipxo_flush(h->st, &t); ipxo_table_free(t); return IPX_OK; }

int ipx_merge_rows_fold(ipx_handle *h, const ipx_merge_rows *rows, ipx_table **out, ipx_text *conflict_log) {
    int rc = ipxo_merge_rows_fold(h->st, rows, out);
    if (conflict_log) {
        const char *all = ipxo_stderr_text(h->st);
        conflict_log->data = strdup(all + h->log_reported);
        conflict_log->len = strlen(conflict_log->data);
        h->log_reported = strlen(all);
    }
    return rc;
}

/* one key range of a sharded merge: the oracle only ever folds everything at once (ipx_device_count() is 0 here), so the
 * lines of its log are already in input order -- they are numbered as they come */
int ipx_merge_rows_fold_part(ipx_handle *h, const ipx_merge_rows *rows, const uint64_t *row_ids, ipx_table **out,
                             ipx_text *conflict_log, uint64_t **conflict_rows, uint64_t *n_conflict_rows, uint64_t *fail_row) {
    (void)row_ids;
    ipx_text log = {NULL, 0};
    int rc = ipx_merge_rows_fold(h, rows, out, &log);
    uint64_t n_lines = 0;
    for (size_t i = 0; i < log.len; i++) n_lines += log.data[i] == '\n';
    if (conflict_rows && n_conflict_rows) {
        *conflict_rows = (uint64_t *)malloc((n_lines ? n_lines : 1) * sizeof(uint64_t));
        for (uint64_t i = 0; i < n_lines; i++) (*conflict_rows)[i] = i;
        *n_conflict_rows = n_lines;
    }
    if (fail_row) *fail_row = rc == IPX_ERR_MERGE ? n_lines : ~0ull;
    if (conflict_log) *conflict_log = log; else ipxo_text_free(&log);
    return rc;
}
void ipx_free(void *p) { free(p); }
/* the device tokeniser: not an input this build takes -- the tools then read the rows with the host parser */
int ipx_tsv_tokenize(int32_t device, const char *bytes, uint64_t n_bytes, const ipx_tsv_columns *columns, ipx_tsv_rows **out) {
    (void)device; (void)bytes; (void)n_bytes; (void)columns;
    if (out) *out = NULL;
    return IPX_ERR_ARG;
}
void ipx_tsv_rows_free(ipx_tsv_rows *rows) { (void)rows; }
int ipx_table_set_ijunc(ipx_table *t, const uint32_t *ijunc) {
    if (!t || (t->n_junctions && !ijunc)) return IPX_ERR_ARG;
    memcpy((uint32_t *)t->ijunc, ijunc, t->n_junctions * sizeof(uint32_t)); /* (the oracle's tables own their columns) */
    return IPX_OK;
}

int ipx_format(const ipx_table *t, const char *const *names, int format, float min_confidence, uint64_t begin, uint64_t end,
               ipx_text *out) {
    return ipxo_format(t, names, format, min_confidence, begin, end, out);
}
int ipx_format_tsv_header(ipx_text *out) { return ipxo_format_tsv_header(out); }
void ipx_text_free(ipx_text *t) { ipxo_text_free(t); }
int ipx_get_timings(ipx_handle *h, ipx_timings *out) { (void)h; memset(out, 0, sizeof(*out)); return IPX_OK; }
int ipx_reset_timings(ipx_handle *h) { (void)h; return IPX_OK; }
/* pinned batches of the product's CLI: plain memory and nothing to wait for here */
void *ipx_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void ipx_host_free(void *p) { free(p); }
int ipx_sync_uploads(ipx_handle *h, uint32_t keep_in_flight) { (void)h; (void)keep_in_flight; return IPX_OK; }
