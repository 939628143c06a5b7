/* ipx_oracle.h -- CPU oracle for the junction-extraction path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded, sequential restatement of the reference algorithm
 * (diekhans/intronProspector 1.4.0); every function in ipx_oracle.c cites the reference
 * file:line it follows.  It consumes the same columnar batches and produces the same
 * table layout as the C ABI in include/ipx.h (the struct definitions are shared, nothing
 * else is), so a parity test is "same batch in, same table out".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (intronprospector_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_goldens.py checks this oracle against every
 * extraction and merge golden of the reference's own test-suite, and tests/test_oracle_vs_ref.py
 * against the unmodified reference binary (oracle/_ref) on seeded synthetic inputs.
 */
#ifndef IPX_ORACLE_H
#define IPX_ORACLE_H
#include "../include/ipx.h"

typedef struct ipxo_state ipxo_state;

ipxo_state *ipxo_create(const ipx_params *params);
void ipxo_destroy(ipxo_state *st);
const char *ipxo_last_error(const ipxo_state *st);
/* stderr text the reference would have produced so far (missing-genome warnings, merge conflicts) */
const char *ipxo_stderr_text(const ipxo_state *st);
int ipxo_set_targets(ipxo_state *st, int32_t n, const char *const *names);
int ipxo_set_genome(ipxo_state *st, int32_t n, const char *const *names, const uint64_t *lengths,
                    const char *const *seqs);
int ipxo_push_batch(ipxo_state *st, const ipx_batch *batch);
int ipxo_flush(ipxo_state *st, ipx_table **out);
void ipxo_table_free(ipx_table *t);
int ipxo_merge_rows_fold(ipxo_state *st, const ipx_merge_rows *rows, ipx_table **out);
int ipxo_format(const ipx_table *t, const char *const *names, int format, float min_confidence,
                uint64_t begin, uint64_t end, ipx_text *out);
int ipxo_format_tsv_header(ipx_text *out);
void ipxo_text_free(ipx_text *t);
/* number of junction events that passed junction_qc so far (for algorithmic-byte accounting) */
uint64_t ipxo_event_count(const ipxo_state *st);
uint64_t ipxo_cigar_live(const ipxo_state *st);
#endif
