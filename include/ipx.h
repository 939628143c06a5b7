/* ipx.h -- C ABI of the B200 junction-extraction library (libipx.so).
 *
 * The reference (diekhans/intronProspector 1.4.0) has no plugin or FFI interface;
 * its hot path sits behind a C++ class seam.  This header is that seam as a C ABI:
 * every entry point names the reference interface it replaces (file:line into the
 * reference's src/).  The library is CUDA-only (sm_100a); there is no CPU fallback:
 * every compute entry point fails with IPX_ERR_CUDA when no device is usable.
 *
 * Conventions: all functions return IPX_OK (0) or a negative IPX_ERR_* code; the
 * text of the last error on a handle is ipx_last_error().  A handle is used from
 * one host thread at a time (the reference is single-threaded).  Buffers returned
 * by the library (ipx_table, ipx_text) stay owned by it until the matching *_free.
 */
#ifndef IPX_H
#define IPX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPX_ABI_VERSION 2

#define IPX_OK 0
#define IPX_ERR_ARG (-1)       /* bad argument */
#define IPX_ERR_CUDA (-2)      /* CUDA runtime failure or no usable device */
#define IPX_ERR_CIGAR (-3)     /* unknown CIGAR operation (junctions_extractor.cc:312-313) */
#define IPX_ERR_GENOME (-4)    /* genome fetch failure (genome.cc:64-70) */
#define IPX_ERR_CAPACITY (-5)  /* key space cannot hold (targets x max intron length) */
#define IPX_ERR_MERGE (-6)     /* merge: rows of one intron disagree on splice sites (junctions.cc:91-92) */
#define IPX_ERR_STATE (-7)     /* call sequence error */

/* junction filter, reference splice_juncs.hh:13-17 */
#define IPX_FILTER_NULL 0
#define IPX_FILTER_ALL 1
#define IPX_FILTER_CANON 2
/* strandness, reference junctions_extractor.hh (Strandness): UN, RF, FR */
#define IPX_STRAND_UN 0
#define IPX_STRAND_RF 1
#define IPX_STRAND_FR 2
/* excludes bit set, reference junctions_extractor.hh (EXCLUDE_MULTI) */
#define IPX_EXCLUDE_MULTI 1u

/* read categories, reference junctions.hh:49-54 */
#define IPX_CAT_UNIQ 0
#define IPX_CAT_MULTI 1
#define IPX_CAT_UNSURE 2

/* nh_class column values: what get_num_aligns() (junctions_extractor.cc:49-56) can distinguish */
#define IPX_NH_NONE 0 /* NH absent, not an integer type, or <= 0 */
#define IPX_NH_ONE 1  /* NH == 1 */
#define IPX_NH_MANY 2 /* NH > 1 */

/* xs column values: 0 = no XS tag; IPX_XS_OTHER = XS present but not a non-NUL 'A' value
 * (get_junction_strand_XS, junctions_extractor.cc:322-330, yields '.'); else the XS:A character */
#define IPX_XS_ABSENT 0
#define IPX_XS_OTHER 0xFF

typedef struct ipx_handle ipx_handle;

/* Constructor arguments of JunctionsExtractor (junctions_extractor.hh:176-207) that
 * influence results, after the CLI's defaulting rules (intronProspector.cc:274-292). */
typedef struct ipx_params {
    uint32_t min_anchor_length;     /* -a, default 8 */
    uint32_t max_anchor_indel_size; /* 0 unless --allow-anchor-indels, then -m (default 36) */
    uint32_t min_intron_length;     /* -i, default 70 */
    uint32_t max_intron_length;     /* -I, default 500000 */
    int32_t junction_filter;        /* IPX_FILTER_* (CANON is the CLI default when a genome is given) */
    int32_t strandness;             /* IPX_STRAND_* */
    uint32_t excludes;              /* IPX_EXCLUDE_* */
    int32_t skip_missing_targets;   /* -S */
    int32_t device;                 /* CUDA device ordinal */
    int32_t profile;                /* non-zero: record CUDA events around every stage (ipx_get_timings) */
    void *stream;                   /* cudaStream_t to launch on; NULL = library-owned stream */
    uint64_t table_capacity;        /* junction hash slots (power of two; at most half may be used), 0 = default 4 M */
    uint64_t hist_capacity;         /* entries of the read-offset log (8 bytes each; one per read whose offset the
                                       256-wide dense rows do not take), 0 = default 64 M */
    int32_t deferred_counting;      /* multi-GPU runs with the canon filter: ipx_push_batch only records which reads could
                                       create each junction and keeps the batch's junction events on the device; they are
                                       counted by ipx_count_deferred once the creation thresholds of all ranks are known */
} ipx_params;

/* One batch of decoded alignment records in stream order, structure-of-arrays.
 * Replaces the bam1_t the reference hands to process_alignment (junctions_extractor.cc:533-549).
 * tid indexes the names given to ipx_set_targets(); MAPQ is not part of the path (the reference
 * never reads it).  cigar words are BAM-encoded (len << 4 | op, op index into "MIDNSHP=XB"). */
typedef struct ipx_batch {
    uint64_t n_records;
    uint64_t n_cigar;          /* total CIGAR words in `cigar` (< 2^32) */
    uint64_t seq_base;         /* stream index of record 0 (monotonic across batches and input files) */
    const int32_t *tid;        /* [n_records] target index, < 0 = unmapped */
    const int32_t *pos;        /* [n_records] 0-based leftmost position */
    const uint16_t *flag;      /* [n_records] SAM flag */
    const uint8_t *nh_class;   /* [n_records] IPX_NH_* */
    const uint8_t *xs;         /* [n_records] IPX_XS_* or the XS:A character */
    const uint32_t *cigar_off; /* [n_records + 1] start of each record's words in `cigar` */
    const uint32_t *cigar;     /* [n_cigar] */
} ipx_batch;

/* The same batch in the form the host link likes best (the end-to-end path is bound by it): what BAM stores per record
 * -- the number of CIGAR operations, 16 bits -- instead of 32-bit offsets, and the target index as runs (a coordinate
 * sorted file has one run per chromosome) instead of a column: 10 bytes per record + 4 per CIGAR word.  The library
 * expands both on the device. */
typedef struct ipx_batch_compact {
    uint64_t n_records;
    uint64_t n_cigar;            /* total CIGAR words in `cigar` (< 2^32) = sum of n_cigar_ops */
    uint64_t seq_base;           /* stream index of record 0 */
    const int32_t *pos;          /* [n_records] */
    const uint16_t *flag;        /* [n_records] */
    const uint8_t *nh_class;     /* [n_records] IPX_NH_* */
    const uint8_t *xs;           /* [n_records] IPX_XS_* or the XS:A character */
    const uint16_t *n_cigar_ops; /* [n_records] CIGAR words of each record (records with more than 65535 need ipx_batch) */
    const uint32_t *cigar;       /* [n_cigar] */
    uint64_t n_tid_runs;         /* runs of consecutive records with the same target index */
    const int32_t *run_tid;      /* [n_tid_runs] */
    const uint64_t *run_start;   /* [n_tid_runs] index of the first record of each run; run_start[0] == 0 */
} ipx_batch_compact;

/* Junction table in key order (chromosome name as std::string::operator<, intron_start,
 * intron_end) = iteration order of the reference's JunctionTable (junctions.hh:77-100,225-246).
 * One row per Junction (junctions.hh:103-202). */
typedef struct ipx_table {
    uint64_t n_junctions;
    const int32_t *chrom;          /* [n] target index */
    const uint32_t *intron_start;  /* [n] */
    const uint32_t *intron_end;    /* [n] */
    const uint32_t *anchor_start;  /* [n] min over counted reads */
    const uint32_t *anchor_end;    /* [n] max over counted reads */
    const char *strand;            /* [n] strand of the read that created the junction, after motif correction */
    const char *splice_sites;      /* [n][8] NUL-terminated "GT/AG"-style motif, "" when no genome */
    const uint64_t *read_counts;   /* [n][3] IPX_CAT_* */
    const uint32_t *ijunc;         /* [n] creation number: the N of "sjN" */
    const float *confidence;       /* [n] Shannon entropy of read offsets (junctions.cc:54-70), host float32 */
    const uint64_t *hist_off;      /* [n + 1] offsets into hist_cnt */
    const uint32_t *hist_cnt;      /* [n_hist] reads per distinct uint16 read offset, ascending offset */
    uint64_t n_hist;
    uint64_t first_error_seq;      /* stream index of the record that raised status, if any */
    int32_t status;                /* IPX_OK or the deferred device-side error (IPX_ERR_CIGAR / _GENOME) */
    int32_t error_tid;             /* target of that record */
    char error_text[256];          /* what() of the exception the reference would have thrown */
    const int32_t *missing_targets; /* [n_missing_targets] targets with no genome sequence that were skipped
                                       under -S, in first-miss stream order (junctions_extractor.cc:411-417) */
    uint64_t n_missing_targets;
    const uint64_t *creation_seq;  /* [n] stream position of the read that created the junction (record index << 20 |
                                      ordinal of the junction within the read): sjN numbering is the rank of this value,
                                      which lets per-GPU tables be renumbered globally.  NULL from ipx_merge_rows_fold. */
    /* Merge only (NULL from extraction, whose motifs are five characters): the splice_sites strings of the input at
     * their full length -- row i has site_strings[site_ids[i]]; splice_sites[i] holds its first seven characters. */
    const char *const *site_strings;
    const uint32_t *site_ids;      /* [n] */
} ipx_table;

/* Device-time breakdown (milliseconds, CUDA events on the library's stream) and unit counts
 * accumulated since ipx_create / ipx_reset_timings; filled when params.profile != 0. */
typedef struct ipx_timings {
    double extract_ms;   /* K1: CIGAR walk -> junction events */
    double aggregate_ms; /* K2: keyed aggregation of events */
    double motif_ms;     /* K3: donor/acceptor gather + strand correction + filter */
    double finalize_ms;  /* compaction, key-order sort, sjN ranking, histogram ordering */
    double merge_ms;     /* K4: merge fold */
    uint64_t extract_launches, aggregate_launches, motif_launches, finalize_launches, merge_launches;
    uint64_t records;    /* N  records pushed */
    uint64_t cigar_live; /* C_live  CIGAR words of records that pass the A1 skip test */
    uint64_t events;     /* E  junction events that passed junction_qc */
    uint64_t junctions;  /* J  junctions in flushed tables */
    uint64_t hist_runs;  /* R  histogram runs in flushed tables */
    uint64_t motif_keys; /* U  keys evaluated against the genome */
    uint64_t merge_rows; /* rows folded by ipx_merge_rows */
    double export_ms;    /* multi-GPU exchange: rows of junctions other ranks own stored into their mailboxes */
    double import_ms;    /* ... wait for every source's export (device-side) + fold of what arrived */
    uint64_t exchange_launches;
} ipx_timings;

/* Text produced by the formatters; free with ipx_text_free. */
typedef struct ipx_text {
    char *data;
    size_t len;
} ipx_text;

/* One parsed row per TSV line of an intron-calls file (read_junction, intronProspectorMerge.cc:164-177),
 * structure-of-arrays in (file, row) order, after the --sj-filter row filter (:195). */
typedef struct ipx_merge_rows {
    uint64_t n_rows;
    const int32_t *chrom;         /* index into the names given to ipx_set_targets() */
    const uint32_t *intron_start;
    const uint32_t *intron_end;
    const uint32_t *anchor_start; /* intron_start - max_left_overhang (uint32 arithmetic) */
    const uint32_t *anchor_end;   /* intron_end + max_right_overhang */
    const char *strand;           /* first character of the strand column */
    const uint32_t *sites;        /* index into `site_strings` */
    const uint32_t *counts;       /* [n][3] */
    const float *confidence;
    const char *const *site_strings; /* distinct splice_sites strings, index 0 = "" */
    uint32_t n_site_strings;
} ipx_merge_rows;

const char *ipx_version(void);
int ipx_abi_version(void);
/* number of usable CUDA devices (0 when there is none); never fails */
int ipx_device_count(void);

/* JunctionsExtractor::JunctionsExtractor (junctions_extractor.hh:176-207) */
int ipx_create(const ipx_params *params, ipx_handle **out);
/* JunctionsExtractor::~JunctionsExtractor (junctions_extractor.cc:375-380) */
void ipx_destroy(ipx_handle *h);
const char *ipx_last_error(const ipx_handle *h);

/* save_targets (junctions_extractor.cc:564-568): target index -> chromosome name.  May be called
 * again with a longer list (multi-file mode appends); existing indexes must keep their names. */
int ipx_set_targets(ipx_handle *h, int32_t n_targets, const char *const *names);

/* Genome::Genome + has_seq + fetch (genome.cc:44-74): upload the sequences once; they stay resident
 * in HBM as raw FASTA characters so that any character (N, IUPAC, soft-masked) reproduces.
 * seqs[i] holds lengths[i] characters, newlines already removed. */
int ipx_set_genome(ipx_handle *h, int32_t n_seqs, const char *const *names, const uint64_t *lengths,
                   const char *const *seqs);

/* process_alignment for every record of the batch (junctions_extractor.cc:533-549), in stream order.
 * Host pointers; the copy to the device is part of the call and is stream-ordered. */
int ipx_push_batch(ipx_handle *h, const ipx_batch *batch);
/* Same, with every pointer in `batch` already resident in device memory (no copy). */
int ipx_push_batch_device(ipx_handle *h, const ipx_batch *batch);
/* Same as ipx_push_batch from the compact host form. */
int ipx_push_batch_compact(ipx_handle *h, const ipx_batch_compact *batch);

/* The host -> device copies of ipx_push_batch / ipx_push_batch_compact are asynchronous when the batch lies in pinned
 * memory (ipx_host_alloc): the call returns once they are enqueued and the buffers must stay untouched until they have
 * run.  ipx_sync_uploads waits until at most `keep_in_flight` of the pushes made so far (0 ... 3) still have copies
 * pending: a caller that packs into two buffers in turn calls it with 1 before it refills a buffer. */
void *ipx_host_alloc(size_t bytes);
void ipx_host_free(void *p);
int ipx_sync_uploads(ipx_handle *h, uint32_t keep_in_flight);

/* Finish aggregation on the device (key-order sort, sjN ranking, histogram ordering), asynchronously. */
int ipx_finalize_device(ipx_handle *h);
/* get_junctions() + clear() (junctions_extractor.hh:222-256): copy the finished table to the host,
 * compute float32 confidences with the host libm, empty the device table.  Calls
 * ipx_finalize_device() first if needed.  sjN numbering continues across flushes (ijunc_). */
int ipx_flush(ipx_handle *h, ipx_table **out);
void ipx_table_free(ipx_table *t);
/* Replace the sjN numbers of a table (multi-GPU: ranks of creation_seq over all GPUs' tables); ijunc has n_junctions
 * entries. */
int ipx_table_set_ijunc(ipx_table *t, const uint32_t *ijunc);
/* clear() alone (junctions_extractor.hh:222-224): drop everything accumulated so far, keep sjN numbering */
int ipx_clear(ipx_handle *h);

/* Merge fold: process_junction + Junction::merge over all rows (intronProspectorMerge.cc:179-199,
 * junctions.cc:76-105), then sort_by_anchors + renumber_junctions (:201-205, :221-223).
 * Strand-conflict lines the reference prints to stderr are returned in `conflict_log`. */
int ipx_merge_rows_fold(ipx_handle *h, const ipx_merge_rows *rows, ipx_table **out, ipx_text *conflict_log);
/* The same for one key range of a merge that is sharded over several devices (config 5: one handle per device, every
 * handle gets the rows of a contiguous range of chromosomes -- a key lives in one range, so no exchange is needed).
 * `row_ids[i]` is the position of row i in the whole input (file, row order; NULL: i).  What depends on that order
 * across ranges comes back with row ids, so that the caller can restore it: `conflict_rows[k]` is the row the k-th line
 * of `conflict_log` was raised at (malloc'ed: ipx_free), and on IPX_ERR_MERGE `*fail_row` is the row of the fatal
 * disagreement (the reference stops at the first one in input order, after the conflict lines of the rows before it). */
int ipx_merge_rows_fold_part(ipx_handle *h, const ipx_merge_rows *rows, const uint64_t *row_ids, ipx_table **out,
                             ipx_text *conflict_log, uint64_t **conflict_rows, uint64_t *n_conflict_rows, uint64_t *fail_row);
void ipx_free(void *p);

/* Intron-call TSV rows tokenised on the device (Tsv::next_row + read_junction,
 * tsv.hh:27-74, intronProspectorMerge.cc:164-177).  `bytes` = one uncompressed file WITHOUT its header line (any host
 * memory; ipx_host_alloc'ed memory crosses the link faster), `columns` = where the header put the columns read_junction
 * asks for.  One ipx_tsv_rows entry per line, in file order.  The device parses the plain form only (unsigned decimal
 * integers that fit 32 bits, decimal fractions without sign or exponent and at most 8 fractional digits, the right number
 * of columns); every other row has flags[i] != 0 and none of its values set: the caller parses the line
 * bytes[line_off[i] .. + line_len[i]) with the reference's rules, which also produces the reference's error for it.
 * IPX_ERR_ARG: an input this entry point does not take (4 GB or more, a column missing). */
typedef struct ipx_tsv_columns {
    int32_t n_columns; /* of the header: rows with another number of columns are flagged */
    int32_t chrom, intron_start, intron_end, strand, uniq_mapped_count, multi_mapped_count, unsure_mapped_count, max_left_overhang,
        max_right_overhang, confidence, splice_sites;
} ipx_tsv_columns;
typedef struct ipx_tsv_rows {
    uint64_t n_rows;
    const uint32_t *intron_start, *intron_end;
    const uint32_t *anchor_start, *anchor_end; /* start - max_left_overhang, end + max_right_overhang (uint32 arithmetic) */
    const uint32_t *counts;                    /* [n][3] uniq, multi, unsure */
    const float *confidence;
    const char *strand;                        /* first character of the column, 0 when it is empty */
    const uint64_t *sites8;                    /* the splice_sites string, byte k in bits 8k .. 8k+7 */
    const uint8_t *sites_len;                  /* <= 8 (longer strings are flagged) */
    const uint32_t *chrom_off;                 /* the chrom string is bytes[chrom_off .. + chrom_len) */
    const uint16_t *chrom_len;
    const uint32_t *line_off, *line_len;       /* the line without its newline */
    const uint8_t *flags;
} ipx_tsv_rows;
int ipx_tsv_tokenize(int32_t device, const char *bytes, uint64_t n_bytes, const ipx_tsv_columns *columns, ipx_tsv_rows **out);
void ipx_tsv_rows_free(ipx_tsv_rows *rows);

/* Multi-GPU exchange (one process per GPU).  After ipx_finalize_device():
 *  ipx_partial_sizes   rows per destination rank for the given key splitters (n_parts - 1 of them,
 *                      each = first key owned by the next rank: target index, intron_start)
 *  ipx_partial_export  packed junction / histogram rows, destination-major, into caller-provided
 *                      DEVICE buffers (ipx_partial_row_bytes gives the row sizes)
 *  ipx_partial_import  fold the rows received from ONE source rank (DEVICE buffers: offset rows name their junction
 *                      by its index among the junction rows of the same source) into this rank's emptied table;
 *                      call once per source, own share included, then ipx_flush. */
int ipx_partial_row_bytes(uint32_t *junction_row_bytes, uint32_t *hist_row_bytes);
int ipx_partial_sizes(ipx_handle *h, int32_t n_parts, const int32_t *split_tid, const uint32_t *split_start,
                      uint64_t *junction_rows, uint64_t *hist_rows);
int ipx_partial_export(ipx_handle *h, void *dev_junction_rows, void *dev_hist_rows);
int ipx_partial_import(ipx_handle *h, const void *dev_junction_rows, uint64_t n_junction_rows,
                       const void *dev_hist_rows, uint64_t n_hist_rows);

/* The same exchange over peer memory (NVLink / NVSwitch), without send buffers, size exchange or data-path collective:
 * every rank owns a mailbox with one slot per source rank, and the exporting kernels store each row straight into its
 * owner's slot.  Set-up, once:
 *  ipx_mailbox_create        allocate this rank's mailbox: `world` sources x 2 (consecutive exchanges alternate),
 *                            each with room for the given numbers of junction and offset rows
 *  ipx_mailbox_ipc_handle    64 bytes (cudaIpcMemHandle_t) for the other processes;  ipx_mailbox_ptr: the device
 *                            pointer itself, for handles that live in the same process
 *  ipx_mailbox_connect       the mailboxes of all ranks, in rank order: ipc != 0: `peers` = world x 64-byte IPC
 *                            handles (own entry ignored); ipc == 0: `peers` = world device pointers
 * Per exchange, after the local batches are pushed (and ipx_count_deferred with the canon filter):
 *  ipx_partial_export_mail   enqueue the export into the owners' slots (asynchronous, no host round trip) and empty
 *                            the local table
 *  -- the caller makes every rank's export complete before any import starts: any collective or barrier ordered
 *     after ipx_stream(h) on every rank and before it on the importing side (a one-element all-reduce will do) --
 *  ipx_partial_import_mail   fold all the slots of this rank's mailbox (row counts come from the slots' headers);
 *                            then ipx_flush.  A slot that was too small is reported by ipx_flush as IPX_ERR_CAPACITY. */
int ipx_mailbox_create(ipx_handle *h, int32_t world, int32_t rank, uint64_t junction_rows_per_source,
                       uint64_t hist_rows_per_source);
int ipx_mailbox_ipc_handle(ipx_handle *h, void *handle_64_bytes);
void *ipx_mailbox_ptr(ipx_handle *h);
int ipx_mailbox_connect(ipx_handle *h, int32_t ipc, const void *peers);
int ipx_partial_export_mail(ipx_handle *h, const int32_t *split_tid, const uint32_t *split_start);
int ipx_partial_import_mail(ipx_handle *h);
/* The export is enqueued without a host round trip, i.e. before the host knows whether the fold of the last batch was
 * complete (a junction table that fills up is grown and the fold repeated, lazily).  ipx_exchange_settle (every rank,
 * before ipx_flush) waits for this rank's work and reports in *repeat_exchange whether events were folded after its
 * export: if any rank reports 1, every rank repeats ipx_partial_export_mail / ipx_partial_import_mail (cheap: only
 * what is still held for other ranks moves).  ipx_flush refuses to return a table that skipped this. */
int ipx_exchange_settle(ipx_handle *h, int32_t *repeat_exchange);
/* Every stride-th junction key of the local table (target index, intron start), at most max_keys, into host arrays:
 * a sample for load-balanced key-range splitters.  Synchronises. */
int ipx_sample_keys(ipx_handle *h, uint64_t stride, uint64_t max_keys, int32_t *tid, uint32_t *intron_start, uint64_t *n_keys);

/* Multi-GPU with the canon filter.  Which event creates a junction (add_junction, junctions_extractor.cc:488-506) depends
 * on every earlier read of the whole stream, so with ipx_params.deferred_counting the ranks first agree on it:
 *  ipx_thresholds_rows    settle this rank's local creation thresholds; number and size of the rows to exchange
 *  ipx_thresholds_export  one row per junction key seen here (key + threshold) into a caller-provided DEVICE buffer;
 *                         returns when the rows are complete
 *  ipx_thresholds_import  rows of any rank (DEVICE buffer): a key known here takes the earlier threshold
 *  ipx_count_deferred     count the kept junction events against the agreed thresholds; afterwards the
 *                         ipx_partial_* exchange and ipx_flush work as without the filter */
int ipx_thresholds_rows(ipx_handle *h, uint64_t *n_rows, uint32_t *row_bytes);
int ipx_thresholds_export(ipx_handle *h, void *dev_rows);
int ipx_thresholds_import(ipx_handle *h, const void *dev_rows, uint64_t n_rows);
int ipx_count_deferred(ipx_handle *h);

/* Formatters (junctions.cc:162-268).  Rows [begin, end) of the table are taken in key order, sorted as
 * the reference sorts them (sort_by_anchors for IPX_FMT_JUNCTION_BED, sort_by_introns otherwise,
 * junctions.hh:205-222) and rows with confidence < min_confidence are suppressed. */
#define IPX_FMT_JUNCTION_BED 0 /* print_anchor_bed, BED12 */
#define IPX_FMT_INTRON_BED9 1  /* print_intron_bed(9) */
#define IPX_FMT_INTRON_BED6 2  /* print_intron_bed(6) */
#define IPX_FMT_INTRON_TSV 3   /* print_intron_call_tsv rows (no header) */
int ipx_format(const ipx_table *t, const char *const *target_names, int format, float min_confidence,
               uint64_t begin, uint64_t end, ipx_text *out);
/* print_junction_call_header (junctions.cc:247-252) */
int ipx_format_tsv_header(ipx_text *out);
void ipx_text_free(ipx_text *t);

int ipx_get_timings(ipx_handle *h, ipx_timings *out);
int ipx_reset_timings(ipx_handle *h);
/* cudaStream_t the library launches on (for callers that time with their own events) */
void *ipx_stream(ipx_handle *h);

#ifdef __cplusplus
}
#endif
#endif
