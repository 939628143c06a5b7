/* ipx_io.h -- C ABI of the host I/O library (libipx_io.so): everything that sits on the host side of the
 * ipx.h boundary and is not arithmetic -- alignment file reading into columnar batches, FASTA loading,
 * intron-calls TSV parsing.  It replaces, for this path, what the reference gets from htslib
 * (sam_open/sam_hdr_read/sam_read1, junctions_extractor.cc:571-634; faidx, genome.cc:44-74) and from its
 * Tsv class (tsv.hh, intronProspectorMerge.cc:164-199).  No CUDA here.
 */
#ifndef IPX_IO_H
#define IPX_IO_H
#include <stddef.h>
#include <stdint.h>
#include "ipx.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipxio_reader ipxio_reader;
typedef struct ipxio_genome ipxio_genome;
typedef struct ipxio_calls ipxio_calls;

/* text of the last error raised on the calling thread by any ipxio_* function */
const char *ipxio_last_error(void);

/* JunctionsExtractor::open (junctions_extractor.cc:571-586): SAM, BAM or raw (BGZF-less) BAM */
int ipxio_open(const char *path, ipxio_reader **out);
void ipxio_close(ipxio_reader *r);
int32_t ipxio_n_targets(const ipxio_reader *r);
const char *ipxio_target_name(const ipxio_reader *r, int32_t i);
/* Decode up to max_records records (and at most max_cigar CIGAR words) into the caller's columns, in file
 * order, exactly the fields ipx_batch carries.  Returns the number of records read, 0 at end of file,
 * < 0 on a read error.  cigar_off gets n + 1 entries. */
int64_t ipxio_read_batch(ipxio_reader *r, uint64_t max_records, uint64_t max_cigar, int32_t *tid, int32_t *pos,
                         uint16_t *flag, uint8_t *nh_class, uint8_t *xs, uint32_t *cigar_off, uint32_t *cigar,
                         uint64_t *n_cigar_out);
/* The same columns for a whole BAM file through the multi-threaded BAM pipeline (host/bam_pipeline.h: what the
 * command-line tool reads BAM files with; `threads` <= 0: one per core).  Returns the number of records, -1 on a damaged
 * file, -2 when the buffers are too small.  Unlike ipxio_read_batch the NH / XS columns are only filled for reads with
 * an N operation (the only reads the path looks at them for). */
int64_t ipxio_pipeline_read(const char *path, int32_t threads, uint64_t max_records, uint64_t max_cigar, int32_t *tid, int32_t *pos,
                            uint16_t *flag, uint8_t *nh_class, uint8_t *xs, uint32_t *cigar_off, uint32_t *cigar,
                            uint64_t *n_cigar_out);

/* Genome::Genome (genome.cc:44-50): needs "<fasta>.fai"; sequences are returned whole, newlines removed */
int ipxio_genome_load(const char *fasta, ipxio_genome **out);
void ipxio_genome_free(ipxio_genome *g);
int32_t ipxio_genome_n_seqs(const ipxio_genome *g);
const char *const *ipxio_genome_names(const ipxio_genome *g);
const uint64_t *ipxio_genome_lengths(const ipxio_genome *g);
const char *const *ipxio_genome_seqs(const ipxio_genome *g);
/* samtools faidx stand-in */
int ipxio_faidx_build(const char *fasta);

/* process_calls_tsv (intronProspectorMerge.cc:190-199) for every file, in order: parse, apply the row filter
 * (IPX_FILTER_*), produce merge rows and the chromosome name list.  On a malformed file the rows parsed so far
 * are kept, the function returns IPX_ERR_ARG and ipxio_last_error() holds the reference's message. */
int ipxio_calls_load(const char *const *paths, int32_t n_paths, int32_t junction_filter, ipxio_calls **out);
/* The same with the rows of every file tokenised by `tokenize` -- ipx_tsv_tokenize of ipx.h bound to a device, or any
 * stand-in with its contract -- and only the rows it flags parsed here (this library itself needs no CUDA).  `rows_free`
 * releases what `tokenize` returned.  Rows, dictionaries and errors are those of ipxio_calls_load. */
typedef int (*ipxio_tokenize_fn)(const char *bytes, uint64_t n_bytes, const ipx_tsv_columns *columns, ipx_tsv_rows **out);
typedef void (*ipxio_rows_free_fn)(ipx_tsv_rows *rows);
int ipxio_calls_load_tokenized(const char *const *paths, int32_t n_paths, int32_t junction_filter, ipxio_tokenize_fn tokenize,
                               ipxio_rows_free_fn rows_free, ipxio_calls **out);
void ipxio_calls_free(ipxio_calls *c);
const ipx_merge_rows *ipxio_calls_rows(const ipxio_calls *c);
int32_t ipxio_calls_n_chroms(const ipxio_calls *c);
const char *const *ipxio_calls_chroms(const ipxio_calls *c);

#ifdef __cplusplus
}
#endif
#endif
