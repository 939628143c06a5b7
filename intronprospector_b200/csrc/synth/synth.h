/* synth.h -- seeded synthetic alignment generator for the benchmark configs of BASELINE.json / SURVEY.md 8(d).
 * Host only.  Produces coordinate-sorted record columns in the ipx_batch layout and, optionally, the same
 * records as an uncompressed BAM stream for the CPU reference to read. */
#ifndef IPX_SYNTH_H
#define IPX_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define IPX_SYNTH_SHORT_UNSTRANDED 0 /* config 2: 2x150 short reads, XS:A on spliced reads, NH tags */
#define IPX_SYNTH_SHORT_STRANDED 1   /* config 3: same reads, flags follow an RF library, planted motifs in the genome */
#define IPX_SYNTH_LONG 2             /* config 4: ~10 kb reads, one op per ~12 bases, 5-15 N, clips, 5 % supplementary */

typedef struct ipx_synth_spec {
    uint64_t n_records;
    uint64_t seed;        /* read sampling seed */
    uint64_t model_seed;  /* gene model seed (shared by all shards of one workload) */
    int32_t kind;         /* IPX_SYNTH_* */
    int32_t n_contigs;    /* 25 */
    uint64_t genome_len;  /* total, split evenly */
    int32_t n_genes;      /* 20000 */
    int32_t threads;      /* 0 = hardware concurrency */
} ipx_synth_spec;

typedef struct ipx_synth_data ipx_synth_data;

/* generate; returns NULL on failure */
ipx_synth_data *ipx_synth_generate(const ipx_synth_spec *spec);
void ipx_synth_free(ipx_synth_data *d);
uint64_t ipx_synth_n_records(const ipx_synth_data *d);
uint64_t ipx_synth_n_cigar(const ipx_synth_data *d);
/* copy the columns into caller buffers (e.g. pinned memory); any pointer may be NULL */
void ipx_synth_copy(const ipx_synth_data *d, int32_t *tid, int32_t *pos, uint16_t *flag, uint8_t *nh_class, uint8_t *xs,
                    uint32_t *cigar_off, uint32_t *cigar);
int32_t ipx_synth_n_contigs(const ipx_synth_data *d);
const char *ipx_synth_contig_name(const ipx_synth_data *d, int32_t i);
uint64_t ipx_synth_contig_len(const ipx_synth_data *d, int32_t i);
/* write records [begin, end) as a BAM stream: raw != 0 -> uncompressed, no BGZF framing (hts_compat reads it) */
int ipx_synth_write_bam(const ipx_synth_data *d, const char *path, uint64_t begin, uint64_t end, int raw);
/* the same records as an aligner writes them (read names, bases, qualities, mate fields, STAR-style tags incl. B arrays),
 * BGZF-compressed at `level` by `threads` threads (0 = all): the input of the command-line tools */
int ipx_synth_write_bam_full(const ipx_synth_data *d, const char *path, uint64_t begin, uint64_t end, int threads, int level);
/* synthetic genome consistent with the gene model (planted splice motifs); writes FASTA + .fai */
int ipx_synth_write_genome(const ipx_synth_data *d, const char *fasta_path);

#ifdef __cplusplus
}
#endif
#endif
