/* faidx.h -- minimal stand-in for <htslib/faidx.h>; see hts.h in this directory.
 * Semantics follow htslib: fai_load3() needs an existing "<fn>.fai" (it is not
 * built on the fly when flags == 0), faidx_fetch_seq() takes a 0-based inclusive
 * end, clips to the sequence and returns a malloc'd NUL-terminated string with
 * the file's own letter case (reference use: genome.cc:46-64).
 */
#ifndef IPX_COMPAT_FAIDX_H
#define IPX_COMPAT_FAIDX_H
#include <stdint.h>
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct faidx_t faidx_t;
#define FAI_CREATE 0x01
faidx_t *fai_load(const char *fn);
faidx_t *fai_load3(const char *fn, const char *fnfai, const char *fngzi, int flags);
void fai_destroy(faidx_t *fai);
int fai_build(const char *fn);
int faidx_has_seq(const faidx_t *fai, const char *seq);
int faidx_nseq(const faidx_t *fai);
const char *faidx_iseq(const faidx_t *fai, int i);
int faidx_seq_len(const faidx_t *fai, const char *seq);
char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len);
/* not htslib: 1 when the handle holds a bgzip-compressed FASTA inflated in memory (fetches need no file position then,
 * so threads may share the handle) */
int ipx_fai_in_memory(const faidx_t *fai);
#ifdef __cplusplus
}
#endif
#endif
