/* ipx_motif.cuh -- donor/acceptor lookup in the HBM-resident genome and the canonical-motif test, shared by K3
 * (k_resolve_T) and the finalize row gather.  Reference: get_splice_sites / correct_strand
 * (junctions_extractor.cc:405-444), splice_juncs.cc:9-61, genome.cc:11-74. */
#ifndef IPX_MOTIF_CUH
#define IPX_MOTIF_CUH
#include "ipx_engine.h"

namespace ipx {

__device__ __forceinline__ char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
__device__ __forceinline__ char low(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }
__device__ __forceinline__ char comp(char c) {
    /* dna_reverse_complement's table (genome.cc:14-28): only ACGT / acgt are mapped */
    switch (c) {
        case 'A': return 'T';
        case 'T': return 'A';
        case 'G': return 'C';
        case 'C': return 'G';
        case 'a': return 't';
        case 't': return 'a';
        case 'g': return 'c';
        case 'c': return 'g';
        default: return c;
    }
}
/* is_canonical (splice_juncs.cc:25-38) for an upper-cased donor pair d0 d1 and acceptor pair a0 a1 */
__device__ __forceinline__ bool canon4(char d0, char d1, char a0, char a1) {
    return (d0 == 'G' && d1 == 'T' && a0 == 'A' && a1 == 'G') || (d0 == 'G' && d1 == 'C' && a0 == 'A' && a1 == 'G') ||
           (d0 == 'A' && d1 == 'T' && a0 == 'A' && a1 == 'C');
}

enum : int { MOTIF_OK = 0, MOTIF_NONE = 1, MOTIF_ERR_MISSING = 2, MOTIF_ERR_LEN_DONOR = 3, MOTIF_ERR_LEN_ACCEPTOR = 4 };

/* Genome::fetch clipping (genome.cc:62-74 over faidx_fetch_seq): number of characters a fetch of [s, e) returns */
__device__ __forceinline__ int64_t fetch_len(int64_t len, int64_t s, int64_t e) {
    int64_t b = s, en = e - 1;
    if (en < b - 1) en = b - 1;
    if (b < 0) b = 0;
    if (b >= len) b = len;
    if (en < 0) en = 0; else if (en >= len) en = len - 1;
    int64_t got = en - b + 1;
    return got < 0 ? 0 : got;
}

/* get_splice_sites' fetches (junctions_extractor.cc:405-420): the four upper-cased characters, forward strand */
__device__ __forceinline__ int motif_fetch(const GenomeView &g, int32_t tid, uint32_t i0, uint32_t i1, char f[4]) {
    if (!g.have_genome) return MOTIF_NONE;
    const int32_t gi = (tid >= 0 && tid < g.n_targets) ? g.target_seq[tid] : -1;
    if (gi < 0) return g.skip_missing ? MOTIF_NONE : MOTIF_ERR_MISSING;
    const int64_t len = (int64_t)g.len[gi];
    if (fetch_len(len, i0, (int64_t)i0 + 2) != 2) return MOTIF_ERR_LEN_DONOR;
    if (fetch_len(len, (int64_t)i1 - 2, i1) != 2) return MOTIF_ERR_LEN_ACCEPTOR;
    const char *s = g.seq[gi];
    f[0] = up(s[i0]);
    f[1] = up(s[i0 + 1]);
    f[2] = up(s[i1 - 2]);
    f[3] = up(s[i1 - 1]);
    return MOTIF_OK;
}

}  // namespace ipx
#endif
