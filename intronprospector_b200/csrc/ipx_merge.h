/* ipx_merge.h -- device structures of the merge fold (ipx_merge.cu). */
#ifndef IPX_MERGE_H
#define IPX_MERGE_H
#include <stdint.h>
#include <cuda_runtime.h>
#include "ipx_table.cuh"

namespace ipx {

#if defined(__CUDACC__)
struct __align__(64) MStat {
#else
struct alignas(64) MStat {
#endif
    unsigned long long cnt[3];
    unsigned long long first_row; /* first row (file, row order) of this intron */
    uint32_t a0, a1;
    uint32_t conf_bits;           /* max confidence as float bits (non-negative floats order like uints) */
    uint32_t strand_min, strand_max; /* over strands other than '.'; max == 0 means none */
    uint32_t site_min, site_max;  /* over splice-site string ids */
    uint32_t odd;                 /* a row with negative / NaN confidence: fold on the host */
};

struct MergeTables {
    unsigned long long *keys; /* [cap] intron_start << 32 | intron_end */
    uint32_t *chrom;          /* [cap] chromosome index, 0xFFFFFFFF = unclaimed */
    MStat *stat;              /* [cap] */
    uint64_t cap_mask;
    unsigned long long *n_keys;
    uint32_t *overflow;
};

struct MergeRowsView {
    const int32_t *chrom;
    const uint32_t *i0, *i1, *a0, *a1;
    const char *strand;
    const uint32_t *sites;
    const uint32_t *counts; /* [n][3] */
    const float *conf;
    uint64_t n;
    uint64_t row_base;
};

struct MergeOut {
    int32_t *chrom;
    uint32_t *i0, *i1, *a0, *a1, *site;
    unsigned long long *counts;
    float *conf;
    char *strand;
    uint8_t *slow;
};

void launch_merge_init(const MergeTables &t, cudaStream_t st);
void launch_merge_fold(const MergeRowsView &r, const MergeTables &t, int sm_count, cudaStream_t st);
void launch_merge_compact(const MergeTables &t, unsigned long long *d_count, unsigned long long *sort_lo, uint32_t *slot_out,
                          cudaStream_t st);
void launch_merge_rank_keys(const MergeTables &t, const uint32_t *name_rank, const uint32_t *slots, uint64_t n,
                            unsigned long long *hi_out, cudaStream_t st);
void launch_merge_gather(const MergeTables &t, const uint32_t *sorted_slot, uint64_t n, const MergeOut &out,
                         cudaStream_t st);

}  // namespace ipx
#endif
