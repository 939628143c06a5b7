/* ipx_merge.h -- device structures of the merge fold (ipx_merge.cu). */
#ifndef IPX_MERGE_H
#define IPX_MERGE_H
#include <stdint.h>
#include <cuda_runtime.h>
#include "ipx_table.cuh"

namespace ipx {

/* what an intron accumulates, indexed by junction id (ipx_table.cuh) */
#if defined(__CUDACC__)
struct __align__(64) MStat {
#else
struct alignas(64) MStat {
#endif
    unsigned long long cnt[3];
    uint32_t a0, a1;
    uint32_t conf_bits; /* max confidence as float bits (non-negative floats order like uints) */
    uint32_t ev_lo;     /* min over rows of (splice-site string id << 9 | strand code, '.' as 0x1FF): see ev_for_min */
    uint32_t ev_hi;     /* max over rows of (splice-site string id << 9 | strand code, '.' as 0) */
    uint32_t odd;       /* a row with negative / -0.0 / NaN confidence: fold on the host */
    uint32_t pad[4];
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

/* `t`: slots, cap_mask, jkey, jtid, jslot, max_j, n_keys, id_pools, overflow, kl.wide = 1 (nothing else is used) */
void launch_merge_init(const Tables &t, MStat *stat, unsigned long long *n_unfolded, cudaStream_t st);
void launch_merge_fold(const MergeRowsView &r, const Tables &t, MStat *stat, unsigned long long *n_unfolded, int sm_count, cudaStream_t st);
/* d_count[4] = ids in use (<= B), d_count[3] = 1 when B was too small */
void launch_merge_keys(const Tables &t, const uint32_t *name_rank, uint64_t B, unsigned long long *d_count, unsigned long long *sort_key,
                       uint32_t *sort_sub, cudaStream_t st);
void launch_merge_gather(const Tables &t, const MStat *stat, const uint32_t *order, uint64_t n, const MergeOut &out, uint8_t *slow_of_id,
                         cudaStream_t st);
void launch_merge_slow_rows(const MergeRowsView &r, const Tables &t, const uint8_t *slow_of_id, unsigned long long *n_slow, uint64_t cap,
                            unsigned long long *slow_rows, int sm_count, cudaStream_t st);

}  // namespace ipx
#endif
