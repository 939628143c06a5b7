/* ipx_merge.cu -- K4: the intronProspectorMerge fold on the device (sm_100a).
 *
 * Reference: process_junction + Junction::merge (intronProspectorMerge.cc:179-188, junctions.cc:76-105).
 * The reference folds rows of one intron in (file, row) order.  The fold is order-free whenever the rows of
 * an intron (a) carry at most one distinct strand other than '.', and (b) all carry the same splice-site
 * string: then no strand-conflict skip and no splice-site exception can occur, the result strand is that
 * strand (or '.'), the motif is the common string, and counts / anchors / confidence are sum / min / max /
 * max.  The kernel computes exactly that, plus the evidence for (a) and (b); introns that violate either are
 * flagged, their rows are picked out of the input by a second pass, and the host folds those few rows in row order
 * (ipx_api.cu), which also produces the reference's stderr lines and fatal error.
 *
 * The table is the extraction path's (ipx_table.cuh): hash slots -> junction ids, everything an intron owns indexed by
 * id, ids handed out from per-block pools.  It is sized by the number of introns, not of rows (64 tables over the same
 * genes fold into a few hundred thousand introns: tens of MB, resident in L2), and grown when it fills.  Key = intron
 * start << 32 | intron end (merge rows are not bounded by --max-intron-length) with the chromosome index compared
 * beside it -- the "wide" layout.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipx_merge.h"

namespace ipx {

__global__ void __launch_bounds__(256) k_merge_init_rows(Tables t, MStat *stat, uint64_t first, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        MStat s;
        s.cnt[0] = s.cnt[1] = s.cnt[2] = 0ull;
        s.a0 = 0xFFFFFFFFu;
        s.a1 = 0u;
        s.conf_bits = 0u;
        s.ev_lo = 0xFFFFFFFFu;
        s.ev_hi = 0u;
        s.odd = 0u;
        s.pad[0] = s.pad[1] = s.pad[2] = s.pad[3] = 0u;
        stat[j] = s;
        t.jslot[j] = SLOT_NONE;
    }
}

__global__ void k_merge_reset(Tables t, unsigned long long *n_unfolded) {
    t.id_pools[threadIdx.x] = (unsigned long long)ID_BLOCK;
    if (threadIdx.x != 0) return;
    *t.n_keys = 0ull;
    *t.overflow = 0u;
    *n_unfolded = 0ull;
}

__global__ void k_merge_init_slots(Slot *slots, uint64_t cap) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
        uint4 *p = reinterpret_cast<uint4 *>(&slots[i]);
        p[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, JID_UNSET, BASE_UNSET);
        p[1] = make_uint4(BASE_UNSET, BASE_UNSET, 0u, 0u);
    }
}

/* strand and splice-site evidence of a row as one word each way: site << 9 | strand code.  The code of a strand character
 * is the character + 1 (any byte, 0 included: an empty strand column); '.' counts as the largest code for the minimum and
 * as 0 for the maximum, so that it never decides either */
__device__ __forceinline__ uint32_t ev_for_min(uint32_t site, uint32_t strand) { return (site << 9) | (strand == (uint32_t)'.' ? 0x1FFu : strand + 1u); }
__device__ __forceinline__ uint32_t ev_for_max(uint32_t site, uint32_t strand) { return (site << 9) | (strand == (uint32_t)'.' ? 0u : strand + 1u); }

/* Thread per row: one probe (a load of the intron's slot, out of L2 once the table is warm), then fire-and-forget atomics
 * into its 64-byte record.  Rows of one warp that name the same intron (the last rows of one file and the first of the
 * next) are grouped with match.any and probed once by the group's first lane. */
__global__ void __launch_bounds__(256) k_merge_fold(MergeRowsView r, Tables t, MStat *stat, unsigned long long *n_unfolded) {
    __shared__ IdPool s_pool;
    if (threadIdx.x == 0) id_pool_load(t, &s_pool);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    uint32_t spare = JID_NONE; /* see junction_probe */
    /* whole warps run the loop together: the bound is the warp's first row */
    for (uint64_t i0w = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0w < r.n; i0w += stride) {
        const uint64_t i = i0w + lane;
        const bool valid = i < r.n;
        uint32_t i0 = 0u, i1 = 0u;
        int32_t chrom = -1 - lane; /* idle lanes group with nobody */
        if (valid) {
            i0 = r.i0[i];
            i1 = r.i1[i];
            chrom = r.chrom[i];
        }
        const unsigned long long key = ((unsigned long long)i0 << 32) | i1; /* never ~0: start < end */
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, key) & __match_any_sync(0xFFFFFFFFu, chrom);
        const int leader = __ffs(grp) - 1;
        uint32_t jid = JID_NONE;
        if (valid && lane == leader) jid = junction_probe<true>(t, &s_pool, key, chrom, spare).jid;
        __syncwarp();
        jid = __shfl_sync(0xFFFFFFFFu, jid, leader);
        if (!valid) continue;
        if (jid == JID_NONE) {
            atomicAdd(n_unfolded, 1ull); /* the table is full: the host grows it and folds again */
            continue;
        }
        MStat *s = &stat[jid];
        const uint32_t c0 = r.counts[3 * i + 0], c1 = r.counts[3 * i + 1], c2 = r.counts[3 * i + 2];
        if (c0) atomicAdd(&s->cnt[0], (unsigned long long)c0);
        if (c1) atomicAdd(&s->cnt[1], (unsigned long long)c1);
        if (c2) atomicAdd(&s->cnt[2], (unsigned long long)c2);
        atomicMin(&s->a0, r.a0[i]);
        atomicMax(&s->a1, r.a1[i]);
        const float c = r.conf[i];
        /* max(j1.confidence, confidence) on the bits: non-negative floats order like unsigned integers; a value with the
         * sign bit set (negative, -0.0) or a NaN does not, and sends the intron to the host fold */
        const uint32_t cb = __float_as_uint(c);
        if (!(cb >> 31) && c == c) atomicMax(&s->conf_bits, cb); else atomicOr(&s->odd, 1u);
        const uint32_t st = (uint8_t)r.strand[i], site = r.sites[i];
        atomicMin(&s->ev_lo, ev_for_min(site, st));
        atomicMax(&s->ev_hi, ev_for_max(site, st));
    }
    __syncthreads();
    if (threadIdx.x == 0) id_pool_store(t, &s_pool);
}

/* sort keys per intron id: (chromosome-name rank << 32 | start, end); ids no intron took are left out.
 * d_count (zeroed by the caller): [0] introns, [3] |= 1 when B was too small, [4] ids looked at */
__global__ void k_merge_keys(Tables t, const uint32_t *name_rank, uint64_t B, unsigned long long *d_count, unsigned long long *sort_key,
                             uint32_t *sort_sub) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) {
        const unsigned long long nk = junctions_in_use(t);
        d_count[4] = nk < B ? nk : B;
        if (nk > B) atomicOr(&d_count[3], 1ull);
    }
    unsigned long long sk = EMPTY64;
    uint32_t sub = 0u;
    const bool take = j < B && j < junctions_in_use(t) && t.jslot[j] != SLOT_NONE;
    if (take) {
        const unsigned long long key = t.jkey[j];
        sk = ((unsigned long long)name_rank[t.jtid[j]] << 32) | (key >> 32);
        sub = (uint32_t)key;
    }
    if (j < B) {
        sort_key[j] = sk;
        sort_sub[j] = sub;
    }
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, take);
    if (mask && (threadIdx.x & 31) == 0) atomicAdd(&d_count[0], (unsigned long long)__popc(mask));
}

__global__ void k_merge_gather(Tables t, const MStat *stat, const uint32_t *order, uint64_t n, MergeOut out, uint8_t *slow_of_id) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = order[i];
    const unsigned long long key = t.jkey[j];
    const MStat s = stat[j];
    out.chrom[i] = t.jtid[j];
    out.i0[i] = (uint32_t)(key >> 32);
    out.i1[i] = (uint32_t)key;
    out.a0[i] = s.a0;
    out.a1[i] = s.a1;
    out.counts[3 * i + 0] = s.cnt[0];
    out.counts[3 * i + 1] = s.cnt[1];
    out.counts[3 * i + 2] = s.cnt[2];
    out.conf[i] = __uint_as_float(s.conf_bits);
    const bool one_site = (s.ev_lo >> 9) == (s.ev_hi >> 9);
    const uint32_t st_min = s.ev_lo & 0x1FFu, st_max = s.ev_hi & 0x1FFu; /* 0x1FF / 0 when every row has '.' */
    const bool one_strand = st_max == 0u || st_min == st_max;
    out.strand[i] = st_max == 0u ? '.' : (char)(st_min - 1u);
    out.site[i] = s.ev_lo >> 9;
    const uint8_t slow = (one_strand && one_site && !s.odd) ? 0 : 1;
    out.slow[i] = slow;
    slow_of_id[j] = slow;
}

/* the rows of the introns that were flagged, for the host's fold in row order */
__global__ void __launch_bounds__(256) k_merge_slow_rows(MergeRowsView r, Tables t, const uint8_t *slow_of_id, unsigned long long *n_slow,
                                                         uint64_t cap, unsigned long long *slow_rows) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r.n; i += stride) {
        const unsigned long long key = ((unsigned long long)r.i0[i] << 32) | r.i1[i];
        const Found f = junction_find(t, key, r.chrom[i]);
        if (f.jid == JID_NONE || !slow_of_id[f.jid]) continue;
        const unsigned long long at = atomicAdd(n_slow, 1ull);
        if (at < cap) slow_rows[at] = r.row_base + i;
    }
}

void launch_merge_init(const Tables &t, MStat *stat, unsigned long long *n_unfolded, cudaStream_t st) {
    k_merge_init_slots<<<148 * 4, 256, 0, st>>>(t.slots, t.cap_mask + 1);
    k_merge_init_rows<<<148 * 4, 256, 0, st>>>(t, stat, 0, t.max_j);
    k_merge_reset<<<1, N_ID_POOLS, 0, st>>>(t, n_unfolded);
}

void launch_merge_fold(const MergeRowsView &r, const Tables &t, MStat *stat, unsigned long long *n_unfolded, int sm_count, cudaStream_t st) {
    if (r.n == 0) return;
    uint64_t want = (r.n + 255) / 256, maxb = (uint64_t)sm_count * 4;
    if (maxb > N_ID_POOLS) maxb = N_ID_POOLS;
    k_merge_fold<<<(unsigned)(want < maxb ? want : maxb), 256, 0, st>>>(r, t, stat, n_unfolded);
}

void launch_merge_keys(const Tables &t, const uint32_t *name_rank, uint64_t B, unsigned long long *d_count, unsigned long long *sort_key,
                       uint32_t *sort_sub, cudaStream_t st) {
    k_merge_keys<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(t, name_rank, B, d_count, sort_key, sort_sub);
}

void launch_merge_gather(const Tables &t, const MStat *stat, const uint32_t *order, uint64_t n, const MergeOut &out, uint8_t *slow_of_id,
                         cudaStream_t st) {
    if (n == 0) return;
    k_merge_gather<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t, stat, order, n, out, slow_of_id);
}

void launch_merge_slow_rows(const MergeRowsView &r, const Tables &t, const uint8_t *slow_of_id, unsigned long long *n_slow, uint64_t cap,
                            unsigned long long *slow_rows, int sm_count, cudaStream_t st) {
    if (r.n == 0) return;
    uint64_t want = (r.n + 255) / 256, maxb = (uint64_t)sm_count * 8;
    k_merge_slow_rows<<<(unsigned)(want < maxb ? want : maxb), 256, 0, st>>>(r, t, slow_of_id, n_slow, cap, slow_rows);
}

}  // namespace ipx
