/* ipx_merge.cu -- K4: the intronProspectorMerge fold on the device.
 *
 * Reference: process_junction + Junction::merge (intronProspectorMerge.cc:179-188, junctions.cc:76-105).
 * The reference folds rows of one intron in (file, row) order.  The fold is order-free whenever the rows of
 * an intron (a) carry at most one distinct strand other than '.', and (b) all carry the same splice-site
 * string: then no strand-conflict skip and no splice-site exception can occur, the result strand is that
 * strand (or '.'), the motif is the common string, and counts / anchors / confidence are sum / min / max /
 * max.  The kernel computes exactly that, plus the evidence for (a) and (b); introns that violate either are
 * flagged and re-folded on the host in row order (merge_refold in ipx_api.cu), which also produces the
 * reference's stderr lines and fatal error.  Such introns are rare (a handful per million).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipx_merge.h"

namespace ipx {

__global__ void k_merge_init(MergeTables t) {
    const uint64_t cap = t.cap_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
        t.keys[i] = EMPTY64;
        t.chrom[i] = 0xFFFFFFFFu;
        MStat s;
        s.cnt[0] = s.cnt[1] = s.cnt[2] = 0ull;
        s.a0 = 0xFFFFFFFFu;
        s.a1 = 0u;
        s.conf_bits = 0u;
        s.strand_min = 0xFFFFFFFFu;
        s.strand_max = 0u;
        s.site_min = 0xFFFFFFFFu;
        s.site_max = 0u;
        s.odd = 0u;
        s.first_row = EMPTY64;
        t.stat[i] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *t.n_keys = 0ull;
        *t.overflow = 0u;
    }
}

__global__ void __launch_bounds__(256) k_merge_fold(MergeRowsView r, MergeTables t) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r.n; i += stride) {
        const uint32_t i0 = r.i0[i], i1 = r.i1[i];
        const uint32_t chrom = (uint32_t)r.chrom[i];
        /* Merge rows are not bounded by --max-intron-length, so the key is two words: (start << 32 | end) in the
         * key array and the chromosome beside it.  Both are claimed with CAS and never waited on: whoever first
         * sets the chromosome word of a slot whose key word matches owns the slot for that chromosome. */
        const unsigned long long k0 = ((unsigned long long)i0 << 32) | i1; /* never ~0: start < end */
        uint64_t h = mix64(k0 ^ ((unsigned long long)chrom * 0x9E3779B97F4A7C15ull)) & t.cap_mask;
        uint64_t slot = t.cap_mask + 1;
        for (uint32_t probe = 0; probe < (1u << 16); ++probe) {
            unsigned long long k = *((volatile unsigned long long *)&t.keys[h]);
            if (k == EMPTY64) {
                const unsigned long long old = atomicCAS(&t.keys[h], EMPTY64, k0);
                k = (old == EMPTY64) ? k0 : old;
            }
            if (k == k0) {
                const uint32_t c = atomicCAS(&t.chrom[h], 0xFFFFFFFFu, chrom);
                if (c == 0xFFFFFFFFu) {
                    atomicAdd(t.n_keys, 1ull);
                    slot = h;
                    break;
                }
                if (c == chrom) {
                    slot = h;
                    break;
                }
            }
            h = (h + 1) & t.cap_mask;
        }
        if (slot > t.cap_mask) {
            atomicExch(t.overflow, 1u);
            continue;
        }
        MStat *s = &t.stat[slot];
        atomicAdd(&s->cnt[0], (unsigned long long)r.counts[3 * i + 0]);
        atomicAdd(&s->cnt[1], (unsigned long long)r.counts[3 * i + 1]);
        atomicAdd(&s->cnt[2], (unsigned long long)r.counts[3 * i + 2]);
        atomicMin(&s->a0, r.a0[i]);
        atomicMax(&s->a1, r.a1[i]);
        const float c = r.conf[i];
        if (c >= 0.0f) atomicMax(&s->conf_bits, __float_as_uint(c)); else atomicOr(&s->odd, 1u); /* negative or NaN */
        const uint32_t st = (uint8_t)r.strand[i];
        if (st != (uint32_t)'.') {
            atomicMin(&s->strand_min, st);
            atomicMax(&s->strand_max, st);
        }
        atomicMin(&s->site_min, r.sites[i]);
        atomicMax(&s->site_max, r.sites[i]);
        atomicMin(&s->first_row, (unsigned long long)(r.row_base + i));
    }
}

__global__ void k_merge_compact(MergeTables t, unsigned long long *d_count, unsigned long long *sort_lo,
                                uint32_t *slot_out) {
    const uint64_t cap = t.cap_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < cap; base += stride) {
        const uint64_t i = base + threadIdx.x;
        unsigned long long key = EMPTY64;
        bool take = false;
        if (i < cap) {
            key = t.keys[i];
            take = key != EMPTY64 && t.chrom[i] != 0xFFFFFFFFu;
        }
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, take);
        if (mask == 0u) continue;
        const int lane = threadIdx.x & 31;
        unsigned long long wb = 0ull;
        if (lane == 0) wb = atomicAdd(d_count, (unsigned long long)__popc(mask));
        wb = __shfl_sync(0xFFFFFFFFu, wb, 0);
        if (take) {
            const uint64_t o = wb + __popc(mask & ((1u << lane) - 1u));
            sort_lo[o] = key; /* start << 32 | end */
            slot_out[o] = (uint32_t)i;
        }
    }
}

__global__ void k_merge_rank_keys(MergeTables t, const uint32_t *name_rank, const uint32_t *slots, uint64_t n,
                                  unsigned long long *hi_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hi_out[i] = name_rank[t.chrom[slots[i]]];
}

__global__ void k_merge_gather(MergeTables t, const uint32_t *sorted_slot, uint64_t n, MergeOut out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = sorted_slot[i];
    const unsigned long long key = t.keys[slot];
    const MStat s = t.stat[slot];
    out.chrom[i] = (int32_t)t.chrom[slot];
    out.i0[i] = (uint32_t)(key >> 32);
    out.i1[i] = (uint32_t)key;
    out.a0[i] = s.a0;
    out.a1[i] = s.a1;
    out.counts[3 * i + 0] = s.cnt[0];
    out.counts[3 * i + 1] = s.cnt[1];
    out.counts[3 * i + 2] = s.cnt[2];
    out.conf[i] = __uint_as_float(s.conf_bits);
    const bool one_strand = (s.strand_max == 0u) || (s.strand_min == s.strand_max);
    const bool one_site = s.site_min == s.site_max;
    out.strand[i] = (s.strand_max == 0u) ? '.' : (char)s.strand_min;
    out.site[i] = s.site_min;
    out.slow[i] = (one_strand && one_site && !s.odd) ? 0 : 1;
}

void launch_merge_init(const MergeTables &t, cudaStream_t st) { k_merge_init<<<148 * 4, 256, 0, st>>>(t); }

void launch_merge_fold(const MergeRowsView &r, const MergeTables &t, int sm_count, cudaStream_t st) {
    if (r.n == 0) return;
    uint64_t want = (r.n + 255) / 256, maxb = (uint64_t)sm_count * 8;
    k_merge_fold<<<(unsigned)(want < maxb ? want : maxb), 256, 0, st>>>(r, t);
}

void launch_merge_compact(const MergeTables &t, unsigned long long *d_count, unsigned long long *sort_lo, uint32_t *slot_out,
                          cudaStream_t st) {
    k_merge_compact<<<148 * 4, 256, 0, st>>>(t, d_count, sort_lo, slot_out);
}

void launch_merge_rank_keys(const MergeTables &t, const uint32_t *name_rank, const uint32_t *slots, uint64_t n,
                            unsigned long long *hi_out, cudaStream_t st) {
    if (n == 0) return;
    k_merge_rank_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t, name_rank, slots, n, hi_out);
}

void launch_merge_gather(const MergeTables &t, const uint32_t *sorted_slot, uint64_t n, const MergeOut &out,
                         cudaStream_t st) {
    if (n == 0) return;
    k_merge_gather<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t, sorted_slot, n, out);
}

}  // namespace ipx
