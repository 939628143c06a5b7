/* ipx_finalize.cu -- the finished table on the device (sm_100a): key order, sjN ranking, ordered read-offset
 * histograms, output rows.  One CUDA graph of hand-written kernels over the junction rows in use, never over the
 * table's capacity, inside buffers of grow-only bounds (see FinalizeWork); no library sort.
 *
 *   key order   (chromosome-name rank, intron_start, intron_end) -- the iteration order of the reference's
 *               JunctionTable (junctions.hh:77-100,225-246) -- by a sample sort: splitters from a sorted sample (one
 *               block), one scatter pass into buckets of ~512 keys, one block per bucket sorting in shared memory
 *   sjN         rank of the creating event's stream position (junctions_extractor.cc:458-460), the same sort on T
 *   histograms  Junction::read_offsets (junctions.hh:120) in run-length form, ascending offset: the dense row gives the
 *               offsets < DENSE_W in order; the junction's entries of the read-offset log are gathered into a private
 *               segment (their number is known per junction: olog_cnt), ordered there -- a warp's bitonic sort in
 *               shared memory for up to FIN_SMALL entries, a 65536-bin bitmap + counter pass by a block for the
 *               few junctions with more -- and run-length reduced in place
 * The three chains run as parallel branches of the graph and join in the row gather; an exclusive scan of the runs per
 * junction (ipx_scan.cuh) gives hist_off, and a last pass writes the runs.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <utility>
#include <vector>

#include "ipx_engine.h"
#include "ipx_motif.cuh"
#include "ipx_scan.cuh"

namespace ipx {

enum : unsigned long long { FLAG_B = 1ull, FLAG_L = 2ull, FLAG_R = 4ull, FLAG_C = 8ull };
/* d_count: 0 junctions, 1 histogram runs, 2 log-segment cursor, 3 flags, 4 junction ids in use, 5 log entries in use,
 *          6 junctions whose log entries need the block pass, 7 largest bucket seen */

__global__ void k_fin_begin(Tables t, unsigned long long *d_count, uint64_t B, uint64_t L) {
    const unsigned long long nk = junctions_in_use(t), nl = log_entries_in_use(t);
    d_count[0] = d_count[1] = d_count[2] = 0ull;
    d_count[3] = (nk > B ? FLAG_B : 0ull) | (nl > L ? FLAG_L : 0ull);
    d_count[4] = nk;
    d_count[5] = nl;
    d_count[6] = d_count[7] = 0ull;
}

/* per junction id: the two sort keys, and a private segment for its log entries */
__global__ void __launch_bounds__(256) k_fin_list(Tables t, const uint32_t *name_rank, uint64_t B, unsigned long long *d_count,
                                                  unsigned long long *sort_key, uint32_t *sort_sub, unsigned long long *tkey,
                                                  uint32_t *seg_off, uint32_t *seg_fill) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint64_t nk = d_count[4] < B ? d_count[4] : B;
    bool take = false;
    uint32_t lc = 0u;
    unsigned long long sk = EMPTY64, tk = EMPTY64;
    uint32_t sub = 0u;
    if (j < nk) {
        const unsigned long long T = t.stat[j].T;
        take = T != EMPTY64; /* keys whose every event was rejected by the canon filter are not junctions */
        if (take) {
            const unsigned long long key = t.jkey[j];
            sk = ((unsigned long long)name_rank[t.jtid[j]] << 31) | key_start(t.kl, key);
            sub = key_len(t.kl, key);
            tk = T >> 8;
            lc = t.olog_cnt[j];
        }
    }
    uint32_t incl = lc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
    }
    unsigned long long seg0 = 0ull;
    const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (lane == 31 && warp_total) seg0 = atomicAdd(&d_count[2], (unsigned long long)warp_total);
    seg0 = __shfl_sync(0xFFFFFFFFu, seg0, 31);
    if (j < B) {
        sort_key[j] = sk;
        sort_sub[j] = sub;
        tkey[j] = tk;
        seg_off[j] = (uint32_t)(seg0 + incl - lc);
        seg_fill[j] = 0u;
    }
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, take);
    if (mask && lane == 0) atomicAdd(&d_count[0], (unsigned long long)__popc(mask));
}

/* ------------------------------------------------------------------ sample sort of (u64 key, u32 sub-key) -> u32 value
 * Keys equal to EMPTY64 are not part of the input.  (key, sub) pairs are unique. */
static const int SS_MAX_BUCKETS = 4096;
static const int SS_MAX_SAMPLE = 8192;
static const int SS_BUCKET = 256;       /* keys per bucket the splitters aim at */
/* Elements a bucket may have to be sorted in shared memory.  Junction ids are handed out to thread blocks in blocks of
 * 64, so the keys arrive shuffled over tens of thousands of positions and the splitters behave like a random sample of 3
 * per bucket: one sort in four saw a bucket of more than 4 times the average (1024), which then went the slow way. */
static const int SS_SMEM_BUCKET = 2048;


__device__ __forceinline__ bool pair_less(unsigned long long ka, uint32_t sa, unsigned long long kb, uint32_t sb) {
    return ka < kb || (ka == kb && sa < sb);
}

/* bitonic sort of n = 2^k elements in shared memory by the whole block */
template <bool WITH_PAYLOAD>
__device__ __forceinline__ void block_bitonic(unsigned long long *k, uint32_t *s, uint32_t *v, uint32_t n) {
    for (uint32_t size = 2u; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0u; stride >>= 1) {
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                const uint32_t lo = 2u * i - (i & (stride - 1u)); /* index with bit `stride` clear */
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0u;
                const unsigned long long ka = k[lo], kb = k[hi];
                uint32_t sa = 0u, sb = 0u;
                if (WITH_PAYLOAD) { sa = s[lo]; sb = s[hi]; }
                const bool swap = up ? pair_less(kb, sb, ka, sa) : pair_less(ka, sa, kb, sb);
                if (swap) {
                    k[lo] = kb; k[hi] = ka;
                    if (WITH_PAYLOAD) {
                        s[lo] = sb; s[hi] = sa;
                        const uint32_t x = v[lo]; v[lo] = v[hi]; v[hi] = x;
                    }
                }
            }
        }
    }
    __syncthreads();
}

/* one block: a sorted sample of the keys -> splitters; also resets the bucket counters */
__global__ void __launch_bounds__(1024) k_ss_sample(const unsigned long long *keys, const unsigned long long *n_ptr, uint64_t bound,
                                                    SampleSort ss) {
    extern __shared__ unsigned long long s_sample[];
    const uint64_t n = *n_ptr < bound ? *n_ptr : bound;
    uint32_t P = (uint32_t)((n + SS_BUCKET - 1) / SS_BUCKET);
    if (P < 1u) P = 1u;
    if (P > (uint32_t)SS_MAX_BUCKETS) P = SS_MAX_BUCKETS;
    uint64_t S = 3ull * P;
    if (S > (uint64_t)SS_MAX_SAMPLE) S = SS_MAX_SAMPLE;
    if (S > n) S = n;
    uint32_t S2 = 1u;
    while (S2 < S) S2 <<= 1;
    if (S2 < 2u) S2 = 2u;
    for (uint32_t i = threadIdx.x; i < S2; i += blockDim.x) s_sample[i] = i < S ? keys[(uint64_t)i * n / S] : EMPTY64;
    for (uint32_t i = threadIdx.x; i <= (uint32_t)SS_MAX_BUCKETS; i += blockDim.x) ss.bucket_cnt[i] = i == (uint32_t)SS_MAX_BUCKETS ? P : 0u;
    block_bitonic<false>(s_sample, nullptr, nullptr, S2);
    /* splitter p = first key of bucket p + 1; padding and keys that are not part of the input (EMPTY64) sort last and
     * make empty buckets */
    for (uint32_t p = threadIdx.x; p + 1u < P; p += blockDim.x) ss.splitters[p] = S ? s_sample[(uint64_t)(p + 1u) * S / P] : EMPTY64;
}

/* every key into its bucket: the number of splitters <= key */
__global__ void __launch_bounds__(256) k_ss_scatter(const unsigned long long *keys, const uint32_t *sub, const unsigned long long *n_ptr,
                                                    uint64_t bound, SampleSort ss, unsigned long long *d_count) {
    __shared__ unsigned long long s_split[SS_MAX_BUCKETS];
    const uint32_t P = ss.bucket_cnt[SS_MAX_BUCKETS];
    for (uint32_t p = threadIdx.x; p + 1u < P; p += blockDim.x) s_split[p] = ss.splitters[p];
    __syncthreads();
    const uint64_t n = *n_ptr < bound ? *n_ptr : bound;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned long long key = i < n ? keys[i] : EMPTY64;
    uint32_t lo = 0xFFFFFFFFu - (uint32_t)lane; /* lanes with nothing to place group with nobody */
    if (key != EMPTY64) {
        uint32_t hi = P - 1u; /* number of splitters <= key lies in [lo, hi] */
        lo = 0u;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_split[mid] <= key) lo = mid + 1u; else hi = mid;
        }
    }
    /* keys arrive nearly in order, so a warp's keys mostly share a bucket: one atomic per (warp, bucket) */
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, lo);
    const int leader = __ffs(grp) - 1;
    uint32_t pos = 0u;
    if (key != EMPTY64 && lane == leader) pos = atomicAdd(&ss.bucket_cnt[lo], (uint32_t)__popc(grp));
    pos = __shfl_sync(0xFFFFFFFFu, pos, leader) + (uint32_t)__popc(grp & ((1u << lane) - 1u));
    if (key == EMPTY64) return;
    if (pos < ss.C) {
        const uint64_t at = (uint64_t)lo * ss.C + pos;
        ss.bk[at] = key;
        ss.bs[at] = sub ? sub[i] : 0u;
        ss.bv[at] = (uint32_t)i;
    } else {
        atomicOr(&d_count[3], FLAG_C);
        atomicMax(&d_count[7], (unsigned long long)pos + 1ull);
    }
}

/* One block per bucket: a bitonic sort in shared memory (its cost grows with n log^2 n, so a bucket a few times the
 * average does not hold the launch up the way a quadratic ranking would), then the values go to the bucket's place in
 * the global order (and, when asked, the position of every value).  A bucket that does not fit shared memory is
 * ranked by counting out of global memory (only a pathological input makes such buckets). */
__global__ void __launch_bounds__(256) k_ss_local(SampleSort ss, uint32_t *out_vals, uint32_t *out_rank) {
    __shared__ unsigned long long s_k[SS_SMEM_BUCKET];
    __shared__ uint32_t s_s[SS_SMEM_BUCKET], s_v[SS_SMEM_BUCKET];
    __shared__ uint32_t s_part[8];
    const uint32_t P = ss.bucket_cnt[SS_MAX_BUCKETS];
    const uint32_t b = blockIdx.x;
    if (b >= P) return;
    const uint32_t cnt = ss.bucket_cnt[b] < ss.C ? ss.bucket_cnt[b] : (uint32_t)ss.C;
    /* where the bucket starts: the sizes of the buckets before it */
    uint32_t before = 0u;
    for (uint32_t p = threadIdx.x; p < b; p += blockDim.x) before += ss.bucket_cnt[p] < ss.C ? ss.bucket_cnt[p] : (uint32_t)ss.C;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = before;
    __syncthreads();
    uint32_t base = 0u;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w) base += s_part[w];
    if (cnt == 0u) return;
    const unsigned long long *gk = ss.bk + (uint64_t)b * ss.C;
    const uint32_t *gs = ss.bs + (uint64_t)b * ss.C, *gv = ss.bv + (uint64_t)b * ss.C;
    if (cnt <= (uint32_t)SS_SMEM_BUCKET) {
        uint32_t n2 = 2u;
        while (n2 < cnt) n2 <<= 1;
        for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
            s_k[i] = i < cnt ? gk[i] : EMPTY64;
            s_s[i] = i < cnt ? gs[i] : 0xFFFFFFFFu;
            s_v[i] = i < cnt ? gv[i] : 0u;
        }
        block_bitonic<true>(s_k, s_s, s_v, n2);
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
            out_vals[base + i] = s_v[i];
            if (out_rank) out_rank[s_v[i]] = base + i;
        }
    } else {
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
            const unsigned long long ki = gk[i];
            const uint32_t si = gs[i];
            uint32_t rank = 0u;
            for (uint32_t x = 0; x < cnt; ++x) rank += pair_less(gk[x], gs[x], ki, si) ? 1u : 0u;
            out_vals[base + rank] = gv[i];
            if (out_rank) out_rank[gv[i]] = base + rank;
        }
    }
}

/* IPX_FIN_TRACE=1 (with IPX_NO_GRAPH=1): CUDA events between the steps, everything on one stream, times on stderr */
struct FinTrace {
    bool on;
    cudaStream_t st;
    std::vector<std::pair<const char *, cudaEvent_t>> marks;
    const unsigned long long *counts = nullptr; /* pinned copy of d_count, valid after the last mark */
    void mark(const char *what) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.emplace_back(what, e);
    }
    void report() {
        if (!on || marks.empty()) return;
        cudaEventSynchronize(marks.back().second);
        fprintf(stderr, "[ipx finalize]");
        for (size_t i = 1; i < marks.size(); i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            fprintf(stderr, " %s %.1f", marks[i].first, ms * 1000.f);
        }
        float total = 0.f;
        cudaEventElapsedTime(&total, marks.front().second, marks.back().second);
        fprintf(stderr, " | total %.1f us", total * 1000.f);
        if (counts) fprintf(stderr, " | junctions %llu runs %llu ids %llu log %llu big %llu", counts[0], counts[1], counts[4], counts[5], counts[6]);
        fprintf(stderr, "\n");
        for (auto &m : marks) cudaEventDestroy(m.second);
        marks.clear();
    }
};

/* the sample kernel's shared memory is beyond the default limit: opt in, once per device (not under stream capture) */
static cudaError_t sample_sort_configure() {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_ss_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SS_MAX_SAMPLE * sizeof(unsigned long long)));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    return cudaSuccess;
}

static cudaError_t sample_sort(const unsigned long long *keys, const uint32_t *sub, const unsigned long long *n_ptr, uint64_t bound,
                               const SampleSort &ss, uint32_t *out_vals, uint32_t *out_rank, unsigned long long *d_count,
                               cudaStream_t st, FinTrace *tr) {
    uint64_t p_bound = (bound + SS_BUCKET - 1) / SS_BUCKET;
    if (p_bound < 1) p_bound = 1;
    if (p_bound > (uint64_t)SS_MAX_BUCKETS) p_bound = SS_MAX_BUCKETS;
    k_ss_sample<<<1, 1024, SS_MAX_SAMPLE * sizeof(unsigned long long), st>>>(keys, n_ptr, bound, ss);
    if (tr) tr->mark("sample");
    k_ss_scatter<<<(unsigned)((bound + 255) / 256), 256, 0, st>>>(keys, sub, n_ptr, bound, ss, d_count);
    if (tr) tr->mark("scatter");
    k_ss_local<<<(unsigned)p_bound, 256, 0, st>>>(ss, out_vals, out_rank);
    if (tr) tr->mark("local");
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ read-offset log -> ordered runs per junction */
static const uint32_t FIN_SMALL = 256; /* log entries a warp orders in shared memory */

/* every log entry into its junction's segment: offset << 16 | count */
__global__ void __launch_bounds__(256) k_fin_scatter(Tables t, uint64_t B, uint64_t L, const unsigned long long *d_count,
                                                     const uint32_t *seg_off, uint32_t *seg_fill, uint32_t *seg) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nl = d_count[5] < L ? d_count[5] : L;
    if (i >= nl) return;
    const unsigned long long e = t.olog[i];
    const uint32_t j = (uint32_t)(e >> 32);
    /* a junction beyond the bound this finalize was sized for has no segment: the bound is reported as too small
     * anyway (k_fin_begin) and the finalize runs again */
    if (j >= B || t.stat[j].T == EMPTY64) return;
    const uint32_t pos = atomicAdd(&seg_fill[j], 1u);
    if ((uint64_t)seg_off[j] + pos < L) seg[seg_off[j] + pos] = (uint32_t)e;
}

/* the bases of junction j's dense rows 1 .. (BASE_UNSET: row not given out) */
__device__ __forceinline__ void row_bases(const Tables &t, uint64_t j, uint32_t base[DENSE_ROWS - 1]) {
    const Slot *s = &t.slots[t.jslot[j]];
    base[0] = s->base[0];
    base[1] = s->base[1];
    base[2] = s->base[2];
}

/* non-zero counters of one dense row (all lanes get the sum) */
__device__ __forceinline__ uint32_t dense_nonzeros(const Tables &t, uint64_t j, uint32_t row, int lane) {
    const uint4 *p = reinterpret_cast<const uint4 *>(t.dense + (j * DENSE_ROWS + row) * DENSE_W);
    const uint4 a = p[lane * 2], b = p[lane * 2 + 1];
    uint32_t nz = (a.x != 0) + (a.y != 0) + (a.z != 0) + (a.w != 0) + (b.x != 0) + (b.y != 0) + (b.z != 0) + (b.w != 0);
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) nz += __shfl_xor_sync(0xFFFFFFFFu, nz, k);
    return nz;
}

/* nz4[4 j + r] = non-zero counters of junction j's row r; returns their sum */
__device__ __forceinline__ uint32_t count_dense_rows(const Tables &t, uint64_t j, const uint32_t base[DENSE_ROWS - 1], int lane, uint16_t *nz4) {
    uint32_t total = 0u;
    if (!t.dense) {
        if (lane < (int)DENSE_ROWS) nz4[4 * j + lane] = 0;
        return 0u;
    }
#pragma unroll
    for (uint32_t r = 0; r < DENSE_ROWS; ++r) {
        const uint32_t nz = (r == 0u || base[r - 1] != BASE_UNSET) ? dense_nonzeros(t, j, r, lane) : 0u;
        if (lane == 0) nz4[4 * j + r] = (uint16_t)nz;
        total += nz;
    }
    return total;
}

/* warp per junction id: non-zero counters of its dense rows; its log segment ordered by offset and run-length reduced
 * in place (the counts of the runs, ascending offset, replace the entries); runs_j = runs of both kinds, below3[3 j + r]
 * = log runs with offsets below the first offset of dense row r + 1 (they come before it in the output).  Junctions
 * with more than FIN_SMALL log entries are left to k_fin_big. */
__global__ void __launch_bounds__(256) k_fin_small(Tables t, uint64_t B, uint64_t L, unsigned long long *d_count, const uint32_t *seg_off,
                                                   uint32_t *seg, uint32_t *runs_j, uint16_t *nz4, uint32_t *below3, uint32_t *big_list) {
    __shared__ uint32_t s_w[8][FIN_SMALL];
    __shared__ uint32_t s_c[8][FIN_SMALL];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nk = d_count[4] < B ? d_count[4] : B;
    if (j >= nk) return;
    if (lane < 3) below3[3 * j + lane] = 0u;
    if (t.stat[j].T == EMPTY64) {
        if (lane == 0) runs_j[j] = 0u;
        if (lane < (int)DENSE_ROWS) nz4[4 * j + lane] = 0;
        return;
    }
    uint32_t base[DENSE_ROWS - 1];
    row_bases(t, j, base);
    const uint32_t nz = count_dense_rows(t, j, base, lane, nz4);
    const uint32_t n = t.olog_cnt[j];
    const uint64_t s0 = seg_off[j];
    if (n == 0u || s0 + n > L) { /* (beyond L: the bound was too small, the finalize runs again) */
        if (lane == 0) runs_j[j] = nz;
        return;
    }
    if (n > FIN_SMALL) {
        if (lane == 0) big_list[atomicAdd(&d_count[6], 1ull)] = (uint32_t)j;
        return;
    }
    uint32_t *w = s_w[warp], *c = s_c[warp];
    uint32_t n2 = 2u;
    while (n2 < n) n2 <<= 1;
    for (uint32_t i = lane; i < n2; i += 32u) w[i] = i < n ? seg[s0 + i] : 0xFFFFFFFFu;
    for (uint32_t size = 2u; size <= n2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0u; stride >>= 1) {
            __syncwarp();
            for (uint32_t i = lane; i < (n2 >> 1); i += 32u) {
                const uint32_t lo = 2u * i - (i & (stride - 1u)), hi = lo + stride;
                const uint32_t a = w[lo], b = w[hi];
                if (((lo & size) == 0u) ? (b < a) : (a < b)) { w[lo] = b; w[hi] = a; }
            }
        }
    }
    __syncwarp();
    /* runs of equal offsets (the high half of an entry), their counts summed */
    uint32_t n_runs = 0u, below[DENSE_ROWS - 1] = {0u, 0u, 0u};
    for (uint32_t i0 = 0; i0 < n; i0 += 32u) {
        const uint32_t i = i0 + lane;
        const bool valid = i < n;
        const uint32_t x = valid ? w[i] : 0u;
        const bool head = valid && (i == 0u || (w[i - 1] >> 16) != (x >> 16));
        const unsigned hm = __ballot_sync(0xFFFFFFFFu, head);
        const uint32_t run = n_runs + (uint32_t)__popc(hm & (0xFFFFFFFFu >> (31 - lane))) - 1u;
        if (head) c[run] = 0u;
        __syncwarp();
        if (valid) atomicAdd(&c[run], x & 0xFFFFu);
        n_runs += (uint32_t)__popc(hm);
#pragma unroll
        for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r)
            below[r] += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, head && base[r] != BASE_UNSET && (x >> 16) < base[r]));
        __syncwarp();
    }
    for (uint32_t r = lane; r < n_runs; r += 32u) seg[s0 + r] = c[r];
    if (lane == 0) {
        runs_j[j] = nz + n_runs;
        below3[3 * j + 0] = below[0]; below3[3 * j + 1] = below[1]; below3[3 * j + 2] = below[2];
    }
}

/* Block per junction with many log entries: the entries' counts are summed per offset in a 65536-counter scratch of the
 * block (global memory, zero between uses) while a two-level bitmap in shared memory remembers which offsets occur
 * (64 summary words over 2048 words over 65536 offsets); the runs are then read off the bitmap in ascending order by
 * the 64 threads that own a summary word each.  The cost follows the entries and the distinct offsets, not 65536. */
static const int FIN_BIG_BLOCKS = 148 * 4;
__device__ __forceinline__ uint32_t bits_below(uint32_t bits, uint32_t lo, uint32_t base) {
    /* set bits of a bitmap word that starts at offset lo whose offsets lie below base */
    if (base == BASE_UNSET || lo >= base) return 0u;
    return lo + 32u <= base ? (uint32_t)__popc(bits) : (uint32_t)__popc(bits & ((1u << (base - lo)) - 1u));
}
__global__ void __launch_bounds__(256) k_fin_big(Tables t, uint64_t L, const unsigned long long *d_count, const uint32_t *big_list,
                                                 const uint32_t *seg_off, uint32_t *seg, uint32_t *runs_j, const uint16_t *nz4,
                                                 uint32_t *below3, uint32_t *scratch) {
    __shared__ uint32_t s_bits[2048];
    __shared__ uint32_t s_sum[64];
    __shared__ uint32_t s_runs[64], s_below[DENSE_ROWS - 1][64];
    uint32_t *cnt = scratch + (uint64_t)blockIdx.x * 65536u;
    for (uint32_t i = threadIdx.x; i < 2048u; i += blockDim.x) s_bits[i] = 0u;
    if (threadIdx.x < 64) s_sum[threadIdx.x] = 0u;
    __syncthreads();
    const uint64_t n_big = d_count[6];
    for (uint64_t q = blockIdx.x; q < n_big; q += gridDim.x) {
        const uint32_t j = big_list[q];
        const uint32_t n = t.olog_cnt[j];
        const uint64_t s0 = seg_off[j];
        uint32_t base[DENSE_ROWS - 1] = {BASE_UNSET, BASE_UNSET, BASE_UNSET};
        if (t.dense) row_bases(t, j, base);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t x = seg[s0 + i];
            const uint32_t off = x >> 16;
            const uint32_t bit = 1u << (off & 31u);
            if (!(atomicOr(&s_bits[off >> 5], bit) & bit)) atomicOr(&s_sum[off >> 10], 1u << ((off >> 5) & 31u));
            atomicAdd(&cnt[off], x & 0xFFFFu);
        }
        __threadfence();
        __syncthreads();
        /* thread i < 64 owns summary word i = offsets [1024 i, 1024 i + 1024) */
        if (threadIdx.x < 64) {
            uint32_t mine = 0u, mb[DENSE_ROWS - 1] = {0u, 0u, 0u};
            uint32_t words = s_sum[threadIdx.x];
            while (words) {
                const uint32_t wi = threadIdx.x * 32u + (uint32_t)__ffs(words) - 1u;
                words &= words - 1u;
                const uint32_t bits = s_bits[wi];
                mine += (uint32_t)__popc(bits);
#pragma unroll
                for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r) mb[r] += bits_below(bits, wi * 32u, base[r]);
            }
            s_runs[threadIdx.x] = mine;
#pragma unroll
            for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r) s_below[r][threadIdx.x] = mb[r];
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            uint32_t pre = 0u, total = 0u, below[DENSE_ROWS - 1] = {0u, 0u, 0u};
            for (uint32_t k = 0; k < 64u; ++k) {
                const uint32_t v = s_runs[k];
                if (k < threadIdx.x) pre += v;
                total += v;
#pragma unroll
                for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r) below[r] += s_below[r][k];
            }
            /* total <= n: the runs fit where the entries were (every entry has been read: the barrier above) */
            uint32_t words = s_sum[threadIdx.x];
            while (words) {
                const uint32_t wi = threadIdx.x * 32u + (uint32_t)__ffs(words) - 1u;
                words &= words - 1u;
                uint32_t bits = s_bits[wi];
                s_bits[wi] = 0u;
                while (bits) {
                    const uint32_t off = wi * 32u + (uint32_t)__ffs(bits) - 1u;
                    bits &= bits - 1u;
                    seg[s0 + pre++] = __ldcg(&cnt[off]);
                    cnt[off] = 0u;
                }
            }
            s_sum[threadIdx.x] = 0u;
            if (threadIdx.x == 0) {
                runs_j[j] = (uint32_t)nz4[4 * j] + nz4[4 * j + 1] + nz4[4 * j + 2] + nz4[4 * j + 3] + total;
                below3[3 * j + 0] = below[0]; below3[3 * j + 1] = below[1]; below3[3 * j + 2] = below[2];
            }
        }
        __syncthreads();
    }
    (void)L;
}

/* ------------------------------------------------------------------ output rows */
/* one output row per junction, in key order: attributes of the creating event after motif correction
 * (get_splice_sites + correct_strand, junctions_extractor.cc:419-444); rows [n_j, B] get zero runs for the scan */
__global__ void __launch_bounds__(256) k_fin_rows(Tables t, GenomeView g, uint64_t B, const uint32_t *order, const unsigned long long *d_count,
                                                  const uint32_t *t_rank, const uint32_t *runs_j, uint32_t *runs_r, DeviceTable out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > B) return;
    /* (a sort bucket that overflowed left keys out: `order` is incomplete, the host grows the buckets and replays) */
    if (i >= d_count[0] || i >= B || (d_count[3] & FLAG_C)) {
        runs_r[i] = 0u;
        return;
    }
    const uint32_t j = order[i];
    runs_r[i] = runs_j[j];
    out.ijunc[i] = t_rank[j];
    const unsigned long long key = t.jkey[j];
    const JStat s = t.stat[j];
    const int32_t tid = t.jtid[j];
    const uint32_t i0 = key_start(t.kl, key), i1 = i0 + key_len(t.kl, key);
    out.chrom[i] = tid;
    out.i0[i] = i0;
    out.i1[i] = i1;
    out.a0[i] = s.a0;
    out.a1[i] = s.a1;
    out.counts[3 * i + 0] = s.cnt[0];
    out.counts[3 * i + 1] = s.cnt[1];
    out.counts[3 * i + 2] = s.cnt[2];
    out.T[i] = s.T;
    char strand = (char)(s.T & 0xFFull);
    char site[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    char f[4];
    if (motif_fetch(g, tid, i0, i1, f) == MOTIF_OK) {
        char d0 = f[0], d1 = f[1], a0 = f[2], a1 = f[3];
        const char r0 = comp(f[3]), r1 = comp(f[2]), r2 = comp(f[1]), r3 = comp(f[0]);
        bool canon;
        if (strand == '-') {
            d0 = r0; d1 = r1; a0 = r2; a1 = r3;
            canon = canon4(d0, d1, a0, a1);
        } else if (strand == '.') {
            canon = canon4(d0, d1, a0, a1);
            if (canon) {
                strand = '+';
            } else if (canon4(r0, r1, r2, r3)) {
                strand = '-';
                d0 = r0; d1 = r1; a0 = r2; a1 = r3;
                canon = true;
            }
        } else {
            canon = canon4(d0, d1, a0, a1);
        }
        if (!canon) { d0 = low(d0); d1 = low(d1); a0 = low(a0); a1 = low(a1); }
        site[0] = d0; site[1] = d1; site[2] = '/'; site[3] = a0; site[4] = a1;
    }
    out.strand[i] = strand;
    for (int k = 0; k < 8; ++k) out.sites[8 * i + k] = site[k];
}

struct RunsIn {
    const uint32_t *runs_r;
    __device__ __forceinline__ unsigned long long operator()(uint64_t i) const { return runs_r[i]; }
};

/* warp per junction, key order: its runs in ascending offset order -- row 0 (offsets below DENSE_W), then the other dense
 * rows in the order of their first offsets, each preceded by the log runs that lie below it, then the rest of the log */
__global__ void __launch_bounds__(256) k_fin_write(Tables t, uint64_t B, uint64_t R, const uint32_t *order, unsigned long long *d_count,
                                                   const unsigned long long *hist_off, const uint16_t *nz4, const uint32_t *below3,
                                                   const uint32_t *seg_off, const uint32_t *seg, uint32_t *hist_cnt) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned long long n_runs = hist_off[B]; /* rows beyond the junctions count zero runs */
    if (r == 0 && lane == 0) {
        d_count[1] = n_runs;
        if (n_runs > R) atomicOr(&d_count[3], FLAG_R);
    }
    if (n_runs > R) return; /* does not fit: the host grows R and replays */
    if (r >= d_count[0] || r >= B || (d_count[3] & FLAG_C)) return;
    const uint32_t j = order[r];
    const unsigned long long o0 = hist_off[r], o1 = hist_off[r + 1];
    const uint64_t s0 = seg_off[j];
    uint32_t base[DENSE_ROWS - 1] = {BASE_UNSET, BASE_UNSET, BASE_UNSET};
    uint32_t nz[DENSE_ROWS] = {0u, 0u, 0u, 0u};
    if (t.dense) {
        row_bases(t, j, base);
#pragma unroll
        for (uint32_t k = 0; k < DENSE_ROWS; ++k) nz[k] = nz4[4 * (uint64_t)j + k];
    }
    const uint32_t n_dense = nz[0] + nz[1] + nz[2] + nz[3];
    const uint32_t n_log = (uint32_t)(o1 - o0) - n_dense;
    /* the dense rows in output order: row 0, then rows 1 .. by ascending base (rows not given out last) */
    uint32_t ord[DENSE_ROWS - 1] = {0u, 1u, 2u};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2 - a; ++b)
            if (base[ord[b]] > base[ord[b + 1]]) { const uint32_t x = ord[b]; ord[b] = ord[b + 1]; ord[b + 1] = x; }
    uint64_t o = o0;
    uint32_t q = 0u; /* log runs written so far */
#pragma unroll
    for (uint32_t k = 0; k < DENSE_ROWS; ++k) {
        const uint32_t row = k == 0u ? 0u : ord[k - 1] + 1u;
        if (k > 0u && base[row - 1u] == BASE_UNSET) break;
        if (k > 0u) {
            /* the log runs below this row's first offset */
            const uint32_t upto = below3[3 * (uint64_t)j + (row - 1u)];
            for (uint32_t x = q + lane; x < upto; x += 32u) hist_cnt[o + (x - q)] = seg[s0 + x];
            o += upto - q;
            q = upto;
        }
        if (nz[row]) {
            const uint4 *p = reinterpret_cast<const uint4 *>(t.dense + ((uint64_t)j * DENSE_ROWS + row) * DENSE_W);
            const uint4 a = p[lane * 2], b = p[lane * 2 + 1];
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint32_t mine = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) mine += (v[i] != 0);
            uint32_t pre = mine;
#pragma unroll
            for (int i = 1; i < 32; i <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, pre, i);
                if (lane >= i) pre += y;
            }
            uint64_t at = o + (pre - mine);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (v[i] != 0) hist_cnt[at++] = v[i];
            }
            o += nz[row];
        }
    }
    for (uint32_t x = q + lane; x < n_log; x += 32u) hist_cnt[o + (x - q)] = seg[s0 + x];
}

/* ------------------------------------------------------------------ host side */
#define IPX_CU(x)                          \
    do {                                   \
        cudaError_t e__ = (x);             \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

static inline unsigned blocks_for(uint64_t n, unsigned threads) {
    uint64_t b = (n + threads - 1) / threads;
    return (unsigned)(b ? b : 1);
}

/* a bound for v with some headroom: v plus an eighth, rounded up to a multiple of 4096, at least floor_ */
static inline uint64_t bound_for(uint64_t v, uint64_t floor_) {
    uint64_t b = ((v + v / 8 + 4095) / 4096) * 4096;
    return b > floor_ ? b : floor_;
}

template <class T>
static cudaError_t regrow(T **p, size_t count) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    return cudaMalloc((void **)p, (count ? count : 1) * sizeof(T));
}

void finalize_release(FinalizeWork &w) {
    if (w.exec) cudaGraphExecDestroy(w.exec);
    w.exec = nullptr;
    cudaFree(w.sort_key); cudaFree(w.sort_sub); cudaFree(w.tkey); cudaFree(w.order); cudaFree(w.t_order); cudaFree(w.t_rank);
    cudaFree(w.seg_off); cudaFree(w.seg_fill); cudaFree(w.runs_j); cudaFree(w.nz4); cudaFree(w.below3); cudaFree(w.big_list); cudaFree(w.runs_r);
    cudaFree(w.seg); cudaFree(w.big_scratch); cudaFree(w.scan_state); cudaFree(w.d_count);
    for (int k = 0; k < 2; ++k) {
        cudaFree(w.ss_splitters[k]); cudaFree(w.ss_bucket_cnt[k]); cudaFree(w.ss_bk[k]); cudaFree(w.ss_bs[k]); cudaFree(w.ss_bv[k]);
    }
    if (w.s2) cudaStreamDestroy(w.s2);
    if (w.s3) cudaStreamDestroy(w.s3);
    if (w.ev_fork) cudaEventDestroy(w.ev_fork);
    if (w.ev_join2) cudaEventDestroy(w.ev_join2);
    if (w.ev_join3) cudaEventDestroy(w.ev_join3);
    if (w.h_count) cudaFreeHost(w.h_count);
    DeviceTable &o = w.out;
    cudaFree(o.chrom); cudaFree(o.i0); cudaFree(o.i1); cudaFree(o.a0); cudaFree(o.a1); cudaFree(o.ijunc); cudaFree(o.strand);
    cudaFree(o.sites); cudaFree(o.counts); cudaFree(o.T); cudaFree(o.hist_off); cudaFree(o.hist_cnt);
    w = FinalizeWork();
}

/* the same sort for other callers (the merge fold): flags in d_count[3] (FLAG_C = 8: a bucket overflowed; d_count[7] = its size) */
uint64_t sample_sort_buckets(uint64_t bound) {
    uint64_t p = (bound + SS_BUCKET - 1) / SS_BUCKET;
    if (p < 1) p = 1;
    return p > (uint64_t)SS_MAX_BUCKETS ? (uint64_t)SS_MAX_BUCKETS : p;
}
uint64_t sample_sort_max_buckets() { return SS_MAX_BUCKETS; }
cudaError_t sample_sort_enqueue(const unsigned long long *keys, const uint32_t *sub, const unsigned long long *n_ptr, uint64_t bound,
                                const SampleSort &ss, uint32_t *out_vals, uint32_t *out_rank, unsigned long long *d_count, cudaStream_t st) {
    cudaError_t e = sample_sort_configure();
    if (e != cudaSuccess) return e;
    return sample_sort(keys, sub, n_ptr, bound, ss, out_vals, out_rank, d_count, st, nullptr);
}

static uint64_t bucket_bound(uint64_t B) {
    uint64_t p = (B + SS_BUCKET - 1) / SS_BUCKET;
    if (p < 1) p = 1;
    return p > (uint64_t)SS_MAX_BUCKETS ? (uint64_t)SS_MAX_BUCKETS : p;
}

/* (re)allocate everything whose size follows the bounds */
static cudaError_t finalize_allocate(FinalizeWork &w, uint64_t B, uint64_t L, uint64_t R, uint64_t C) {
    if (!w.d_count) IPX_CU(cudaMalloc(&w.d_count, 8 * sizeof(unsigned long long)));
    if (!w.h_count) IPX_CU(cudaMallocHost(&w.h_count, 8 * sizeof(unsigned long long)));
    if (!w.s2) {
        IPX_CU(cudaStreamCreateWithFlags(&w.s2, cudaStreamNonBlocking));
        IPX_CU(cudaStreamCreateWithFlags(&w.s3, cudaStreamNonBlocking));
        IPX_CU(cudaEventCreateWithFlags(&w.ev_fork, cudaEventDisableTiming));
        IPX_CU(cudaEventCreateWithFlags(&w.ev_join2, cudaEventDisableTiming));
        IPX_CU(cudaEventCreateWithFlags(&w.ev_join3, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            IPX_CU(regrow(&w.ss_splitters[k], SS_MAX_BUCKETS));
            IPX_CU(regrow(&w.ss_bucket_cnt[k], SS_MAX_BUCKETS + 1));
        }
        IPX_CU(sample_sort_configure());
        IPX_CU(cudaMalloc(&w.big_scratch, (size_t)FIN_BIG_BLOCKS * 65536 * sizeof(uint32_t)));
        IPX_CU(cudaMemset(w.big_scratch, 0, (size_t)FIN_BIG_BLOCKS * 65536 * sizeof(uint32_t)));
    }
    if (B != w.B || C != w.C) {
        const uint64_t cells = bucket_bound(B) * C;
        for (int k = 0; k < 2; ++k) {
            IPX_CU(regrow(&w.ss_bk[k], cells)); IPX_CU(regrow(&w.ss_bs[k], cells)); IPX_CU(regrow(&w.ss_bv[k], cells));
        }
        w.C = C;
    }
    if (B != w.B) {
        IPX_CU(regrow(&w.sort_key, B)); IPX_CU(regrow(&w.sort_sub, B)); IPX_CU(regrow(&w.tkey, B)); IPX_CU(regrow(&w.order, B));
        IPX_CU(regrow(&w.t_order, B)); IPX_CU(regrow(&w.t_rank, B));
        /* what a sort leaves unwritten (see FLAG_C) must at least be junction ids */
        IPX_CU(cudaMemset(w.order, 0, B * sizeof(uint32_t))); IPX_CU(cudaMemset(w.t_order, 0, B * sizeof(uint32_t)));
        IPX_CU(cudaMemset(w.t_rank, 0, B * sizeof(uint32_t))); IPX_CU(regrow(&w.seg_off, B)); IPX_CU(regrow(&w.seg_fill, B));
        IPX_CU(regrow(&w.runs_j, B)); IPX_CU(regrow(&w.nz4, B * 4)); IPX_CU(regrow(&w.below3, B * 3)); IPX_CU(regrow(&w.big_list, B)); IPX_CU(regrow(&w.runs_r, B + 1));
        IPX_CU(regrow(&w.scan_state, scan_state_words(B + 1)));
        DeviceTable &o = w.out;
        IPX_CU(regrow(&o.chrom, B)); IPX_CU(regrow(&o.i0, B)); IPX_CU(regrow(&o.i1, B)); IPX_CU(regrow(&o.a0, B));
        IPX_CU(regrow(&o.a1, B)); IPX_CU(regrow(&o.ijunc, B)); IPX_CU(regrow(&o.strand, B)); IPX_CU(regrow(&o.sites, B * 8));
        IPX_CU(regrow(&o.counts, B * 3)); IPX_CU(regrow(&o.T, B)); IPX_CU(regrow(&o.hist_off, B + 2));
        w.B = B;
    }
    if (L != w.L) {
        IPX_CU(regrow(&w.seg, L));
        w.L = L;
    }
    if (R != w.R) {
        IPX_CU(regrow(&w.out.hist_cnt, R));
        w.R = R;
    }
    return cudaSuccess;
}

/* the sequence itself, issued into `st` (under stream capture, or directly when capture is not possible) */
static cudaError_t finalize_issue(const Tables &t, const GenomeView &g, const uint32_t *d_name_rank, FinalizeWork &w, cudaStream_t st,
                                  bool trace = false) {
    const uint64_t B = w.B, L = w.L, R = w.R;
    DeviceTable &out = w.out;
    FinTrace tr{trace, st, {}};
    cudaStream_t s2 = trace ? st : w.s2, s3 = trace ? st : w.s3;
    tr.mark("start");
    launch_resolve_T(t, g, st); /* creation thresholds for anything still unresolved */
    k_fin_begin<<<1, 1, 0, st>>>(t, w.d_count, B, L);
    k_fin_list<<<blocks_for(B, 256), 256, 0, st>>>(t, d_name_rank, B, w.d_count, w.sort_key, w.sort_sub, w.tkey, w.seg_off, w.seg_fill);
    tr.mark("resolve+begin+list");
    /* three independent chains from here: key order (st), sjN ranking (s2), read-offset log (s3) */
    if (!trace) {
        IPX_CU(cudaEventRecord(w.ev_fork, st));
        IPX_CU(cudaStreamWaitEvent(w.s2, w.ev_fork, 0));
        IPX_CU(cudaStreamWaitEvent(w.s3, w.ev_fork, 0));
    }
    const unsigned long long *n_ptr = w.d_count + 4;

    /* st: key order -- chromosome-name rank, intron_start, intron_end */
    SampleSort ssa{w.ss_splitters[0], w.ss_bucket_cnt[0], w.ss_bk[0], w.ss_bs[0], w.ss_bv[0], w.C};
    IPX_CU(sample_sort(w.sort_key, w.sort_sub, n_ptr, B, ssa, w.order, nullptr, w.d_count, st, &tr));
    tr.mark("key-sort");

    /* s2: sjN = rank of the creating event in stream order */
    SampleSort ssb{w.ss_splitters[1], w.ss_bucket_cnt[1], w.ss_bk[1], w.ss_bs[1], w.ss_bv[1], w.C};
    IPX_CU(sample_sort(w.tkey, nullptr, n_ptr, B, ssb, w.t_order, w.t_rank, w.d_count, s2, &tr));
    tr.mark("T-sort");
    if (!trace) IPX_CU(cudaEventRecord(w.ev_join2, s2));

    /* s3: the log entries gathered per junction, ordered and run-length reduced */
    k_fin_scatter<<<blocks_for(L, 256), 256, 0, s3>>>(t, B, L, w.d_count, w.seg_off, w.seg_fill, w.seg);
    tr.mark("log-scatter");
    k_fin_small<<<blocks_for(B * 32, 256), 256, 0, s3>>>(t, B, L, w.d_count, w.seg_off, w.seg, w.runs_j, w.nz4, w.below3, w.big_list);
    tr.mark("small");
    k_fin_big<<<FIN_BIG_BLOCKS, 256, 0, s3>>>(t, L, w.d_count, w.big_list, w.seg_off, w.seg, w.runs_j, w.nz4, w.below3, w.big_scratch);
    tr.mark("big");
    if (!trace) {
        IPX_CU(cudaEventRecord(w.ev_join3, s3));
        /* join: rows in key order (motif, strand correction, sjN), then the histograms */
        IPX_CU(cudaStreamWaitEvent(st, w.ev_join2, 0));
        IPX_CU(cudaStreamWaitEvent(st, w.ev_join3, 0));
    }
    k_fin_rows<<<blocks_for(B + 1, 256), 256, 0, st>>>(t, g, B, w.order, w.d_count, w.t_rank, w.runs_j, w.runs_r, out);
    tr.mark("rows");
    IPX_CU(scan_exclusive(RunsIn{w.runs_r}, out.hist_off, B + 1, w.scan_state, st));
    tr.mark("scan");
    k_fin_write<<<blocks_for(B * 32, 256), 256, 0, st>>>(t, B, R, w.order, w.d_count, out.hist_off, w.nz4, w.below3, w.seg_off, w.seg,
                                                        out.hist_cnt);
    tr.mark("write");
    IPX_CU(cudaMemcpyAsync(w.h_count, w.d_count, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    w.launches_per_replay = 16;
    tr.mark("counts");
    tr.counts = w.h_count;
    tr.report();
    return cudaGetLastError();
}

cudaError_t finalize_enqueue(const Tables &t, const GenomeView &g, const uint32_t *d_name_rank, FinalizeWork &w, cudaStream_t st) {
    uint64_t B = w.B ? w.B : w.first_B, L = w.L ? w.L : w.first_L, R = w.R ? w.R : w.first_R, C = w.C ? w.C : 2048;
    if (w.wish_B > B) B = w.wish_B;
    if (w.wish_L > L) L = w.wish_L;
    if (w.wish_R > R) R = w.wish_R;
    if (w.wish_C > C) C = w.wish_C;
    FinalizeWork::Key key{};
    key.B = B; key.L = L; key.R = R; key.C = C;
    key.tab_keys = t.slots; key.tab_rows = t.stat; key.tab_log = t.olog; key.max_j = t.max_j; key.olog_cap = t.olog_cap;
    key.tab_base = t.jslot; key.name_rank = d_name_rank; key.gseq = g.seq; key.target_seq = g.target_seq; key.dense = t.dense;
    key.n_targets = g.n_targets; key.have_genome = g.have_genome; key.skip_missing = g.skip_missing; key.filter_canon = g.filter_canon;
    key.wide = t.kl.wide; key.len_bits = t.kl.len_bits;
    if (!w.exec || !(key == w.key)) {
        IPX_CU(cudaStreamSynchronize(st)); /* a replay may still be running */
        if (w.exec) {
            cudaGraphExecDestroy(w.exec);
            w.exec = nullptr;
        }
        IPX_CU(finalize_allocate(w, B, L, R, C));
        w.key = key;
        static const bool no_graph = getenv("IPX_NO_GRAPH") != nullptr;
        if (!no_graph) {
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal); /* s2 / s3 join the capture through the fork event */
            if (e == cudaSuccess) {
                cudaError_t ei = finalize_issue(t, g, d_name_rank, w, st);
                e = cudaStreamEndCapture(st, &graph);
                if (ei != cudaSuccess) e = ei;
            }
            if (e == cudaSuccess && graph) e = cudaGraphInstantiate(&w.exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                cudaGetLastError(); /* capture is not available here: issue the sequence directly every time */
                w.exec = nullptr;
            }
        }
    }
    if (w.exec) return cudaGraphLaunch(w.exec, st);
    static const bool trace = getenv("IPX_FIN_TRACE") != nullptr;
    return finalize_issue(t, g, d_name_rank, w, st, trace);
}

bool finalize_check(FinalizeWork &w) {
    const unsigned long long flags = w.h_count[3];
    if (flags == 0ull) {
        w.out.n = w.h_count[0];
        w.out.n_hist = w.h_count[1];
        return true;
    }
    /* grow what was too small; the next finalize_enqueue reallocates and re-captures */
    if (flags & FLAG_B) w.wish_B = bound_for(w.h_count[4], w.B + 4096);
    if (flags & FLAG_L) w.wish_L = bound_for(w.h_count[5], w.L + 4096);
    if (flags & FLAG_R) w.wish_R = bound_for(w.h_count[1], w.R + 4096);
    if (flags & FLAG_C) {
        uint64_t c = w.C * 2;
        while (c < w.h_count[7]) c *= 2;
        w.wish_C = c;
    }
    w.out.n = 0;
    w.out.n_hist = 0;
    return false;
}

}  // namespace ipx
