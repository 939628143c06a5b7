/* ipx_kernels.cu -- sm_100a kernels of the junction-extraction path.
 *
 *  k_extract<MODE>   K1+K2 fused: one thread walks one alignment's CIGAR (ipx_walk.cuh) and folds every
 *                    junction event straight into the L2-resident junction / offset tables, so events
 *                    never travel through HBM.  HBM traffic is the record columns and the live CIGAR
 *                    words, read once, coalesced.
 *                      MODE_SINGLE  no canon filter: insert + first-seen + counts + anchors + offsets
 *                      MODE_PASS1   canon filter:    insert + first-seen per strand class only
 *                      MODE_PASS2   canon filter:    count events at or after the creation threshold T
 *  k_resolve_T       K3: donor/acceptor gather from the HBM genome, per junction, and the choice of the
 *                    creating event T = first event whose (key, strand class) passes the filter
 *  finalize_tables   compaction, key-order sort, sjN ranking, offset-histogram ordering, output gather
 *  k_import_*        fold partial tables received from other GPUs (multi-GPU exchange)
 *
 * Reference semantics: SURVEY.md section 8(a) A1-A8 (junctions_extractor.cc:49-95,124-319,322-530).
 */
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipx_engine.h"

namespace ipx {

/* ------------------------------------------------------------------ table initialisation */
/* zero the dense rows in use; runs before k_init_tables resets n_keys */
__global__ void k_clear_dense(Tables t) {
    unsigned long long rows = *t.n_keys;
    if (rows > t.dense_rows) rows = t.dense_rows;
    const uint64_t n4 = rows * (DENSE_W / 4);
    uint4 *p = reinterpret_cast<uint4 *>(t.dense);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void k_init_tables(Tables t) {
    const uint64_t cap = t.cap_mask + 1, hcap = t.hcap_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
        t.keys[i] = EMPTY64;
        t.dense_idx[i] = DENSE_NONE;
        JStat s;
        s.first[0] = s.first[1] = s.first[2] = EMPTY64;
        s.T = EMPTY64;
        s.cnt[0] = s.cnt[1] = s.cnt[2] = 0ull;
        s.a0 = 0xFFFFFFFFu;
        s.a1 = 0u;
        t.stat[i] = s;
    }
    HEntry e;
    e.key = EMPTY64;
    e.cnt = 0u;
    e.pad = 0u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hcap; i += stride) t.hist[i] = e;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *t.n_keys = 0ull;
        *t.n_hist = 0ull;
        *t.err_sv = EMPTY64;
        *t.overflow = 0u;
    }
}

void launch_init_tables(const Tables &t, cudaStream_t st) {
    k_clear_dense<<<148 * 8, 256, 0, st>>>(t);
    k_init_tables<<<148 * 8, 256, 0, st>>>(t);
}

/* ------------------------------------------------------------------ K1: extract */
static const int EX_THREADS = 256;
static const int EX_RPT = 4;                       /* records per thread per tile */
static const int EX_TILE = EX_THREADS * EX_RPT;    /* records per tile */
static const int EX_EVCAP = 1024;                  /* staged events per tile (32 KB); more go out one by one */

/* K1 sink: junction_qc, then stage the event in shared memory (or, past the staging capacity, write it out
 * through a per-event reservation) */
struct EventSink {
    const Tables &t;
    const ExtractParams &p;
    Event *stage;
    uint32_t *stage_count;
    Event *out;
    unsigned long long *out_count;
    int32_t tid;
    int64_t pos;
    uint64_t record_seq;
    uint32_t cat;
    char strand;
    int cls;
    bool may_emit;
    uint32_t n_events;

    __device__ __forceinline__ void operator()(const Candidate &c) {
        if (!may_emit || !junction_qc(p, c)) return;
        n_events++;
        if (c.intron_start < 0 || c.intron_start > 0x7FFFFFFFll) {
            atomicExch(t.overflow, 3u); /* outside what a BAM position can hold */
            return;
        }
        Event e;
        e.key = make_key(t.kl, tid, c.intron_start, c.intron_end);
        e.sv = make_sv(record_seq, c.ordinal, strand);
        e.a0 = (uint32_t)c.anchor_start;
        e.a1 = (uint32_t)c.anchor_end;
        e.off_cat = ((uint32_t)(c.intron_start - pos) & 0xFFFFu) | (cat << 16) | ((uint32_t)cls << 18);
        e.pad = 0u;
        const uint32_t idx = atomicAdd(stage_count, 1u);
        if (idx < (uint32_t)EX_EVCAP) stage[idx] = e;
        else out[atomicAdd(out_count, 1ull)] = e; /* a tile of very long reads: rare, still exact */
    }
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
    return v;
}

/* K1.  One block works through tiles of EX_TILE consecutive records.  Every thread loads the fixed-width columns of
 * its EX_RPT records up front (coalesced, all loads in flight together), walks the CIGAR of the live ones
 * (ipx_walk.cuh) and stages their junction events in shared memory; the tile then reserves a contiguous range of
 * the event buffer with ONE atomic and copies its events out with coalesced 16-byte stores.  The event buffer is
 * sized by the host for the worst case (one event per two CIGAR words), so nothing can be lost. */
__global__ void __launch_bounds__(EX_THREADS) k_extract(BatchView b, ExtractParams p, Tables t, Event *out,
                                                        unsigned long long *out_count) {
    __shared__ Event stage[EX_EVCAP];
    __shared__ uint32_t stage_count;
    __shared__ unsigned long long tile_base;
    unsigned long long live_words = 0ull, events = 0ull;
    const uint64_t n_tiles = (b.n + EX_TILE - 1) / EX_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) stage_count = 0u;
        __syncthreads();
        const uint64_t base = tile * EX_TILE + threadIdx.x;
        uint32_t c0[EX_RPT], c1[EX_RPT], flag[EX_RPT];
        int32_t tid[EX_RPT];
#pragma unroll
        for (int j = 0; j < EX_RPT; ++j) {
            const uint64_t r = base + (uint64_t)j * EX_THREADS;
            c0[j] = c1[j] = 0u;
            flag[j] = 0u;
            tid[j] = -1;
            if (r < b.n) {
                c0[j] = __ldg(&b.cigar_off[r]);
                c1[j] = __ldg(&b.cigar_off[r + 1]);
                tid[j] = __ldg(&b.tid[r]);
                flag[j] = __ldg(&b.flag[r]);
            }
        }
#pragma unroll
        for (int j = 0; j < EX_RPT; ++j) {
            const uint32_t n_cigar = c1[j] - c0[j];
            if (!record_is_live(n_cigar, tid[j], flag[j])) continue;
            const uint64_t r = base + (uint64_t)j * EX_THREADS;
            live_words += n_cigar;
            const int64_t pos = __ldg(&b.pos[r]);
            const uint32_t nh = __ldg(&b.nh[r]);
            const uint32_t xs = __ldg(&b.xs[r]);
            const char strand = read_strand(p.strandness, flag[j], xs);
            EventSink sink{t, p, stage, &stage_count, out, out_count, tid[j], pos, b.seq_base + r,
                           read_category(flag[j], nh), strand, strand_class(strand), read_may_emit(p, flag[j], nh), 0u};
            WalkState w;
            w.init(pos);
            uint32_t bad = 0u;
            const uint32_t *cig = b.cigar + c0[j];
            for (uint32_t k = 0; k < n_cigar; ++k) {
                bad = walk_op(w, __ldg(&cig[k]), p.max_anchor_indel_size, sink);
                if (bad) break;
            }
            if (bad) {
                /* the reference throws here; keep the first such record in stream order */
                atomicMin(t.err_sv, make_err_cigar(b.seq_base + r, bad & 15u));
            } else {
                walk_finish(w, sink);
            }
            events += sink.n_events;
        }
        __syncthreads();
        const uint32_t n_ev = stage_count < (uint32_t)EX_EVCAP ? stage_count : (uint32_t)EX_EVCAP;
        if (threadIdx.x == 0 && n_ev) tile_base = atomicAdd(out_count, (unsigned long long)n_ev);
        __syncthreads();
        if (n_ev) {
            /* 32-byte events as two 16-byte halves: consecutive threads write consecutive 16-byte words */
            const uint4 *src = reinterpret_cast<const uint4 *>(stage);
            uint4 *dst = reinterpret_cast<uint4 *>(out + tile_base);
            for (uint32_t i = threadIdx.x; i < n_ev * 2u; i += EX_THREADS) dst[i] = src[i];
        }
        __syncthreads();
    }
    /* one atomic per warp for the unit counters */
    live_words = warp_sum(live_words);
    events = warp_sum(events);
    if ((threadIdx.x & 31) == 0) {
        if (live_words) atomicAdd(t.n_live, live_words);
        if (events) atomicAdd(t.n_events, events);
    }
}

void launch_extract(const BatchView &b, const ExtractParams &p, const Tables &t, Event *out, unsigned long long *out_count,
                    int sm_count, cudaStream_t st) {
    if (b.n == 0) return;
    const uint64_t n_tiles = (b.n + EX_TILE - 1) / EX_TILE;
    const uint64_t max_blocks = (uint64_t)sm_count * 6; /* 33 KB of shared memory per block */
    const int blocks = (int)(n_tiles < max_blocks ? n_tiles : max_blocks);
    k_extract<<<blocks, EX_THREADS, 0, st>>>(b, p, t, out, out_count);
}

/* ------------------------------------------------------------------ K2: keyed aggregation */
/* fold one event into the tables on its own (import kernels and tests of the slow path) */
template <int MODE>
__device__ __forceinline__ uint32_t fold_event_direct(const Tables &t, unsigned long long key, unsigned long long sv,
                                                      uint32_t a0, uint32_t a1, uint32_t off, uint32_t cat, int cls) {
    if (MODE == MODE_PASS2) {
        const uint64_t slot = junction_find(t, key);
        if (slot > t.cap_mask) return 0u;
        JStat *s = &t.stat[slot];
        const unsigned long long T = *((volatile unsigned long long *)&s->T);
        if (T == EMPTY64 || sv < T) return 0u; /* before the junction exists: dropped */
        stat_accumulate(s, cat, a0, a1);
        return hist_count(t, slot, off, 1u);
    } else {
        const uint64_t slot = junction_find_or_insert(t, key);
        if (slot > t.cap_mask) return 0u;
        JStat *s = &t.stat[slot];
        stat_first(s, cls, sv);
        if (MODE == MODE_SINGLE) {
            stat_accumulate(s, cat, a0, a1);
            return hist_count(t, slot, off, 1u);
        }
        return 0u;
    }
}

/* K2.  Every warp takes 32 consecutive events (neighbouring reads, so mostly the same few junctions), groups them
 * by key with match.any, and lets the lowest lane of each group probe the junction table once and apply the
 * group's combined read counts and anchor extremes; each lane then counts its own read offset (one fire-and-forget
 * atomic into the junction's dense offset row, or the offset hash table for large offsets).
 *   MODE_SINGLE  no canon filter: insert + first-seen + counts + anchors + offsets
 *   MODE_PASS1   canon filter:    insert + first-seen per strand class only
 *   MODE_PASS2   canon filter:    count events at or after the creation threshold T */
template <int MODE>
__global__ void __launch_bounds__(256) k_fold(Tables t, const Event *ev, const unsigned long long *n_events_ptr) {
    const unsigned long long n_ev = *n_events_ptr;
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t inserts = 0u;
    for (uint64_t base = warp0 * 32; base < n_ev; base += n_warps * 32) {
        const uint64_t i = base + lane;
        const bool valid = i < n_ev;
        Event e;
        if (valid) {
            const uint4 *src = reinterpret_cast<const uint4 *>(ev + i);
            const uint4 lo = __ldg(src), hi = __ldg(src + 1);
            e.key = ((unsigned long long)lo.y << 32) | lo.x;
            e.sv = ((unsigned long long)lo.w << 32) | lo.z;
            e.a0 = hi.x; e.a1 = hi.y; e.off_cat = hi.z;
        } else {
            e.key = EMPTY64 - 2 - (unsigned)lane; e.sv = EMPTY64; e.a0 = 0xFFFFFFFFu; e.a1 = 0u; e.off_cat = 0u;
        }
        const uint32_t cat = (e.off_cat >> 16) & 3u;
        const int cls = (int)((e.off_cat >> 18) & 3u);
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, e.key);
        const int leader = __ffs(grp) - 1;
        uint64_t slot = t.cap_mask + 1;
        uint32_t drow = DENSE_NONE;
        unsigned long long T = 0ull;
        if (lane == leader && valid) {
            if (MODE == MODE_PASS2) {
                slot = junction_find(t, e.key);
                if (slot <= t.cap_mask) T = *((volatile unsigned long long *)&t.stat[slot].T);
            } else {
                slot = junction_find_or_insert(t, e.key, &drow);
            }
            if (MODE != MODE_PASS1 && slot <= t.cap_mask && drow == DENSE_NONE)
                drow = *((volatile uint32_t *)&t.dense_idx[slot]);
        }
        slot = __shfl_sync(grp, slot, leader);
        bool counted = valid && slot <= t.cap_mask;
        if (MODE == MODE_PASS2) {
            T = __shfl_sync(grp, T, leader);
            counted = counted && T != EMPTY64 && e.sv >= T; /* earlier events are dropped */
        }
        if (MODE != MODE_PASS2 && counted) stat_first(&t.stat[slot], cls, e.sv);
        if (MODE != MODE_PASS1) {
            drow = __shfl_sync(grp, drow, leader);
            const unsigned c0 = __ballot_sync(0xFFFFFFFFu, counted && cat == 0u) & grp;
            const unsigned c1 = __ballot_sync(0xFFFFFFFFu, counted && cat == 1u) & grp;
            const unsigned c2 = __ballot_sync(0xFFFFFFFFu, counted && cat == 2u) & grp;
            const uint32_t a0 = __reduce_min_sync(grp, counted ? e.a0 : 0xFFFFFFFFu);
            const uint32_t a1 = __reduce_max_sync(grp, counted ? e.a1 : 0u);
            if (lane == leader && (c0 | c1 | c2)) {
                JStat *s = &t.stat[slot];
                if (c0) atomicAdd(&s->cnt[0], (unsigned long long)__popc(c0));
                if (c1) atomicAdd(&s->cnt[1], (unsigned long long)__popc(c1));
                if (c2) atomicAdd(&s->cnt[2], (unsigned long long)__popc(c2));
                if (a0 < *((volatile uint32_t *)&s->a0)) atomicMin(&s->a0, a0);
                if (a1 > *((volatile uint32_t *)&s->a1)) atomicMax(&s->a1, a1);
            }
            if (counted) {
                const uint32_t off = e.off_cat & 0xFFFFu;
                if (off < DENSE_W && drow != DENSE_NONE) atomicAdd(&t.dense[(uint64_t)drow * DENSE_W + off], 1u);
                else inserts += hist_add_hash(t, slot, off, 1u);
            }
        }
    }
    unsigned long long ins = warp_sum((unsigned long long)inserts);
    if (lane == 0 && ins) atomicAdd(t.n_hist, ins);
}

void launch_fold(int mode, const Tables &t, const Event *ev, const unsigned long long *n_events_ptr, int sm_count,
                 cudaStream_t st) {
    const int blocks = sm_count * 8;
    switch (mode) {
        case MODE_SINGLE: k_fold<MODE_SINGLE><<<blocks, 256, 0, st>>>(t, ev, n_events_ptr); break;
        case MODE_PASS1: k_fold<MODE_PASS1><<<blocks, 256, 0, st>>>(t, ev, n_events_ptr); break;
        default: k_fold<MODE_PASS2><<<blocks, 256, 0, st>>>(t, ev, n_events_ptr); break;
    }
}

/* ------------------------------------------------------------------ K3: motif gather and threshold */
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

/* For every junction slot without a creation threshold yet: T = first event (stream order) whose strand class
 * passes junction_filter_check (add_junction, junctions_extractor.cc:494-502; SURVEY.md 8a "A6"). */
__global__ void k_resolve_T(Tables t, GenomeView g) {
    const uint64_t cap = t.cap_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
        const unsigned long long key = t.keys[i];
        if (key == EMPTY64) continue;
        JStat *s = &t.stat[i];
        if (s->T != EMPTY64) continue;
        const unsigned long long f0 = s->first[CLS_PLUS], f1 = s->first[CLS_MINUS], f2 = s->first[CLS_DOT];
        bool p0 = true, p1 = true, p2 = true;
        const int32_t tid = key_tid(t.kl, key);
        const uint32_t i0 = key_start(t.kl, key), i1 = i0 + key_len(t.kl, key);
        char f[4];
        const int rc = motif_fetch(g, tid, i0, i1, f);
        if (rc >= MOTIF_ERR_MISSING) {
            /* Genome::fetch throws at the first event that needs the sequence */
            unsigned long long fm = f0 < f1 ? f0 : f1;
            fm = fm < f2 ? fm : f2;
            if (fm != EMPTY64) atomicMin(t.err_sv, make_err_genome(fm, rc));
            continue;
        }
        if (rc == MOTIF_OK && g.filter_canon) {
            p0 = canon4(f[0], f[1], f[2], f[3]);
            p1 = canon4(comp(f[3]), comp(f[2]), comp(f[1]), comp(f[0]));
            p2 = p0 || p1;
        }
        unsigned long long T = EMPTY64;
        if (p0 && f0 < T) T = f0;
        if (p1 && f1 < T) T = f1;
        if (p2 && f2 < T) T = f2;
        s->T = T;
    }
}

void launch_resolve_T(const Tables &t, const GenomeView &g, cudaStream_t st) {
    k_resolve_T<<<148 * 8, 256, 0, st>>>(t, g);
}

/* ------------------------------------------------------------------ finalize */
/* gather every created junction: its slot and its key-order sort key */
__global__ void k_compact_junctions(Tables t, const uint32_t *name_rank, unsigned long long *d_count,
                                    unsigned long long *sort_key, uint32_t *slot_out) {
    const uint64_t cap = t.cap_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < cap; base += stride) {
        const uint64_t i = base + threadIdx.x;
        bool take = false;
        unsigned long long key = EMPTY64;
        if (i < cap) {
            key = t.keys[i];
            take = (key != EMPTY64) && (t.stat[i].T != EMPTY64);
        }
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, take);
        if (mask == 0u) continue;
        const int lane = threadIdx.x & 31;
        unsigned long long warp_base = 0ull;
        if (lane == 0) warp_base = atomicAdd(&d_count[0], (unsigned long long)__popc(mask));
        warp_base = __shfl_sync(0xFFFFFFFFu, warp_base, 0);
        if (take) {
            const uint64_t o = warp_base + __popc(mask & ((1u << lane) - 1u));
            const uint64_t tid = (uint64_t)key_tid(t.kl, key);
            const uint64_t low = key & ((1ull << t.kl.tid_shift) - 1ull);
            sort_key[o] = ((unsigned long long)name_rank[tid] << t.kl.tid_shift) | low;
            slot_out[o] = (uint32_t)i;
        }
    }
}

__global__ void k_fill_slot_rank(const uint32_t *sorted_slot, uint64_t n, uint32_t *slot_rank) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_rank[sorted_slot[i]] = (uint32_t)i;
}

__global__ void k_gather_T(Tables t, const uint32_t *sorted_slot, uint64_t n, unsigned long long *T_out,
                           uint32_t *idx_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        T_out[i] = t.stat[sorted_slot[i]].T;
        idx_out[i] = (uint32_t)i;
    }
}

__global__ void k_scatter_ijunc(const uint32_t *idx_by_T, uint64_t n, uint32_t *ijunc) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ijunc[idx_by_T[i]] = (uint32_t)i;
}

/* one output row per junction, in key order: attributes of the creating event after motif correction
 * (get_splice_sites + correct_strand, junctions_extractor.cc:419-444) */
__global__ void k_gather_rows(Tables t, GenomeView g, const uint32_t *sorted_slot, uint64_t n, DeviceTable out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = sorted_slot[i];
    const unsigned long long key = t.keys[slot];
    const JStat s = t.stat[slot];
    const int32_t tid = key_tid(t.kl, key);
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
    if (out.slot) out.slot[i] = slot;
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

/* every (junction, offset) pair of a created junction -> (rank << 16 | offset, count); one atomic per block step */
static const int CH_PER_THREAD = 8;
__global__ void __launch_bounds__(256) k_compact_hist(Tables t, const uint32_t *slot_rank, unsigned long long *d_count,
                                                      unsigned long long *hkey_out, uint32_t *hcnt_out, int count_only) {
    typedef cub::BlockScan<uint32_t, 256> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ unsigned long long block_base;
    const uint64_t hcap = t.hcap_mask + 1;
    const uint64_t step = 256ull * CH_PER_THREAD;
    for (uint64_t base = (uint64_t)blockIdx.x * step; base < hcap; base += (uint64_t)gridDim.x * step) {
        unsigned long long hk[CH_PER_THREAD];
        uint32_t cnt[CH_PER_THREAD], rank[CH_PER_THREAD];
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < CH_PER_THREAD; ++j) {
            const uint64_t i = base + (uint64_t)j * 256 + threadIdx.x;
            hk[j] = EMPTY64;
            cnt[j] = 0u;
            if (i < hcap) {
                const HEntry e = t.hist[i];
                hk[j] = e.key;
                cnt[j] = e.cnt;
            }
        }
#pragma unroll
        for (int j = 0; j < CH_PER_THREAD; ++j) {
            rank[j] = 0xFFFFFFFFu;
            if (hk[j] != EMPTY64 && hk[j] != TOMB64) rank[j] = slot_rank[hk[j] >> 16];
            if (rank[j] != 0xFFFFFFFFu) mine++;
        }
        uint32_t offset, total;
        Scan(scan_tmp).ExclusiveSum(mine, offset, total);
        if (threadIdx.x == 0 && total) block_base = atomicAdd(&d_count[1], (unsigned long long)total);
        __syncthreads();
        if (!count_only && total) {
            uint64_t o = block_base + offset;
#pragma unroll
            for (int j = 0; j < CH_PER_THREAD; ++j) {
                if (rank[j] != 0xFFFFFFFFu) {
                    hkey_out[o] = ((unsigned long long)rank[j] << 16) | (hk[j] & 0xFFFFull);
                    hcnt_out[o] = cnt[j];
                    o++;
                }
            }
        }
        __syncthreads();
    }
}

/* entries with a small offset that raced ahead of their junction's dense row: move them into the row */
__global__ void k_drain_small(Tables t) {
    const uint64_t hcap = t.hcap_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hcap; i += stride) {
        const HEntry e = t.hist[i];
        if (e.key == EMPTY64 || e.key == TOMB64) continue;
        const uint32_t off = (uint32_t)(e.key & 0xFFFFull);
        if (off >= DENSE_W) continue;
        const uint32_t d = t.dense_idx[e.key >> 16];
        if (d == DENSE_NONE) continue;
        atomicAdd(&t.dense[(uint64_t)d * DENSE_W + off], e.cnt);
        t.hist[i].key = TOMB64;
    }
}

/* per junction (key order): first index and number of its entries in the sorted hash-run list */
__global__ void k_hash_runs(const unsigned long long *sorted_hkey, uint64_t n_hash, unsigned long long *hash_lo,
                            unsigned long long *total) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hash) return;
    const uint64_t r = sorted_hkey[i] >> 16;
    if (i == 0 || (sorted_hkey[i - 1] >> 16) != r) hash_lo[r] = i;
    atomicAdd(&total[r], 1ull);
}

/* warp per junction: number of non-zero counters of its dense row, added to total[r] */
__global__ void k_dense_count(Tables t, const uint32_t *sorted_slot, uint64_t n, unsigned long long *total,
                              uint32_t *dense_nz) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const uint32_t d = t.dense_idx[sorted_slot[r]];
    uint32_t nz = 0;
    if (d != DENSE_NONE) {
        const uint4 *row = reinterpret_cast<const uint4 *>(t.dense + (uint64_t)d * DENSE_W);
        const uint4 a = row[lane * 2], b = row[lane * 2 + 1];
        nz = (a.x != 0) + (a.y != 0) + (a.z != 0) + (a.w != 0) + (b.x != 0) + (b.y != 0) + (b.z != 0) + (b.w != 0);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) nz += __shfl_xor_sync(0xFFFFFFFFu, nz, k);
    if (lane == 0) {
        dense_nz[r] = nz;
        total[r] += nz;
    }
}

/* warp per junction: write its runs in ascending offset order -- dense row first (offsets < DENSE_W), then its
 * slice of the sorted hash runs (all >= DENSE_W once k_drain_small has run, or everything when it has no row) */
__global__ void k_write_runs(Tables t, const uint32_t *sorted_slot, uint64_t n, const unsigned long long *hist_off,
                             const uint32_t *dense_nz, const unsigned long long *hash_lo,
                             const unsigned long long *sorted_hkey, const uint32_t *sorted_hcnt, uint32_t *hist_cnt,
                             uint32_t *hist_offset) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const unsigned long long o0 = hist_off[r];
    const uint32_t nz = dense_nz[r];
    const uint32_t d = t.dense_idx[sorted_slot[r]];
    if (d != DENSE_NONE && nz) {
        const uint4 *row = reinterpret_cast<const uint4 *>(t.dense + (uint64_t)d * DENSE_W);
        const uint4 a = row[lane * 2], b = row[lane * 2 + 1];
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) mine += (v[k] != 0);
        uint32_t pre = mine;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, pre, k);
            if (lane >= k) pre += y;
        }
        uint64_t o = o0 + (pre - mine);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (v[k] != 0) {
                hist_cnt[o] = v[k];
                if (hist_offset) hist_offset[o] = (uint32_t)(lane * 8 + k);
                o++;
            }
        }
    }
    const unsigned long long n_hash = (hist_off[r + 1] - o0) - nz;
    if (n_hash) {
        const unsigned long long lo = hash_lo[r];
        for (unsigned long long j = lane; j < n_hash; j += 32) {
            hist_cnt[o0 + nz + j] = sorted_hcnt[lo + j];
            if (hist_offset) hist_offset[o0 + nz + j] = (uint32_t)(sorted_hkey[lo + j] & 0xFFFFull);
        }
    }
}

#define IPX_CU(x)                          \
    do {                                   \
        cudaError_t e__ = (x);             \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

static cudaError_t ensure_tmp(FinalizeWork &w, size_t bytes, cudaStream_t st) {
    if (bytes > w.cub_tmp_bytes) {
        if (w.cub_tmp) dfree(st, w.cub_tmp);
        w.cub_tmp = nullptr;
        w.cub_tmp_bytes = 0;
        IPX_CU(dmalloc(st, &w.cub_tmp, bytes));
        w.cub_tmp_bytes = bytes;
    }
    return cudaSuccess;
}

void free_device_table(DeviceTable &t, cudaStream_t st) {
    dfree(st, t.chrom); dfree(st, t.i0); dfree(st, t.i1); dfree(st, t.a0); dfree(st, t.a1); dfree(st, t.ijunc);
    dfree(st, t.strand); dfree(st, t.sites); dfree(st, t.counts); dfree(st, t.T); dfree(st, t.hist_off);
    dfree(st, t.hist_cnt); dfree(st, t.hist_offset); dfree(st, t.slot);
    t = DeviceTable();
}

static inline unsigned blocks_for(uint64_t n, unsigned threads) {
    uint64_t b = (n + threads - 1) / threads;
    return (unsigned)(b ? b : 1);
}

cudaError_t finalize_tables(const Tables &t, const GenomeView &g, const uint32_t *d_name_rank, int32_t n_targets,
                            FinalizeWork &w, DeviceTable &out, bool keep_export, cudaStream_t st, int *own_launches) {
    (void)n_targets;
    int launches = 0;
    const uint64_t cap = t.cap_mask + 1;
    out = DeviceTable();
    if (!w.d_count) IPX_CU(dmalloc(st, &w.d_count, 4 * sizeof(unsigned long long)));
    if (w.slot_rank_cap < cap) {
        if (w.slot_rank) dfree(st, w.slot_rank);
        IPX_CU(dmalloc(st, &w.slot_rank, cap * sizeof(uint32_t)));
        w.slot_rank_cap = cap;
    }
    IPX_CU(cudaMemsetAsync(w.d_count, 0, 4 * sizeof(unsigned long long), st));
    IPX_CU(cudaMemsetAsync(w.slot_rank, 0xFF, cap * sizeof(uint32_t), st));

    /* creation thresholds for anything still unresolved */
    k_resolve_T<<<148 * 8, 256, 0, st>>>(t, g);
    launches++;

    /* number of keys bounds the number of junctions */
    unsigned long long n_keys = 0, n_hash_entries = 0;
    IPX_CU(cudaMemcpyAsync(&n_keys, t.n_keys, sizeof(n_keys), cudaMemcpyDeviceToHost, st));
    IPX_CU(cudaMemcpyAsync(&n_hash_entries, t.n_hist, sizeof(n_hash_entries), cudaMemcpyDeviceToHost, st));
    IPX_CU(cudaStreamSynchronize(st));
    const uint64_t nk = n_keys ? n_keys : 1;

    unsigned long long *sort_key = nullptr, *sort_key2 = nullptr;
    uint32_t *slot_a = nullptr, *slot_b = nullptr;
    IPX_CU(dmalloc(st, &sort_key, nk * 8));
    IPX_CU(dmalloc(st, &sort_key2, nk * 8));
    IPX_CU(dmalloc(st, &slot_a, nk * 4));
    IPX_CU(dmalloc(st, &slot_b, nk * 4));
    k_compact_junctions<<<148 * 8, 256, 0, st>>>(t, d_name_rank, w.d_count, sort_key, slot_a);
    launches++;
    unsigned long long n_j = 0;
    IPX_CU(cudaMemcpyAsync(&n_j, &w.d_count[0], sizeof(n_j), cudaMemcpyDeviceToHost, st));
    IPX_CU(cudaStreamSynchronize(st));
    const uint64_t n = n_j, na = n ? n : 1;

    /* key order: chromosome-name rank, intron_start, intron_end */
    size_t tmp_bytes = 0;
    IPX_CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, sort_key, sort_key2, slot_a, slot_b, (int)n, 0, 64, st));
    IPX_CU(ensure_tmp(w, tmp_bytes, st));
    if (n) IPX_CU(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp_bytes, sort_key, sort_key2, slot_a, slot_b, (int)n, 0, 64, st));
    const uint32_t *sorted_slot = slot_b;
    if (n) {
        k_fill_slot_rank<<<blocks_for(n, 256), 256, 0, st>>>(sorted_slot, n, w.slot_rank);
        launches++;
    }

    IPX_CU(dmalloc(st, &out.chrom, na * 4));
    IPX_CU(dmalloc(st, &out.i0, na * 4));
    IPX_CU(dmalloc(st, &out.i1, na * 4));
    IPX_CU(dmalloc(st, &out.a0, na * 4));
    IPX_CU(dmalloc(st, &out.a1, na * 4));
    IPX_CU(dmalloc(st, &out.ijunc, na * 4));
    IPX_CU(dmalloc(st, &out.strand, na));
    IPX_CU(dmalloc(st, &out.sites, na * 8));
    IPX_CU(dmalloc(st, &out.counts, na * 24));
    IPX_CU(dmalloc(st, &out.T, na * 8));
    IPX_CU(dmalloc(st, &out.hist_off, (na + 2) * 8));
    if (keep_export) IPX_CU(dmalloc(st, &out.slot, na * 4));
    out.n = n;

    /* sjN: rank of the creating event in stream order */
    if (n) {
        uint32_t *idx_a = slot_a; /* reuse */
        uint32_t *idx_b = nullptr;
        IPX_CU(dmalloc(st, &idx_b, na * 4));
        k_gather_T<<<blocks_for(n, 256), 256, 0, st>>>(t, sorted_slot, n, sort_key, idx_a);
        size_t tb = 0;
        IPX_CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, sort_key, sort_key2, idx_a, idx_b, (int)n, 0, 64, st));
        IPX_CU(ensure_tmp(w, tb, st));
        IPX_CU(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, sort_key, sort_key2, idx_a, idx_b, (int)n, 0, 64, st));
        k_scatter_ijunc<<<blocks_for(n, 256), 256, 0, st>>>(idx_b, n, out.ijunc);
        k_gather_rows<<<blocks_for(n, 256), 256, 0, st>>>(t, g, sorted_slot, n, out);
        launches += 3; /* gather_T, scatter_ijunc, gather_rows */
        IPX_CU(cudaStreamSynchronize(st));
        dfree(st, idx_b);
    }

    /* offset histogram.  Hash part first (only long reads / pool overflow put anything there): fold raced
     * small-offset entries into their dense rows, then order the rest by (junction rank, offset). */
    unsigned long long *hk_a = nullptr, *hk_b = nullptr, *hash_lo = nullptr, *total = nullptr;
    uint32_t *hc_a = nullptr, *hc_b = nullptr, *dense_nz = nullptr;
    uint64_t nh = 0;
    if (n_hash_entries) {
        k_drain_small<<<148 * 8, 256, 0, st>>>(t);
        k_compact_hist<<<148 * 8, 256, 0, st>>>(t, w.slot_rank, w.d_count, nullptr, nullptr, 1);
        launches += 2;
        unsigned long long n_h = 0;
        IPX_CU(cudaMemcpyAsync(&n_h, &w.d_count[1], sizeof(n_h), cudaMemcpyDeviceToHost, st));
        IPX_CU(cudaStreamSynchronize(st));
        nh = n_h;
    }
    const uint64_t nha = nh ? nh : 1;
    IPX_CU(dmalloc(st, &hash_lo, na * 8));
    IPX_CU(dmalloc(st, &total, (na + 1) * 8));
    IPX_CU(dmalloc(st, &dense_nz, na * 4));
    IPX_CU(cudaMemsetAsync(total, 0, (na + 1) * 8, st));
    if (nh) {
        IPX_CU(dmalloc(st, &hk_a, nha * 8));
        IPX_CU(dmalloc(st, &hk_b, nha * 8));
        IPX_CU(dmalloc(st, &hc_a, nha * 4));
        IPX_CU(dmalloc(st, &hc_b, nha * 4));
        IPX_CU(cudaMemsetAsync(&w.d_count[1], 0, sizeof(unsigned long long), st));
        k_compact_hist<<<148 * 8, 256, 0, st>>>(t, w.slot_rank, w.d_count, hk_a, hc_a, 0);
        int end_bit = 16;
        while (end_bit < 64 && (n >> (end_bit - 16)) != 0) end_bit++;
        size_t tb = 0;
        IPX_CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, hk_a, hk_b, hc_a, hc_b, (int)nh, 0, end_bit, st));
        IPX_CU(ensure_tmp(w, tb, st));
        IPX_CU(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, hk_a, hk_b, hc_a, hc_b, (int)nh, 0, end_bit, st));
        k_hash_runs<<<blocks_for(nh, 256), 256, 0, st>>>(hk_b, nh, hash_lo, total);
        launches += 2;
    }
    unsigned long long n_runs = 0;
    if (n) {
        k_dense_count<<<blocks_for(n * 32, 256), 256, 0, st>>>(t, sorted_slot, n, total, dense_nz);
        launches++;
    }
    {
        size_t tb = 0;
        IPX_CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, total, out.hist_off, (int)(n + 1), st));
        IPX_CU(ensure_tmp(w, tb, st));
        IPX_CU(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, total, out.hist_off, (int)(n + 1), st));
    }
    IPX_CU(cudaMemcpyAsync(&n_runs, out.hist_off + n, sizeof(n_runs), cudaMemcpyDeviceToHost, st));
    IPX_CU(cudaStreamSynchronize(st));
    out.n_hist = n_runs;
    IPX_CU(dmalloc(st, &out.hist_cnt, (n_runs ? n_runs : 1) * 4));
    if (keep_export) IPX_CU(dmalloc(st, &out.hist_offset, (n_runs ? n_runs : 1) * 4));
    if (n) {
        k_write_runs<<<blocks_for(n * 32, 256), 256, 0, st>>>(t, sorted_slot, n, out.hist_off, dense_nz, hash_lo, hk_b, hc_b,
                                                            out.hist_cnt, out.hist_offset);
        launches++;
    }
    if (own_launches) *own_launches = launches;
    IPX_CU(cudaStreamSynchronize(st));
    dfree(st, hk_a); dfree(st, hk_b); dfree(st, hc_a); dfree(st, hc_b);
    dfree(st, hash_lo); dfree(st, total); dfree(st, dense_nz);
    dfree(st, sort_key); dfree(st, sort_key2); dfree(st, slot_a); dfree(st, slot_b);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ multi-GPU exchange */
__global__ void k_export_rows(Tables t, DeviceTable dt, JunctionRow *jr, HistRow *hr) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dt.n) {
        const JStat s = t.stat[dt.slot[i]];
        JunctionRow r;
        r.first[0] = s.first[0]; r.first[1] = s.first[1]; r.first[2] = s.first[2];
        r.cnt[0] = s.cnt[0]; r.cnt[1] = s.cnt[1]; r.cnt[2] = s.cnt[2];
        r.tid = dt.chrom[i]; r.i0 = dt.i0[i]; r.i1 = dt.i1[i]; r.a0 = s.a0; r.a1 = s.a1; r.pad = 0;
        jr[i] = r;
    }
    if (i < dt.n_hist) {
        /* owning junction of run i: last r with hist_off[r] <= i */
        uint64_t lo = 0, hi = dt.n;
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (dt.hist_off[mid] <= i) lo = mid; else hi = mid;
        }
        HistRow h;
        h.tid = dt.chrom[lo]; h.i0 = dt.i0[lo]; h.i1 = dt.i1[lo];
        h.cnt = dt.hist_cnt[i]; h.offset = dt.hist_offset[i]; h.pad = 0;
        hr[i] = h;
    }
}

void launch_export_rows(const Tables &t, const DeviceTable &dt, JunctionRow *jr, HistRow *hr, cudaStream_t st) {
    const uint64_t n = dt.n > dt.n_hist ? dt.n : dt.n_hist;
    if (n == 0) return;
    k_export_rows<<<blocks_for(n, 256), 256, 0, st>>>(t, dt, jr, hr);
}

__global__ void k_import_junction_rows(Tables t, const JunctionRow *jr, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const JunctionRow r = jr[i];
    const uint64_t key = make_key(t.kl, r.tid, r.i0, r.i1);
    const uint64_t slot = junction_find_or_insert(t, key);
    if (slot > t.cap_mask) return;
    JStat *s = &t.stat[slot];
    for (int c = 0; c < 3; ++c) {
        if (r.first[c] != EMPTY64) atomicMin(&s->first[c], r.first[c]);
        if (r.cnt[c]) atomicAdd(&s->cnt[c], r.cnt[c]);
    }
    atomicMin(&s->a0, r.a0);
    atomicMax(&s->a1, r.a1);
}

__global__ void k_import_hist_rows(Tables t, const HistRow *hr, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const HistRow r = hr[i];
    const uint64_t key = make_key(t.kl, r.tid, r.i0, r.i1);
    const uint64_t slot = junction_find_or_insert(t, key);
    if (slot > t.cap_mask) return;
    const uint32_t ins = hist_count(t, slot, r.offset & 0xFFFFu, r.cnt);
    if (ins) atomicAdd(t.n_hist, (unsigned long long)ins);
}

void launch_import_rows(const Tables &t, const JunctionRow *jr, uint64_t nj, const HistRow *hr, uint64_t nh,
                        cudaStream_t st) {
    if (nj) k_import_junction_rows<<<blocks_for(nj, 256), 256, 0, st>>>(t, jr, nj);
    if (nh) k_import_hist_rows<<<blocks_for(nh, 256), 256, 0, st>>>(t, hr, nh);
}

}  // namespace ipx
