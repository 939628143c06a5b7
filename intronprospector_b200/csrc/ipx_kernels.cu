/* ipx_kernels.cu -- sm_100a kernels of the junction-extraction path.
 *
 *  k_extract_tma     K1, the hot kernel: persistent blocks (two per SM: a producer warp + 15 worker warps) stream
 *                    tiles of EX_TILE consecutive records through a 4-stage shared-memory ring filled by TMA bulk
 *                    copies (cp.async.bulk + mbarrier), so every record column and CIGAR word is read from HBM
 *                    exactly once, coalesced, with the latency hidden behind the previous tiles' work.  The workers
 *                    (a) classify the tile's records and compact the live ones (process_alignment's skip test) into
 *                    two lists -- CIGARs of at most three operations, which can only be "aligned, N, aligned", and
 *                    longer ones -- then (b) walk the listed records with full warps (ipx_walk.cuh) and write each
 *                    junction event straight into the tile's private region of the event buffer.
 *  k_extract_sorted  K1 for long reads: records taken longest first, lean 32-bit automaton, 32-byte CIGAR groups;
 *                    k_seg_* walk outliers (tens of thousands of operations) segment by segment.
 *  k_extract_simple  K1 for what is left: the last partial tile, columns that are not 16-byte aligned.
 *  k_fold<MODE>      K2: warp-grouped keyed aggregation of the events into the junction / offset tables
 *                      MODE_SINGLE  no canon filter: insert + first-seen + counts + anchors + offsets
 *                      MODE_PASS1   canon filter:    insert + first-seen per strand class only
 *                      MODE_PASS2   canon filter:    count events at or after the creation threshold T
 *  k_resolve_T       K3: donor/acceptor gather from the HBM genome, per junction, and the choice of the
 *                    creating event T = first event whose (key, strand class) passes the filter
 *  finalize_*        key-order sort, sjN ranking, offset-histogram ordering, output gather -- one CUDA graph over
 *                    the list of slots in use, never over the table's capacity
 *  k_export_* / k_import_*   partial tables packed for / folded from other GPUs (all-to-all transport)
 *  k_mail_*          the same exchange over peer memory: rows stored straight into the owners' mailboxes
 *
 * Reference semantics: SURVEY.md section 8(a) A1-A8 (junctions_extractor.cc:49-95,124-319,322-530).
 */
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "ipx_engine.h"

namespace ipx {

/* ------------------------------------------------------------------ table initialisation */
__device__ __forceinline__ void reset_stat(JStat *s) {
    uint4 *p = reinterpret_cast<uint4 *>(s);
    p[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); /* first[0], first[1] */
    p[1] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); /* first[2], T */
    p[2] = make_uint4(0u, 0u, 0u, 0u);                                     /* cnt[0], cnt[1] */
    p[3] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);                            /* cnt[2], a0 = none, a1 = 0 */
}

__global__ void k_init_tables(Tables t) {
    const uint64_t cap = t.cap_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
        t.keys[i] = EMPTY64;
        reset_stat(&t.stat[i]);
    }
}

__global__ void k_reset_counters(Tables t) {
    *t.n_keys = 0ull;
    *t.n_olog = 0ull;
    *t.err_sv = EMPTY64;
    *t.overflow = 0u;
}

/* clear(): undo exactly what the slot list says was touched -- a warp per junction slot (key, 64-byte record, 1 KB
 * offset row); the offset log is emptied by resetting its counter */
__global__ void __launch_bounds__(256) k_clear_tables(Tables t) {
    const unsigned long long nk = *t.n_keys;
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp0; i < nk; i += n_warps) {
        const uint64_t slot = t.slot_list[i];
        if (lane == 0) {
            t.keys[slot] = EMPTY64;
            reset_stat(&t.stat[slot]);
        }
        if (t.dense) {
            uint4 *row = reinterpret_cast<uint4 *>(t.dense + slot * DENSE_W);
            row[lane] = make_uint4(0u, 0u, 0u, 0u);
            row[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

void launch_init_tables(const Tables &t, cudaStream_t st) {
    k_init_tables<<<148 * 8, 256, 0, st>>>(t);
    k_reset_counters<<<1, 1, 0, st>>>(t);
}

void launch_clear_tables(const Tables &t, cudaStream_t st) {
    k_clear_tables<<<148 * 4, 256, 0, st>>>(t);
    k_reset_counters<<<1, 1, 0, st>>>(t);
}

/* ------------------------------------------------------------------ compact batch -> columns */
struct OpsOrZero {
    const uint16_t *n_ops;
    uint64_t n;
    __host__ __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return i < n ? (uint32_t)n_ops[i] : 0u; }
};

__device__ __forceinline__ uint64_t run_of(const unsigned long long *run_start, uint64_t n_runs, uint64_t i) {
    /* last run that starts at or before record i */
    uint64_t lo = 0, hi = n_runs;
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (run_start[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

/* a warp fills 128 consecutive records; almost always they lie in one run, found once by lane 0 */
__global__ void __launch_bounds__(256) k_expand_tid(const int32_t *run_tid, const unsigned long long *run_start, uint64_t n_runs,
                                                    uint64_t n, int32_t *tid) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint64_t first = warp * 128;
    if (first >= n) return;
    const uint64_t last = first + 127 < n ? first + 127 : n - 1;
    uint64_t r = 0, r_end = 0;
    if (lane == 0) {
        r = run_of(run_start, n_runs, first);
        r_end = r + 1 < n_runs ? run_start[r + 1] : ~0ull;
    }
    r = __shfl_sync(0xFFFFFFFFu, r, 0);
    r_end = __shfl_sync(0xFFFFFFFFu, r_end, 0);
    if (last < r_end) {
        const int32_t v = run_tid[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t i = first + (uint64_t)j * 32 + lane;
            if (i < n) tid[i] = v;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t i = first + (uint64_t)j * 32 + lane;
            if (i < n) tid[i] = run_tid[run_of(run_start, n_runs, i)];
        }
    }
}

cudaError_t expand_compact(const uint16_t *n_ops, uint64_t n, uint32_t *cigar_off, const int32_t *run_tid,
                           const unsigned long long *run_start, uint64_t n_runs, int32_t *tid, void *tmp, size_t *tmp_bytes,
                           cudaStream_t st) {
    auto in = thrust::make_transform_iterator(thrust::make_counting_iterator<uint64_t>(0), OpsOrZero{n_ops, n});
    if (!tmp) return cub::DeviceScan::ExclusiveSum(nullptr, *tmp_bytes, in, cigar_off, (int)(n + 1), st);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, *tmp_bytes, in, cigar_off, (int)(n + 1), st);
    if (e != cudaSuccess) return e;
    if (n && n_runs) k_expand_tid<<<(unsigned)((n + 1023) / 1024), 256, 0, st>>>(run_tid, run_start, n_runs, n, tid);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ K1: extract */
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
    return v;
}

/* first event index of a tile's region in the event buffer (see ipx_engine.h) */
__host__ __device__ __forceinline__ uint64_t tile_event_base(uint32_t first_cigar_word) {
    return ((uint64_t)first_cigar_word + 1ull) >> 1;
}

/* K1 sink: junction_qc, then write the event into the next free slot of the tile's region */
struct EventSink {
    const Tables &t;
    const ExtractParams &p;
    Event *region;         /* the tile's region of the event buffer */
    uint32_t *region_fill; /* shared-memory counter of the events written to it */
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
        const unsigned long long key = make_key(t.kl, tid, c.intron_start, c.intron_end);
        const unsigned long long sv = make_sv(record_seq, c.ordinal, strand);
        const uint32_t off_cat = ((uint32_t)(c.intron_start - pos) & 0xFFFFu) | (cat << 16) | ((uint32_t)cls << 18);
        const uint32_t idx = atomicAdd(region_fill, 1u);
        /* one 32-byte event = one DRAM sector: two 16-byte stores from the emitting lane */
        uint4 *dst = reinterpret_cast<uint4 *>(region + idx);
        __stcs(dst, make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)sv, (uint32_t)(sv >> 32)));
        __stcs(dst + 1, make_uint4((uint32_t)c.anchor_start, (uint32_t)c.anchor_end, off_cat, 0u));
    }
};

__device__ __forceinline__ EventSink make_sink(const Tables &t, const ExtractParams &p, Event *region, uint32_t *region_fill,
                                               int32_t tid, int64_t pos, uint64_t seq, uint32_t flag, uint32_t nh, uint32_t xs) {
    const char strand = read_strand(p.strandness, flag, xs);
    return EventSink{t, p, region, region_fill, tid, pos, seq, read_category(flag, nh), strand, strand_class(strand),
                     read_may_emit(p, flag, nh), 0u};
}

/* ---- generic K1: one block per tile, records and CIGAR words straight from global memory ---- */
static const int EXS_THREADS = 256;

__global__ void __launch_bounds__(EXS_THREADS) k_extract_simple(BatchView b, ExtractParams p, Tables t, Event *out,
                                                                TileInfo *tile_info, uint64_t first_tile, uint64_t n_tiles) {
    __shared__ uint32_t region_fill;
    unsigned long long live_words = 0ull, events = 0ull;
    for (uint64_t tile = first_tile + blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) region_fill = 0u;
        __syncthreads();
        const uint64_t r0 = tile * EX_TILE;
        const uint32_t first_word = __ldg(&b.cigar_off[r0]);
        Event *region = out + tile_event_base(first_word);
        for (uint64_t r = r0 + threadIdx.x; r < r0 + EX_TILE && r < b.n; r += EXS_THREADS) {
            const uint32_t c0 = __ldg(&b.cigar_off[r]), n_cigar = __ldg(&b.cigar_off[r + 1]) - c0;
            const int32_t tid = __ldg(&b.tid[r]);
            const uint32_t flag = __ldg(&b.flag[r]);
            if (!record_is_live(n_cigar, tid, flag)) continue;
            live_words += n_cigar;
            const int64_t pos = __ldg(&b.pos[r]);
            EventSink sink = make_sink(t, p, region, &region_fill, tid, pos, b.seq_base + r, flag, __ldg(&b.nh[r]), __ldg(&b.xs[r]));
            WalkState w;
            w.init(pos);
            uint32_t bad = 0u;
            const uint32_t *cig = b.cigar + c0;
            for (uint32_t k = 0; k < n_cigar; ++k) {
                bad = walk_op_flat(w, __ldg(&cig[k]), p.max_anchor_indel_size, sink);
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
        if (threadIdx.x == 0) tile_info[tile] = TileInfo{region_fill, first_word, 0u, 0u};
        __syncthreads();
    }
    live_words = warp_sum(live_words);
    events = warp_sum(events);
    if ((threadIdx.x & 31) == 0) {
        if (live_words) atomicAdd(t.n_live, live_words);
        if (events) atomicAdd(t.n_events, events);
    }
}

/* ---- K1 for long reads: records taken longest first, two-path 32-bit automaton, 16-byte CIGAR loads ----
 * A long read has hundreds to thousands of CIGAR operations and read lengths are log-normal.  Thread-per-record in
 * stream order loses twice: a warp idles behind its longest read (ncu: 17.6 of 32 lanes active), and the general
 * 64-bit automaton spends ~100 instructions per operation.  Here
 *   - k_long_keys + a radix sort order the records by operation count, longest first; warps claim 32 consecutive
 *     records of that order from a counter, so the lanes of a warp run the same number of iterations and the long
 *     reads do not end up in the tail;
 *   - lanes walk with lean_op (ipx_walk.cuh): positions relative to the record start in 32 bits and a common path
 *     of a dozen instructions for aligned runs and short indels inside an anchor;
 *   - a lane reads its record's words as aligned 16-byte groups, the next group requested before the current one is
 *     walked (one load instruction per four operations, and the four sectors of a 128-byte line are used while hot).
 * Events go to the region of the record's tile as everywhere else; the tile's event count in tile_info doubles as
 * the region's fill pointer (global atomics: a long read emits a handful of events per thousand operations). */
__global__ void __launch_bounds__(256) k_long_keys(const uint32_t *cigar_off, uint64_t n, uint32_t *n_ops, uint32_t *index,
                                                   TileInfo *tile_info) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c0 = __ldg(&cigar_off[i]);
    n_ops[i] = __ldg(&cigar_off[i + 1]) - c0;
    index[i] = (uint32_t)i;
    if (i % EX_TILE == 0) tile_info[i / EX_TILE] = TileInfo{0u, c0, 0u, 0u};
}

/* K1 sink of the long-read kernel: the walk keeps only the record index; what an event needs of the record (strand,
 * category, region of its tile) is fetched when a candidate comes up -- a few times per thousand operations -- instead
 * of living in registers through the walk.  Candidates arrive relative to the record's position. */
struct LazySink {
    const BatchView &b;
    const Tables &t;
    const ExtractParams &p;
    Event *out;
    TileInfo *tile_info;
    uint64_t r;
    uint32_t n_events;
    __device__ __forceinline__ EventSink open() const {
        const uint64_t tile = r / EX_TILE;
        Event *region = out + tile_event_base(__ldg(&b.cigar_off[tile * EX_TILE]));
        /* the tile's event count in tile_info doubles as the fill pointer of its region */
        return make_sink(t, p, region, &tile_info[tile].count, __ldg(&b.tid[r]), __ldg(&b.pos[r]), b.seq_base + r, __ldg(&b.flag[r]),
                         __ldg(&b.nh[r]), __ldg(&b.xs[r]));
    }
    __device__ __forceinline__ void operator()(const Candidate &rel) {
        EventSink sink = open();
        Candidate c = rel;
        c.anchor_start += sink.pos; c.intron_start += sink.pos; c.intron_end += sink.pos; c.anchor_end += sink.pos;
        sink(c);
        n_events += sink.n_events;
    }
};

/* the rest of a record whose relative positions outgrew 32 bits (a CIGAR spanning > 1.8 Gb: malformed input, kept exact) */
__device__ __forceinline__ uint32_t walk_wide_rest(const LeanState &lw, const uint32_t *cig, uint32_t k, uint32_t n, uint32_t max_indel,
                                                   const LazySink &lazy, uint32_t *n_events) {
    EventSink sink = lazy.open();
    WalkState w;
    lw.widen(sink.pos, w);
    uint32_t bad = 0u;
    for (; k < n && !bad; ++k) bad = walk_op_flat(w, __ldg(&cig[k]), max_indel, sink);
    if (!bad) walk_finish(w, sink);
    *n_events = sink.n_events;
    return bad;
}

struct Words8 {
    uint4 lo, hi;
};
__device__ __forceinline__ Words8 load_words8(const uint32_t *cigar, uint32_t k) { /* k a multiple of 8, k + 8 within the buffer */
    const uint4 *src = reinterpret_cast<const uint4 *>(cigar + k);
    Words8 v;
    v.lo = __ldg(src);
    v.hi = __ldg(src + 1);
    return v;
}

/* one live record, start to end, by the calling lane; returns the number of events written */
__device__ __forceinline__ uint32_t walk_record_lean(const BatchView &b, const ExtractParams &p, const Tables &t, Event *out,
                                                     TileInfo *tile_info, uint64_t r, uint32_t c0, uint32_t c1) {
    const uint32_t max_indel = p.max_anchor_indel_size;
    LazySink sink{b, t, p, out, tile_info, r, 0u};
    LeanState w;
    w.init();
    uint32_t bad = 0u;
    bool wide = false;
    /* the record's words [c0, c1) as: a few up to the first 32-byte boundary, whole groups of 8, a few more */
    const uint32_t head_end = min(c1, (c0 + 7u) & ~7u);
    const uint32_t body_end = head_end + ((c1 - head_end) & ~7u);
    uint32_t k = c0;
    Words8 cur;
    if (head_end < body_end) cur = load_words8(b.cigar, head_end);
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
        const uint32_t stop = part == 0 ? head_end : c1;
        if (part == 1 && w.tpos >= LEAN_TPOS_LIMIT) wide = true;
#pragma unroll 1
        for (; k < stop && !wide && !bad; ++k) lean_op(w, 0, __ldg(&b.cigar[k]), max_indel, sink, bad);
        if (part == 1) break;
        for (; k < body_end; k += 8u) {
            Words8 nxt = cur;
            if (k + 8u < body_end) nxt = load_words8(b.cigar, k + 8u);
            if (w.tpos >= LEAN_TPOS_LIMIT || bad) {
                wide = true;
                break;
            }
            lean_op(w, 0, cur.lo.x, max_indel, sink, bad);
            lean_op(w, 0, cur.lo.y, max_indel, sink, bad);
            lean_op(w, 0, cur.lo.z, max_indel, sink, bad);
            lean_op(w, 0, cur.lo.w, max_indel, sink, bad);
            lean_op(w, 0, cur.hi.x, max_indel, sink, bad);
            lean_op(w, 0, cur.hi.y, max_indel, sink, bad);
            lean_op(w, 0, cur.hi.z, max_indel, sink, bad);
            lean_op(w, 0, cur.hi.w, max_indel, sink, bad);
            cur = nxt;
        }
    }
    uint32_t n_events = sink.n_events;
    if (bad) { /* reported below */
    } else if (wide) {
        uint32_t more = 0u;
        bad = walk_wide_rest(w, b.cigar + c0, k - c0, c1 - c0, max_indel, sink, &more);
        n_events += more;
    } else {
        lean_finish(w, 0, sink);
        n_events = sink.n_events;
    }
    if (bad) atomicMin(t.err_sv, make_err_cigar(b.seq_base + r, bad & 15u)); /* the reference throws here */
    return n_events;
}

/* records with more than `max_ops` operations are left to the segmented kernels below */
__global__ void __launch_bounds__(256, 4) k_extract_sorted(BatchView b, uint64_t n_words, ExtractParams p, Tables t, Event *out,
                                                        TileInfo *tile_info, const uint32_t *order,
                                                        unsigned long long *next_chunk, uint32_t max_ops) {
    const int lane = threadIdx.x & 31;
    unsigned long long live_words = 0ull, events = 0ull;
    for (;;) {
        unsigned long long chunk = 0ull;
        if (lane == 0) chunk = atomicAdd(next_chunk, 1ull);
        chunk = __shfl_sync(0xFFFFFFFFu, chunk, 0);
        if (chunk * 32ull >= b.n) break;
        const uint64_t i = chunk * 32ull + (unsigned)lane;
        if (i >= b.n) continue;
        const uint64_t r = __ldg(&order[i]);
        const uint32_t c0 = __ldg(&b.cigar_off[r]), c1 = __ldg(&b.cigar_off[r + 1]);
        if (c1 - c0 > max_ops || !record_is_live(c1 - c0, __ldg(&b.tid[r]), __ldg(&b.flag[r]))) continue;
        live_words += c1 - c0;
        events += walk_record_lean(b, p, t, out, tile_info, r, c0, c1);
    }
    live_words = warp_sum(live_words);
    events = warp_sum(events);
    if (lane == 0) {
        if (live_words) atomicAdd(t.n_live, live_words);
        if (events) atomicAdd(t.n_events, events);
    }
}

/* ---- the longest records, each by many lanes (ipx_walk.cuh: "one long CIGAR walked by many lanes") ----
 * In the longest-first order the records with more than `min_ops` operations form a prefix; each is cut into
 * segments of about SEG_STRIDE operations.  seg_count[i] is the number of segments of the i-th record of that order
 * (0 for the shorter ones), seg_base its exclusive scan, seg_base[n] the number of segments. */
static const uint32_t SEG_STRIDE = 256;
struct SegCount {
    const uint32_t *n_ops_sorted;
    uint64_t n;
    uint32_t min_ops;
    __host__ __device__ __forceinline__ uint32_t operator()(uint64_t i) const {
        const uint32_t ops = i < n ? n_ops_sorted[i] : 0u;
        return ops > min_ops ? (ops + SEG_STRIDE - 1u) / SEG_STRIDE : 0u;
    }
};

/* the record (position in the longest-first order) that owns segment g: the last i with seg_base[i] <= g */
__device__ __forceinline__ uint32_t seg_owner(const uint32_t *seg_base, uint32_t n, uint32_t g) {
    uint32_t lo = 0u, hi = n; /* seg_base[lo] <= g < seg_base[hi] */
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (seg_base[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

/* pass A: what each segment does to the automaton's registers */
__global__ void __launch_bounds__(256) k_seg_summaries(BatchView b, ExtractParams p, const uint32_t *order, const uint32_t *seg_base,
                                                       uint32_t *seg_first, SegSummary *summary) {
    const uint32_t n = (uint32_t)b.n, n_seg = seg_base[n];
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_seg; g += stride) {
        const uint32_t i = seg_owner(seg_base, n, g), j = g - seg_base[i];
        const uint64_t r = __ldg(&order[i]);
        const uint32_t c0 = __ldg(&b.cigar_off[r]), n_ops = __ldg(&b.cigar_off[r + 1]) - c0;
        const uint32_t *cig = b.cigar + c0;
        auto word_at = [&](uint32_t k) { return __ldg(&cig[k]); };
        const uint32_t first = seg_start(word_at, n_ops, j, SEG_STRIDE);
        const uint32_t last = (j + 1u) * SEG_STRIDE >= n_ops ? n_ops : seg_start(word_at, n_ops, j + 1u, SEG_STRIDE);
        SegWalk w;
        w.init(j == 0u);
        for (uint32_t k = first; k < last; ++k) seg_op(w, __ldg(&cig[k]), p.max_anchor_indel_size);
        w.finish();
        seg_first[g] = first;
        summary[g] = w.s;
    }
}

/* stitch: a thread per long record applies its segments' summaries in order; entry[g] = the registers segment g is
 * entered with, mode[g] = 0 nothing to do / 1 walk the segment / 3 walk it and finish the record (its last segment
 * that has operations) / 2 walk the whole record in one piece (first segment only) */
__global__ void __launch_bounds__(128) k_seg_stitch(BatchView b, Tables t, const uint32_t *order, const uint32_t *n_ops_sorted,
                                                    const uint32_t *seg_base, const uint32_t *seg_first, const SegSummary *summary,
                                                    LeanState *entry, uint8_t *mode, uint32_t min_ops) {
    const uint32_t n = (uint32_t)b.n;
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long live_words = 0ull;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t n_ops = n_ops_sorted[i];
        if (n_ops <= min_ops) break; /* longest first: nothing but shorter records from here on */
        const uint32_t g0 = seg_base[i], m = seg_base[i + 1] - g0;
        const uint64_t r = __ldg(&order[i]);
        const bool live = record_is_live(n_ops, __ldg(&b.tid[r]), __ldg(&b.flag[r]));
        if (live) live_words += n_ops;
        LeanState e;
        e.init();
        bool whole = false;
        uint32_t last_walked = 0u;
        for (uint32_t j = 0; j < m && live && !whole; ++j) {
            const uint32_t g = g0 + j;
            const uint32_t last = j + 1u < m ? seg_first[g + 1u] : n_ops;
            entry[g] = e;
            mode[g] = seg_first[g] < last ? 1 : 0;
            if (seg_first[g] < last) {
                LeanState x;
                whole = !seg_apply(summary[g], e, x);
                e = x;
                last_walked = g;
            }
        }
        if (!live || whole) {
            for (uint32_t j = 0; j < m; ++j) mode[g0 + j] = (j == 0u && live) ? 2 : 0;
        } else {
            mode[last_walked] = 3; /* the first segment always has operations */
        }
    }
    if (live_words) atomicAdd(t.n_live, live_words);
}

/* pass C: every segment again, now from its concrete entry registers, emitting */
__global__ void __launch_bounds__(256) k_seg_emit(BatchView b, ExtractParams p, Tables t, Event *out, TileInfo *tile_info,
                                                  const uint32_t *order, const uint32_t *seg_base, const uint32_t *seg_first,
                                                  const LeanState *entry, const uint8_t *mode) {
    const uint32_t n = (uint32_t)b.n, n_seg = seg_base[n];
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long events = 0ull;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_seg; g += stride) {
        const uint32_t md = mode[g];
        if (md == 0u) continue;
        const uint32_t i = seg_owner(seg_base, n, g), j = g - seg_base[i];
        const uint64_t r = __ldg(&order[i]);
        const uint32_t c0 = __ldg(&b.cigar_off[r]), c1 = __ldg(&b.cigar_off[r + 1]), n_ops = c1 - c0;
        if (md == 2u) {
            events += walk_record_lean(b, p, t, out, tile_info, r, c0, c1);
            continue;
        }
        const uint32_t first = seg_first[g], last = g + 1u == seg_base[i + 1] ? n_ops : seg_first[g + 1u];
        const uint32_t *cig = b.cigar + c0;
        LazySink sink{b, t, p, out, tile_info, r, 0u};
        LeanState w = entry[g];
        uint32_t bad = 0u;
        for (uint32_t k = first; k < last; ++k) lean_op(w, 0, __ldg(&cig[k]), p.max_anchor_indel_size, sink, bad);
        if (md == 3u) lean_finish(w, 0, sink);
        if (bad) atomicMin(t.err_sv, make_err_cigar(b.seq_base + r, bad & 15u)); /* (pass A sends such records to mode 2) */
        events += sink.n_events;
        (void)j;
    }
    events = warp_sum(events);
    if ((threadIdx.x & 31) == 0 && events) atomicAdd(t.n_events, events);
}

/* ---- pipelined K1 ---- */
/* One block = a producer warp and a pool of worker warps; nothing is synchronised block-wide after start-up.
 *   producer (one lane)  TMA bulk copies of the next tile's columns and CIGAR words into a free stage of the ring;
 *                        when a stage comes back it publishes the tile's event count and resets the stage
 *   workers              walk the ring in tile order.  A tile offers two kinds of work through dispensers in its stage:
 *                        "rounds" (process_alignment's skip test for 128 records and compaction of the live ones into
 *                        the stage's lists) and, once all rounds are done, "chunks" (walk 32 listed records).  In
 *                        iteration i a worker first takes rounds of tile i+1 if its data has landed, then makes sure
 *                        tile i is classified (helping with its rounds if any are left), then takes chunks of tile i.
 *                        Classification of a tile therefore overlaps the walk of the tile before it, and a warp busy
 *                        with a long walk simply takes fewer units. */
static const int EX_STAGES = 4;
static const int EX_WORKER_WARPS = 15; /* default; 9 / 11 / 13 are instantiated for tuning (IPX_K1_WORKERS) */
static const int EX_LOOKAHEAD = 1; /* workers classify up to this many tiles ahead of the one they walk */
static const int EX_ROUNDS = EX_TILE / 128;
static const int EX_CIG_CAP = 2048; /* staged CIGAR words per tile (average short-read tile: ~1500); words beyond are
                                       read from global memory */

struct __align__(128) ExStage {
    /* filled by the TMA bulk copies */
    uint32_t off[EX_TILE + 4]; /* cigar_off of the tile's records; [EX_TILE] is stored by the producer */
    int32_t tid[EX_TILE];
    int32_t pos[EX_TILE];
    uint32_t cig[EX_CIG_CAP];
    uint16_t flag[EX_TILE];
    uint8_t nh[EX_TILE];
    uint8_t xs[EX_TILE];
    /* written by the block itself */
    uint16_t list[EX_TILE]; /* live records: short CIGARs (<= 3 operations) from the front, longer ones from the back */
    uint32_t n_lists;       /* number of short | number of long << 16 */
    uint32_t next_round, rounds_done; /* classification rounds handed out / finished */
    uint32_t next_short, next_long;   /* chunk dispensers */
    uint32_t region_fill;   /* events written to the tile's region */
    uint32_t tile;          /* tile index */
    uint32_t a0, a_end;     /* staged CIGAR words are [a0, a_end) of the batch's CIGAR column */
    uint32_t pad[3];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
/* has the phase with the given parity completed?  (the hint lets the hardware park the thread for a while) */
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return done != 0u;
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0u;
}
/* try_wait parks the thread on the barrier for a while (NANOSLEEP.SYNCS) but comes back early often enough that a bare
 * retry loop burns a seventh of the kernel's issue slots; a short sleep between tries gives them to the working warps */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t sleep_ns) {
    while (!mbar_try(bar, parity, 1000000u)) {
        if (sleep_ns) __nanosleep(sleep_ns);
    }
}
/* TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

/* a CIGAR word of the batch: from the stage when it was staged, else from global memory */
__device__ __forceinline__ uint32_t cigar_word(const ExStage &S, const uint32_t *cigar, uint32_t gi) {
    return gi < S.a_end ? S.cig[gi - S.a0] : __ldg(&cigar[gi]);
}

/* walk one listed record with the full automaton */
__device__ __forceinline__ void walk_record(const BatchView &b, const ExtractParams &p, const Tables &t, ExStage &S, Event *region,
                                            uint32_t rl, unsigned long long &events) {
    const uint32_t c0 = S.off[rl], n_cigar = S.off[rl + 1] - c0;
    const int64_t pos = S.pos[rl];
    const uint64_t seq = b.seq_base + (uint64_t)S.tile * EX_TILE + rl;
    EventSink sink = make_sink(t, p, region, &S.region_fill, S.tid[rl], pos, seq, S.flag[rl], S.nh[rl], S.xs[rl]);
    WalkState w;
    w.init(pos);
    uint32_t bad = 0u;
    for (uint32_t k = 0; k < n_cigar; ++k) {
        bad = walk_op_flat(w, cigar_word(S, b.cigar, c0 + k), p.max_anchor_indel_size, sink);
        if (bad) break;
    }
    if (bad) {
        atomicMin(t.err_sv, make_err_cigar(seq, bad & 15u));
    } else {
        walk_finish(w, sink);
    }
    events += sink.n_events;
}

/* one classification round: process_alignment's skip test (junctions_extractor.cc:536) for records
 * [128 r, 128 r + 128) of the stage's tile, four consecutive records per lane, and compaction of the live ones */
__device__ __forceinline__ void classify_round(ExStage &S, uint32_t r, int lane, LongRec *long_recs, int walk_long_here,
                                               unsigned long long &live_words) {
    const uint32_t q = r * 32u + lane; /* records 4q .. 4q+3 */
    const uint4 o = reinterpret_cast<const uint4 *>(S.off)[q];
    const uint32_t o4 = S.off[4 * q + 4];
    const int4 td = reinterpret_cast<const int4 *>(S.tid)[q];
    const uint2 f2 = reinterpret_cast<const uint2 *>(S.flag)[q];
    const uint32_t n[4] = {o.y - o.x, o.z - o.y, o.w - o.z, o4 - o.w};
    const int32_t tdv[4] = {td.x, td.y, td.z, td.w};
    const uint32_t fl[4] = {f2.x & 0xFFFFu, f2.x >> 16, f2.y & 0xFFFFu, f2.y >> 16};
    uint32_t kind[4], mine = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        kind[j] = 0u;
        if (record_is_live(n[j], tdv[j], fl[j])) {
            live_words += n[j];
            kind[j] = n[j] <= 3u ? 1u : 2u;
            mine += n[j] <= 3u ? 1u : 0x10000u;
        }
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
    }
    uint32_t base = 0u;
    if (lane == 31 && incl) base = atomicAdd(&S.n_lists, incl);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    const uint32_t pre = base + incl - mine;
    uint32_t ps = pre & 0xFFFFu, pl = pre >> 16;
    LongRec *long_region = long_recs + (((uint64_t)S.off[0] + 3ull) >> 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t rl = 4 * q + j;
        if (kind[j] == 1u) {
            S.list[ps++] = (uint16_t)rl;
        } else if (kind[j] == 2u && walk_long_here) {
            S.list[EX_TILE - 1 - (pl++)] = (uint16_t)rl;
        } else if (kind[j] == 2u) {
            /* hand the record to k_extract_long */
            const uint32_t c0r = j == 0 ? o.x : (j == 1 ? o.y : (j == 2 ? o.z : o.w));
            uint4 *dst = reinterpret_cast<uint4 *>(long_region + (pl++));
            dst[0] = make_uint4((uint32_t)S.pos[rl], (uint32_t)tdv[j], c0r, n[j]);
            dst[1] = make_uint4(S.tile * EX_TILE + rl, fl[j] | ((uint32_t)S.nh[rl] << 16) | ((uint32_t)S.xs[rl] << 24), 0u, 0u);
        }
    }
}

/* take classification rounds of a stage until none is left */
__device__ __forceinline__ void take_rounds(ExStage &S, int stage, uint64_t *classified, int lane, LongRec *long_recs,
                                            int walk_long_here, unsigned long long &live_words) {
    for (;;) {
        uint32_t r = 0u;
        /* (the plain read spares the atomic when nothing is left; the decision stays lane 0's, so it is warp-uniform) */
        if (lane == 0) r = *((volatile uint32_t *)&S.next_round) >= (uint32_t)EX_ROUNDS ? 0xFFFFFFFFu : atomicAdd(&S.next_round, 1u);
        r = __shfl_sync(0xFFFFFFFFu, r, 0);
        if (r >= (uint32_t)EX_ROUNDS) break;
        classify_round(S, r, lane, long_recs, walk_long_here, live_words);
        __syncwarp();
        if (lane == 0) mbar_arrive(&classified[stage]); /* release: the lists are visible to whoever sees the phase complete */
    }
}

/* the short-CIGAR chunk: "aligned, N, aligned" is the only shape of two or three operations that emits a junction
 * (ReadJunctionExtractor, junctions_extractor.cc:208-254,278-282) */
__device__ __forceinline__ void walk_short(const BatchView &b, const ExtractParams &p, const Tables &t, ExStage &S, Event *region,
                                           uint32_t rl, unsigned long long &events) {
    const uint32_t c0 = S.off[rl], n_cigar = S.off[rl + 1] - c0;
    const uint32_t w0 = cigar_word(S, b.cigar, c0), w1 = cigar_word(S, b.cigar, c0 + 1);
    const uint32_t w2 = n_cigar == 3u ? cigar_word(S, b.cigar, c0 + 2) : 0u;
    const uint32_t o0 = w0 & 15u, o1 = w1 & 15u, o2 = w2 & 15u;
    const uint32_t known = 0x1BFu, aligned = 0x181u; /* MIDNSH=X ; M=X */
    if (!((known >> o0) & (known >> o1) & (known >> o2) & 1u)) {
        walk_record(b, p, t, S, region, rl, events); /* reports the unknown operation exactly */
    } else if (n_cigar == 3u && o1 == OP_N && ((aligned >> o0) & (aligned >> o2) & 1u)) {
        /* the one candidate: anchor_start = pos, intron = [pos + l0, + l1), anchor_end = intron_end + l2, mismatches =
         * the X lengths, no indels.  Operation lengths are below 2^28 and pos is a BAM int32, so 32-bit arithmetic
         * gives what the reference's 64-bit arithmetic gives wherever the result is usable (a start outside
         * [0, 2^31) is reported). */
        const uint32_t l0 = w0 >> 4, l1 = w1 >> 4, l2 = w2 >> 4;
        const uint32_t flag = S.flag[rl], nh = S.nh[rl];
        const uint32_t lbad = o0 == OP_X ? l0 : 0u, rbad = o2 == OP_X ? l2 : 0u;
        /* junction_qc (junctions_extractor.cc:384-401) */
        if (read_may_emit(p, flag, nh) && l1 >= p.min_intron_length && l1 <= p.max_intron_length &&
            l0 - lbad >= p.min_anchor_length && l2 - rbad >= p.min_anchor_length) {
            events++;
            const uint32_t pos = (uint32_t)S.pos[rl];
            const uint32_t i0 = pos + l0;
            if (i0 > 0x7FFFFFFFu) {
                atomicExch(t.overflow, 3u);
            } else {
                const char strand = read_strand(p.strandness, flag, S.xs[rl]);
                const unsigned long long key = ((unsigned long long)(uint32_t)S.tid[rl] << t.kl.tid_shift) |
                                               ((unsigned long long)i0 << t.kl.start_shift) | l1;
                const unsigned long long sv = make_sv(b.seq_base + (uint64_t)S.tile * EX_TILE + rl, 0u, strand);
                const uint32_t off_cat = (l0 & 0xFFFFu) | (read_category(flag, nh) << 16) | ((uint32_t)strand_class(strand) << 18);
                const uint32_t idx = atomicAdd(&S.region_fill, 1u);
                uint4 *dst = reinterpret_cast<uint4 *>(region + idx);
                __stcs(dst, make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)sv, (uint32_t)(sv >> 32)));
                __stcs(dst + 1, make_uint4(pos, i0 + l1 + l2, off_cat, 0u));
            }
        }
    }
}

template <int LOOKAHEAD, int WORKERS>
__global__ void __launch_bounds__((WORKERS + 1) * 32, 2) k_extract_tma(BatchView b, uint32_t n_cigar_words, ExtractParams p, Tables t,
                                                               Event *out, TileInfo *tile_info, LongRec *long_recs,
                                                               uint32_t n_full_tiles, int walk_long_here, uint32_t sleep_ns) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ExStage *stages = reinterpret_cast<ExStage *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + EX_STAGES * sizeof(ExStage)); /* TMA bytes have landed */
    uint64_t *empty = full + EX_STAGES;                                                    /* every worker has left */
    uint64_t *classified = empty + EX_STAGES;                                              /* all rounds are done */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < EX_STAGES; ++s) {
            mbar_init(&full[s], 1u);
            mbar_init(&empty[s], (uint32_t)WORKERS);
            mbar_init(&classified[s], (uint32_t)EX_ROUNDS);
            stages[s].n_lists = stages[s].next_round = stages[s].rounds_done = 0u;
            stages[s].next_short = stages[s].next_long = stages[s].region_fill = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint32_t first = blockIdx.x, stride = gridDim.x;
    const uint32_t my_tiles = first < n_full_tiles ? (n_full_tiles - first + stride - 1) / stride : 0u;

    if (warp == WORKERS) {
        /* ---------------- producer: one lane feeds the ring ---------------- */
        if (lane != 0) return;
        uint32_t nc0 = 0u, nc1 = 0u;
        if (my_tiles) {
            nc0 = __ldg(&b.cigar_off[(uint64_t)first * EX_TILE]);
            nc1 = __ldg(&b.cigar_off[(uint64_t)first * EX_TILE + EX_TILE]);
        }
        for (uint32_t it = 0; it < my_tiles + EX_STAGES; ++it) {
            const int s = (int)(it % EX_STAGES);
            ExStage &S = stages[s];
            if (it >= (uint32_t)EX_STAGES) {
                /* the tile that occupied this stage has been left by every worker */
                mbar_wait(&empty[s], ((it / EX_STAGES) & 1u) ^ 1u, sleep_ns ? 4 * sleep_ns : 0u);
                tile_info[S.tile] = TileInfo{S.region_fill, S.off[0], walk_long_here ? 0u : S.n_lists >> 16, 0u};
                S.n_lists = S.next_round = S.rounds_done = S.next_short = S.next_long = S.region_fill = 0u;
            }
            if (it < my_tiles) {
                const uint32_t tile = first + it * stride;
                const uint64_t r0 = (uint64_t)tile * EX_TILE;
                const uint32_t c0 = nc0, c1 = nc1;
                if (it + 1 < my_tiles) {
                    /* boundaries of the next tile: in flight while this one is issued and the stage wait above runs */
                    nc0 = __ldg(&b.cigar_off[r0 + (uint64_t)stride * EX_TILE]);
                    nc1 = __ldg(&b.cigar_off[r0 + (uint64_t)stride * EX_TILE + EX_TILE]);
                }
                const uint32_t a0 = c0 & ~3u;
                uint32_t a_end = (c1 + 3u) & ~3u;
                if (a_end - a0 > (uint32_t)EX_CIG_CAP) a_end = a0 + (uint32_t)EX_CIG_CAP;
                if (a_end > n_cigar_words) a_end = a0 + ((n_cigar_words - a0) & ~3u);
                const uint32_t cig_bytes = (a_end - a0) * 4u;
                S.tile = tile;
                S.a0 = a0;
                S.a_end = a_end;
                S.off[EX_TILE] = c1;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(EX_TILE * (4 + 4 + 4 + 2 + 1 + 1)) + cig_bytes);
                bulk_g2s(S.off, b.cigar_off + r0, EX_TILE * 4, &full[s]);
                bulk_g2s(S.tid, b.tid + r0, EX_TILE * 4, &full[s]);
                bulk_g2s(S.flag, b.flag + r0, EX_TILE * 2, &full[s]);
                if (cig_bytes) bulk_g2s(S.cig, b.cigar + a0, cig_bytes, &full[s]);
                bulk_g2s(S.pos, b.pos + r0, EX_TILE * 4, &full[s]);
                bulk_g2s(S.nh, b.nh + r0, EX_TILE, &full[s]);
                bulk_g2s(S.xs, b.xs + r0, EX_TILE, &full[s]);
            }
        }
        return;
    }

    /* ---------------- workers ---------------- */
    unsigned long long live_words = 0ull, events = 0ull;
    for (uint32_t it = 0; it < my_tiles; ++it) {
        const int s = (int)(it % EX_STAGES);
        ExStage &S = stages[s];
        /* (1) rounds of the tiles ahead whose data is already here, nearest first */
#pragma unroll
        for (uint32_t d = 1; d <= (uint32_t)LOOKAHEAD; ++d) {
            if (it + d < my_tiles) {
                const int s1 = (int)((it + d) % EX_STAGES);
                if (mbar_test(&full[s1], ((it + d) / EX_STAGES) & 1u))
                    take_rounds(stages[s1], s1, classified, lane, long_recs, walk_long_here, live_words);
            }
        }
        /* (2) this tile must be classified before it can be walked: help with what is left, then wait for the rest */
        mbar_wait(&full[s], (it / EX_STAGES) & 1u, sleep_ns);
        take_rounds(S, s, classified, lane, long_recs, walk_long_here, live_words);
        mbar_wait(&classified[s], (it / EX_STAGES) & 1u, sleep_ns);
        /* (3) chunks of this tile: 32 listed records at a time, longest CIGARs first */
        const uint32_t n_lists = *((volatile uint32_t *)&S.n_lists);
        const uint32_t n_short = n_lists & 0xFFFFu, n_long = walk_long_here ? n_lists >> 16 : 0u;
        Event *region = out + tile_event_base(S.off[0]);
        while (n_long) {
            uint32_t c = 0u;
            if (lane == 0) c = *((volatile uint32_t *)&S.next_long) >= n_long ? 0xFFFFFFFFu : atomicAdd(&S.next_long, 32u);
            c = __shfl_sync(0xFFFFFFFFu, c, 0);
            if (c >= n_long) break;
            if (c + lane < n_long) walk_record(b, p, t, S, region, S.list[EX_TILE - 1 - (c + lane)], events);
        }
        for (;;) {
            uint32_t c = 0u;
            if (lane == 0) c = *((volatile uint32_t *)&S.next_short) >= n_short ? 0xFFFFFFFFu : atomicAdd(&S.next_short, 32u);
            c = __shfl_sync(0xFFFFFFFFu, c, 0);
            if (c >= n_short) break;
            if (c + lane < n_short) walk_short(b, p, t, S, region, S.list[c + lane], events);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    live_words = warp_sum(live_words);
    events = warp_sum(events);
    if (lane == 0) {
        if (live_words) atomicAdd(t.n_live, live_words);
        if (events) atomicAdd(t.n_events, events);
    }
}

/* ---- K1 for the records the pipelined kernel handed over: a thread per record, CIGAR words from global memory ---- */
__global__ void __launch_bounds__(256) k_extract_long(BatchView b, ExtractParams p, Tables t, Event *out, TileInfo *tile_info,
                                                      const LongRec *long_recs, uint64_t n_tiles) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long events = 0ull;
    for (uint64_t tile = warp0; tile < n_tiles; tile += n_warps) {
        const uint4 info = __ldg(reinterpret_cast<const uint4 *>(tile_info + tile));
        const uint32_t n_long = info.z;
        if (n_long == 0u) continue;
        Event *region = out + tile_event_base(info.y);
        const LongRec *recs = long_recs + (((uint64_t)info.y + 3ull) >> 2);
        for (uint32_t i = lane; i < n_long; i += 32u) {
            const uint4 r0 = __ldg(reinterpret_cast<const uint4 *>(recs + i));
            const uint4 r1 = __ldg(reinterpret_cast<const uint4 *>(recs + i) + 1);
            const int64_t pos = (int32_t)r0.x;
            const uint32_t c0 = r0.z, n_cigar = r0.w;
            const uint64_t seq = b.seq_base + r1.x;
            /* the tile's event count doubles as the fill pointer of its region */
            EventSink sink = make_sink(t, p, region, &tile_info[tile].count, (int32_t)r0.y, pos, seq, r1.y & 0xFFFFu,
                                       (r1.y >> 16) & 0xFFu, r1.y >> 24);
            WalkState w;
            w.init(pos);
            uint32_t bad = 0u;
            for (uint32_t k = 0; k < n_cigar; ++k) {
                bad = walk_op_flat(w, __ldg(&b.cigar[c0 + k]), p.max_anchor_indel_size, sink);
                if (bad) break;
            }
            if (bad) {
                atomicMin(t.err_sv, make_err_cigar(seq, bad & 15u));
            } else {
                walk_finish(w, sink);
            }
            events += sink.n_events;
        }
    }
    events = warp_sum(events);
    if (lane == 0 && events) atomicAdd(t.n_events, events);
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

cudaError_t launch_extract(const BatchView &b, uint64_t n_cigar, const ExtractParams &p, const Tables &t, Event *out,
                           TileInfo *tile_info, LongRec *long_recs, unsigned long long *next_record, int sm_count, cudaStream_t st,
                           int *launches) {
    if (launches) *launches = 0;
    if (b.n == 0) return cudaSuccess;
    const uint64_t n_tiles = tiles_of(b.n), n_full = b.n / EX_TILE;
    /* the pipelined kernel takes whole tiles of batches whose columns can be bulk-copied and whose reads are short
     * enough for a tile's CIGAR words to fit its stage */
    static const bool force_simple = getenv("IPX_K1_SIMPLE") != nullptr;
    const bool pipelined = !force_simple && n_full > 0 && n_cigar <= 8 * b.n && aligned16(b.tid) && aligned16(b.pos) &&
                           aligned16(b.flag) && aligned16(b.nh) && aligned16(b.xs) && aligned16(b.cigar_off) && aligned16(b.cigar);
    uint64_t first_simple = 0;
    if (pipelined) {
        const size_t smem = EX_STAGES * sizeof(ExStage) + 3 * EX_STAGES * sizeof(uint64_t);
        static const int walk_long_here = getenv("IPX_K1_SPLIT_LONG") ? 0 : 1;
        static const int workers_env = getenv("IPX_K1_WORKERS") ? atoi(getenv("IPX_K1_WORKERS")) : EX_WORKER_WARPS;
        typedef void (*Kernel)(BatchView, uint32_t, ExtractParams, Tables, Event *, TileInfo *, LongRec *, uint32_t, int, uint32_t);
        static const uint32_t sleep_ns = getenv("IPX_K1_SLEEP") ? (uint32_t)atoi(getenv("IPX_K1_SLEEP")) : 0u;
        /* measured on B200 (100 M short reads): look-ahead 1 / 2 / 3 tiles: 0.74 / 0.77 / 0.81 ms; named barriers 1.05 ms */
        Kernel kernel = k_extract_tma<EX_LOOKAHEAD, EX_WORKER_WARPS>;
        int workers = EX_WORKER_WARPS;
        if (workers_env == 9) { kernel = k_extract_tma<EX_LOOKAHEAD, 9>; workers = 9; }
        if (workers_env == 11) { kernel = k_extract_tma<EX_LOOKAHEAD, 11>; workers = 11; }
        if (workers_env == 13) { kernel = k_extract_tma<EX_LOOKAHEAD, 13>; workers = 13; }
        if (workers_env == 17) { kernel = k_extract_tma<EX_LOOKAHEAD, 17>; workers = 17; }
        if (workers_env == 19) { kernel = k_extract_tma<EX_LOOKAHEAD, 19>; workers = 19; }
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured = true;
        }
        const uint64_t max_blocks = (uint64_t)sm_count * 2;
        const int blocks = (int)(n_full < max_blocks ? n_full : max_blocks);
        kernel<<<blocks, (workers + 1) * 32, smem, st>>>(b, (uint32_t)n_cigar, p, t, out, tile_info, long_recs, (uint32_t)n_full,
                                                        walk_long_here, sleep_ns);
        if (launches) (*launches)++;
        if (!walk_long_here) {
            const uint64_t want = (n_full + 7) / 8;
            const uint64_t max_long = (uint64_t)sm_count * 8;
            k_extract_long<<<(int)(want < max_long ? want : max_long), 256, 0, st>>>(b, p, t, out, tile_info, long_recs, n_full);
            if (launches) (*launches)++;
        }
        first_simple = n_full;
    }
    /* long reads (more than 8 operations per record on average): longest first, lean automaton; the longest of them
     * segment by segment */
    if (first_simple == 0 && !force_simple && n_cigar > 8 * b.n && aligned16(b.cigar) && next_record) {
        /* Records longer than this are walked segment by segment.  That is twice the work per operation, and on
         * log-normal read lengths (longest read 12x the mean) it measured slower at every threshold (B200, 1 M reads:
         * 3.6 / 2.9 / 2.6 / 2.6 / 2.4 ms for 1536 / 2048 / 3072 / 4096 / 6144 against 2.3 ms without): the kernel is
         * bound by throughput, not by its longest record.  So the default only catches outliers that would be -- a
         * single lane needs ~0.3 us per operation.  (Read per call: the tests lower it.) */
        const char *seg_env = getenv("IPX_K1_SEG_MIN");
        const uint32_t seg_min_ops = seg_env ? (uint32_t)atoi(seg_env) : 16384u;
        const bool segmented = seg_min_ops != 0u;
        const int n = (int)b.n;
        size_t sort_bytes = 0, scan_bytes = 0;
        cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, n, 0, 32, st);
        if (e != cudaSuccess) return e;
        auto seg_counts = thrust::make_transform_iterator(thrust::make_counting_iterator<uint64_t>(0), SegCount{nullptr, b.n, seg_min_ops});
        e = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, seg_counts, (uint32_t *)nullptr, n + 1, st);
        if (e != cudaSuccess) return e;
        const size_t tmp_bytes = ((sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 255) & ~(size_t)255;
        const size_t col = ((size_t)(b.n + 1) * sizeof(uint32_t) + 255) & ~(size_t)255;
        /* at most one segment per SEG_STRIDE operations plus one per segmented record */
        const size_t max_seg = segmented ? (size_t)(n_cigar / SEG_STRIDE + n_cigar / (seg_min_ops ? seg_min_ops : 1u) + 2) : 1;
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t seg_bytes = up(max_seg * sizeof(uint32_t)) + up(max_seg * sizeof(SegSummary)) + up(max_seg * sizeof(LeanState)) + up(max_seg);
        char *scratch = nullptr;
        e = cudaMallocAsync(&scratch, 5 * col + tmp_bytes + seg_bytes, st);
        if (e != cudaSuccess) return e;
        uint32_t *n_ops = reinterpret_cast<uint32_t *>(scratch), *n_ops_sorted = reinterpret_cast<uint32_t *>(scratch + col);
        uint32_t *index = reinterpret_cast<uint32_t *>(scratch + 2 * col), *order = reinterpret_cast<uint32_t *>(scratch + 3 * col);
        uint32_t *seg_base = reinterpret_cast<uint32_t *>(scratch + 4 * col);
        char *tmp = scratch + 5 * col, *q = tmp + tmp_bytes;
        uint32_t *seg_first = reinterpret_cast<uint32_t *>(q); q += up(max_seg * sizeof(uint32_t));
        SegSummary *summary = reinterpret_cast<SegSummary *>(q); q += up(max_seg * sizeof(SegSummary));
        LeanState *entry = reinterpret_cast<LeanState *>(q); q += up(max_seg * sizeof(LeanState));
        uint8_t *mode = reinterpret_cast<uint8_t *>(q);
        k_long_keys<<<(unsigned)((b.n + 255) / 256), 256, 0, st>>>(b.cigar_off, b.n, n_ops, index, tile_info);
        size_t sb = sort_bytes;
        e = cub::DeviceRadixSort::SortPairsDescending(tmp, sb, n_ops, n_ops_sorted, index, order, n, 0, 32, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(next_record, 0, sizeof(unsigned long long), st);
        if (e == cudaSuccess && segmented) {
            auto counts = thrust::make_transform_iterator(thrust::make_counting_iterator<uint64_t>(0), SegCount{n_ops_sorted, b.n, seg_min_ops});
            size_t cb = scan_bytes;
            e = cub::DeviceScan::ExclusiveSum(tmp, cb, counts, seg_base, n + 1, st);
            if (e == cudaSuccess) {
                k_seg_summaries<<<sm_count * 8, 256, 0, st>>>(b, p, order, seg_base, seg_first, summary);
                k_seg_stitch<<<sm_count * 2, 128, 0, st>>>(b, t, order, n_ops_sorted, seg_base, seg_first, summary, entry, mode, seg_min_ops);
                k_seg_emit<<<sm_count * 8, 256, 0, st>>>(b, p, t, out, tile_info, order, seg_base, seg_first, entry, mode);
                if (launches) (*launches) += 4;
                e = cudaGetLastError();
            }
        }
        if (e == cudaSuccess) {
            const uint64_t chunks = (b.n + 31) / 32, want = (chunks + 7) / 8, max_blocks = (uint64_t)sm_count * 8;
            k_extract_sorted<<<(int)(want < max_blocks ? want : max_blocks), 256, 0, st>>>(b, n_cigar, p, t, out, tile_info, order,
                                                                                         next_record, segmented ? seg_min_ops : 0xFFFFFFFFu);
            if (e == cudaSuccess) e = cudaGetLastError();
        }
        cudaFreeAsync(scratch, st);
        if (launches) (*launches) += 3; /* keys, extract, and the sort counted as one */
        return e;
    } else if (first_simple < n_tiles) {
        const uint64_t todo = n_tiles - first_simple;
        const uint64_t max_blocks = (uint64_t)sm_count * 8;
        const int blocks = (int)(todo < max_blocks ? todo : max_blocks);
        k_extract_simple<<<blocks, EXS_THREADS, 0, st>>>(b, p, t, out, tile_info, first_simple, n_tiles);
        if (launches) (*launches)++;
    }
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ K2: keyed aggregation */
__device__ __forceinline__ Event load_event(const Event *ev, uint64_t i, bool valid, int lane) {
    Event e;
    if (valid) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ev + i);
        const uint4 lo = __ldcs(src), hi = __ldcs(src + 1); /* streamed once: do not displace the tables in L2 */
        e.key = ((unsigned long long)lo.y << 32) | lo.x;
        e.sv = ((unsigned long long)lo.w << 32) | lo.z;
        e.a0 = hi.x; e.a1 = hi.y; e.off_cat = hi.z;
    } else {
        /* distinct keys that no real event has (bit 63 is never set in a key), so idle lanes group with nobody */
        e.key = EMPTY64 - 2 - (unsigned)lane; e.sv = EMPTY64; e.a0 = 0xFFFFFFFFu; e.a1 = 0u; e.off_cat = 0u;
    }
    e.pad = 0u;
    return e;
}

/* probe for `key` and, in the same round trip, read what the fold wants to know about its junction: the anchor
 * extremes, the first-seen value of strand class `cls` and the creation threshold.  The values may be stale by the
 * time they are used; they only ever tighten, so a stale value can cause a redundant atomic, never a missed one. */
struct Probe {
    uint64_t slot; /* cap_mask + 1 = absent (lookup) or table full (insert) */
    uint32_t a0, a1;
    unsigned long long first, T;
};

template <bool INSERT>
__device__ __forceinline__ Probe probe_junction(const Tables &t, unsigned long long key, int cls) {
    Probe r;
    uint64_t h = mix64(key) & t.cap_mask;
    for (uint32_t probe = 0; probe < MAX_PROBE; ++probe) {
        const JStat *s = &t.stat[h];
        unsigned long long k = *((volatile unsigned long long *)&t.keys[h]);
        uint2 a;
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "l"(&s->a0));
        r.first = *((volatile unsigned long long *)&s->first[cls]);
        r.T = *((volatile unsigned long long *)&s->T);
        r.a0 = a.x;
        r.a1 = a.y;
        r.slot = h;
        if (k == key) return r;
        if (k == EMPTY64) {
            if (!INSERT) break;
            const unsigned long long old = atomicCAS(&t.keys[h], EMPTY64, key);
            if (old == EMPTY64) {
                const unsigned long long id = atomicAdd(t.n_keys, 1ull);
                t.slot_list[id] = (uint32_t)h;
                t.slot_id[h] = (uint32_t)id;
                return r; /* the record of a free slot holds the reset values */
            }
            if (old == key) {
                r.a0 = 0xFFFFFFFFu; r.a1 = 0u; r.first = EMPTY64; r.T = EMPTY64; /* not re-read: treat as unknown */
                return r;
            }
        }
        h = (h + 1) & t.cap_mask;
    }
    if (INSERT) atomicExch(t.overflow, 1u);
    r.slot = t.cap_mask + 1;
    return r;
}

/* K2.  A warp takes the event regions of whole tiles; 32 events at a time (neighbouring reads, so mostly the same
 * few junctions) are grouped by key with match.any and the lowest lane of each group probes the junction table once
 * for the whole group.  Read counts are added once per group; anchor extremes and first-seen values are compared
 * against what the probe saw and only improvements are sent as atomics (almost none once a junction is warm); each
 * lane counts its own read offset with one fire-and-forget atomic into the junction's dense offset row (or the
 * offset hash table for large offsets).  The next 32 events are already in flight while a chunk is folded. */
#ifndef IPX_K2_BLOCKS
#define IPX_K2_BLOCKS 5 /* resident blocks per SM */
#endif
template <int MODE>
__global__ void __launch_bounds__(256, IPX_K2_BLOCKS) k_fold(Tables t, const Event *ev, const TileInfo *tile_info, uint64_t n_tiles) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    OlogWriter olog{0ull, 0u};
    uint2 next = make_uint2(0u, 0u);
    if (warp0 < n_tiles) next = __ldg(reinterpret_cast<const uint2 *>(tile_info + warp0));
    for (uint64_t tile = warp0; tile < n_tiles; tile += n_warps) {
        const uint32_t cnt = next.x;
        const uint64_t base = tile_event_base(next.y);
        if (tile + n_warps < n_tiles) next = __ldg(reinterpret_cast<const uint2 *>(tile_info + tile + n_warps));
        if (cnt == 0u) continue;
        Event nxt = load_event(ev, base + lane, (uint32_t)lane < cnt, lane);
        for (uint32_t k0 = 0; k0 < cnt; k0 += 32u) {
            const Event e = nxt;
            const bool valid = k0 + lane < cnt;
            if (k0 + 32u < cnt) nxt = load_event(ev, base + k0 + 32u + lane, k0 + 32u + lane < cnt, lane);
            const uint32_t cat = (e.off_cat >> 16) & 3u;
            const int cls = (int)((e.off_cat >> 18) & 3u);
            const unsigned grp = __match_any_sync(0xFFFFFFFFu, e.key);
            const int leader = __ffs(grp) - 1;
            Probe pr;
            pr.slot = t.cap_mask + 1;
            pr.a0 = 0xFFFFFFFFu; pr.a1 = 0u; pr.first = EMPTY64; pr.T = EMPTY64;
            if (lane == leader && valid) pr = probe_junction<MODE != MODE_PASS2>(t, e.key, cls);
            const uint64_t slot = __shfl_sync(0xFFFFFFFFu, pr.slot, leader);
            bool counted = valid && slot <= t.cap_mask;
            JStat *s = &t.stat[counted ? slot : 0];
            if (MODE == MODE_PASS2) {
                const unsigned long long T = __shfl_sync(0xFFFFFFFFu, pr.T, leader);
                counted = counted && T != EMPTY64 && e.sv >= T; /* earlier events are dropped */
            } else {
                /* first-seen sequence value of this lane's strand class */
                const int leader_cls = __shfl_sync(0xFFFFFFFFu, cls, leader);
                const unsigned long long seen = __shfl_sync(0xFFFFFFFFu, pr.first, leader);
                if (counted && (cls != leader_cls || e.sv < seen)) atomicMin(&s->first[cls], e.sv);
            }
            if (MODE != MODE_PASS1) {
                const unsigned c0 = __ballot_sync(0xFFFFFFFFu, counted && cat == 0u) & grp;
                const unsigned c1 = __ballot_sync(0xFFFFFFFFu, counted && cat == 1u) & grp;
                const unsigned c2 = __ballot_sync(0xFFFFFFFFu, counted && cat == 2u) & grp;
                if (lane == leader) {
                    if (c0) atomicAdd(&s->cnt[0], (unsigned long long)__popc(c0));
                    if (c1) atomicAdd(&s->cnt[1], (unsigned long long)__popc(c1));
                    if (c2) atomicAdd(&s->cnt[2], (unsigned long long)__popc(c2));
                }
                const uint32_t seen_a0 = __shfl_sync(0xFFFFFFFFu, pr.a0, leader);
                const uint32_t seen_a1 = __shfl_sync(0xFFFFFFFFu, pr.a1, leader);
                const uint32_t off = e.off_cat & 0xFFFFu;
                const bool in_row = t.dense != nullptr && off < DENSE_W;
                if (counted) {
                    if (e.a0 < seen_a0) atomicMin(&s->a0, e.a0);
                    if (e.a1 > seen_a1) atomicMax(&s->a1, e.a1);
                    /* Junction::count_read's push_back (junctions.hh:163-167), run-length form */
                    if (in_row) atomicAdd(&t.dense[slot * DENSE_W + off], 1u);
                }
                olog_append(t, olog, counted && !in_row, olog_entry(slot, off, 1u));
            }
        }
    }
    if (MODE != MODE_PASS1) olog_finish(t, olog);
}

/* K2, one lane per event, EPL events per lane and iteration.  No grouping: every lane probes for its own keys (lanes of
 * a warp mostly ask for the same few table lines, which the memory system merges) and sends its own fire-and-forget
 * atomics.  The probes of a lane's EPL events are independent, so their latencies overlap. */
template <int MODE, int EPL>
__global__ void __launch_bounds__(256, 4) k_fold_lanes(Tables t, const Event *ev, const TileInfo *tile_info, uint64_t n_tiles,
                                                        int dbg) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    OlogWriter olog{0ull, 0u};
    uint2 next = make_uint2(0u, 0u);
    if (warp0 < n_tiles) next = __ldg(reinterpret_cast<const uint2 *>(tile_info + warp0));
    for (uint64_t tile = warp0; tile < n_tiles; tile += n_warps) {
        const uint32_t cnt = next.x;
        const uint64_t base = tile_event_base(next.y);
        if (tile + n_warps < n_tiles) next = __ldg(reinterpret_cast<const uint2 *>(tile_info + tile + n_warps));
        if (cnt == 0u) continue;
        for (uint32_t k0 = 0; k0 < cnt; k0 += 32u * EPL) {
            Event e[EPL];
            Probe pr[EPL];
            bool valid[EPL], counted[EPL];
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                valid[j] = k0 + j * 32u + lane < cnt;
                e[j] = load_event(ev, base + k0 + j * 32u + lane, valid[j], lane);
            }
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                pr[j].slot = t.cap_mask + 1;
                if (valid[j] && !(dbg & 4)) pr[j] = probe_junction<MODE != MODE_PASS2>(t, e[j].key, (int)((e[j].off_cat >> 18) & 3u));
            }
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                const uint32_t cat = (e[j].off_cat >> 16) & 3u;
                const int cls = (int)((e[j].off_cat >> 18) & 3u);
                counted[j] = valid[j] && pr[j].slot <= t.cap_mask;
                if (MODE == MODE_PASS2) counted[j] = counted[j] && pr[j].T != EMPTY64 && e[j].sv >= pr[j].T; /* earlier: dropped */
                if (counted[j]) {
                    JStat *s = &t.stat[pr[j].slot];
                    if (MODE != MODE_PASS2 && e[j].sv < pr[j].first) atomicMin(&s->first[cls], e[j].sv);
                    if (MODE != MODE_PASS1 && !(dbg & 2)) {
                        atomicAdd(&s->cnt[cat], 1ull);
                        if (e[j].a0 < pr[j].a0) atomicMin(&s->a0, e[j].a0);
                        if (e[j].a1 > pr[j].a1) atomicMax(&s->a1, e[j].a1);
                    }
                }
            }
            if (MODE != MODE_PASS1) {
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    const uint32_t off = e[j].off_cat & 0xFFFFu;
                    const bool in_row = t.dense != nullptr && off < DENSE_W;
                    /* Junction::count_read's push_back (junctions.hh:163-167), run-length form */
                    if (counted[j] && in_row && !(dbg & 1)) atomicAdd(&t.dense[pr[j].slot * DENSE_W + off], 1u);
                    __syncwarp();
                    olog_append(t, olog, counted[j] && !in_row, olog_entry(pr[j].slot, off, 1u));
                }
            }
        }
    }
    if (MODE != MODE_PASS1) olog_finish(t, olog);
}

void launch_fold(int mode, const Tables &t, const Event *ev, const TileInfo *tile_info, uint64_t n_tiles, int sm_count,
                 cudaStream_t st) {
    if (n_tiles == 0) return;
    const uint64_t want = (n_tiles + 7) / 8; /* 8 warps per block, one tile per warp at a time */
    /* measured on B200 (100 M short reads): grouped 0.62 ms, one lane per event 0.64 - 0.72 ms */
    static const bool grouped = getenv("IPX_K2_LANES") == nullptr;
    if (!grouped) {
        static const int dbg = getenv("IPX_K2_DBG") ? atoi(getenv("IPX_K2_DBG")) : 0;
        const uint64_t max_lane_blocks = (uint64_t)sm_count * 4;
        static const int epl = getenv("IPX_K2_EPL") ? atoi(getenv("IPX_K2_EPL")) : 2;
        const int lb = (int)(want < max_lane_blocks ? want : max_lane_blocks);
        switch (mode) {
            case MODE_SINGLE: if (epl == 1) k_fold_lanes<MODE_SINGLE, 1><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); else if (epl == 4) k_fold_lanes<MODE_SINGLE, 4><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); else k_fold_lanes<MODE_SINGLE, 2><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); break;
            case MODE_PASS1: if (epl == 1) k_fold_lanes<MODE_PASS1, 1><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); else if (epl == 4) k_fold_lanes<MODE_PASS1, 4><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); else k_fold_lanes<MODE_PASS1, 2><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); break;
            default: if (epl == 1) k_fold_lanes<MODE_PASS2, 1><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); else if (epl == 4) k_fold_lanes<MODE_PASS2, 4><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); else k_fold_lanes<MODE_PASS2, 2><<<lb, 256, 0, st>>>(t, ev, tile_info, n_tiles, dbg); break;
        }
        return;
    }
    const uint64_t max_blocks = (uint64_t)sm_count * IPX_K2_BLOCKS; /* what is resident at once */
    const int blocks = (int)(want < max_blocks ? want : max_blocks);
    switch (mode) {
        case MODE_SINGLE: k_fold<MODE_SINGLE><<<blocks, 256, 0, st>>>(t, ev, tile_info, n_tiles); break;
        case MODE_PASS1: k_fold<MODE_PASS1><<<blocks, 256, 0, st>>>(t, ev, tile_info, n_tiles); break;
        default: k_fold<MODE_PASS2><<<blocks, 256, 0, st>>>(t, ev, tile_info, n_tiles); break;
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
    const unsigned long long nk = *t.n_keys;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nk; j += stride) {
        const uint64_t i = t.slot_list[j];
        const unsigned long long key = t.keys[i];
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
    k_resolve_T<<<148 * 4, 256, 0, st>>>(t, g);
}


/* ------------------------------------------------------------------ finalize */
/* Everything here walks slot_list (the slots in use), never the table's capacity, inside buffers of fixed bounds
 * (see FinalizeWork).  Counts stay on the device: d_count[0] junctions (n_j <= keys in use: keys whose every event was
 * rejected by the canon filter are not junctions), d_count[1] histogram runs, d_count[2] distinct log runs,
 * d_count[3] flags for bounds that were too small. */
enum : unsigned long long { FLAG_B = 1ull, FLAG_L = 2ull, FLAG_R = 4ull };

__global__ void k_finalize_begin(Tables t, unsigned long long *d_count, uint64_t B, uint64_t L) {
    const unsigned long long nk = *t.n_keys, nl = *t.n_olog;
    d_count[0] = d_count[1] = d_count[2] = 0ull;
    d_count[3] = (nk > B ? FLAG_B : 0ull) | (nl > L ? FLAG_L : 0ull);
    d_count[4] = nk;
    d_count[5] = nl;
}

/* sort keys of the slots in use: chromosome-name rank, intron_start, intron_end; rejected keys and padding sort last */
__global__ void k_list_keys(Tables t, const uint32_t *name_rank, uint64_t B, int key_bits, unsigned long long *d_count,
                            unsigned long long *sort_key, uint32_t *slot_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nk = *t.n_keys < B ? *t.n_keys : B;
    bool take = false;
    if (i < B) {
        unsigned long long sk = 1ull << key_bits;
        uint32_t slot = 0u;
        if (i < nk) {
            slot = t.slot_list[i];
            const unsigned long long key = t.keys[slot];
            take = t.stat[slot].T != EMPTY64;
            const uint64_t tid = (uint64_t)key_tid(t.kl, key);
            const uint64_t low = key & ((1ull << t.kl.tid_shift) - 1ull);
            if (take) sk = ((unsigned long long)name_rank[tid] << t.kl.tid_shift) | low;
        }
        sort_key[i] = sk;
        slot_out[i] = slot;
    }
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, take);
    if (mask && (threadIdx.x & 31) == 0) atomicAdd(&d_count[0], (unsigned long long)__popc(mask));
}

/* sjN chain, keyed by position in slot_list (j): sort key = the creating event's position in the stream (sv without
 * its strand byte); keys that are not junctions and padding sort last */
__global__ void k_list_T(Tables t, uint64_t B, int t_bits, unsigned long long *T_out, uint32_t *idx_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const unsigned long long nk = *t.n_keys < B ? *t.n_keys : B;
    unsigned long long T = EMPTY64;
    if (i < nk) T = t.stat[t.slot_list[i]].T;
    T_out[i] = T == EMPTY64 ? (1ull << t_bits) : (T >> 8);
    idx_out[i] = (uint32_t)i;
}

/* t_rank[j] = rank of junction j's creating event = its sjN */
__global__ void k_scatter_trank(const unsigned long long *sorted_T, const uint32_t *j_by_T, uint64_t B, int t_bits, uint32_t *t_rank) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B && !(sorted_T[i] >> t_bits)) t_rank[j_by_T[i]] = (uint32_t)i;
}

/* one output row per junction, in key order: attributes of the creating event after motif correction
 * (get_splice_sites + correct_strand, junctions_extractor.cc:419-444) */
__global__ void k_gather_rows(Tables t, GenomeView g, const uint32_t *sorted_slot, const unsigned long long *d_count,
                              const uint32_t *t_rank, DeviceTable out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d_count[0]) return;
    const uint32_t slot = sorted_slot[i];
    out.ijunc[i] = t_rank[t.slot_id[slot]];
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

/* every offset-log entry -> (junction j << 16 | offset, count), j = the slot's position in slot_list; padding, invalid
 * entries and entries of slots that are not junctions (none are expected) get a key above every real one */
__global__ void k_olog_pairs(Tables t, uint64_t B, uint64_t L, int h_bits, unsigned long long *hkey_out, uint32_t *hcnt_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const unsigned long long nl = *t.n_olog < L ? *t.n_olog : L;
    const unsigned long long e = i < nl ? t.olog[i] : OLOG_INVALID;
    bool ok = e != OLOG_INVALID;
    uint32_t j = 0u;
    if (ok) {
        const uint32_t slot = (uint32_t)(e >> 32);
        j = t.slot_id[slot];
        /* a junction beyond the bound this finalize was sized for has no room in the per-junction arrays: the
         * bound is reported as too small anyway (k_finalize_begin) and the finalize runs again */
        ok = t.stat[slot].T != EMPTY64 && j < B;
    }
    hkey_out[i] = ok ? (((unsigned long long)j << 16) | ((e >> 16) & 0xFFFFull)) : (1ull << h_bits);
    hcnt_out[i] = (uint32_t)(e & 0xFFFFull);
}

/* per junction j: first index and number of its runs in the sorted, run-length reduced log */
__global__ void k_hash_runs(const unsigned long long *sorted_hkey, const uint32_t *n_runs, int h_bits, unsigned long long *hash_lo,
                            unsigned long long *total) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_runs) return;
    if (sorted_hkey[i] >> h_bits) return;
    const uint64_t r = sorted_hkey[i] >> 16;
    if (i == 0 || (sorted_hkey[i - 1] >> 16) != r) hash_lo[r] = i;
    atomicAdd(&total[r], 1ull);
}

/* warp per junction: number of non-zero counters of its dense row, added to total[r] */
__global__ void k_dense_count(Tables t, const uint32_t *sorted_slot, const unsigned long long *d_count,
                              const unsigned long long *total_j, unsigned long long *total, uint32_t *dense_nz) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= d_count[0]) return;
    uint32_t nz = 0;
    if (t.dense) {
        const uint4 *row = reinterpret_cast<const uint4 *>(t.dense + (uint64_t)sorted_slot[r] * DENSE_W);
        const uint4 a = row[lane * 2], b = row[lane * 2 + 1];
        nz = (a.x != 0) + (a.y != 0) + (a.z != 0) + (a.w != 0) + (b.x != 0) + (b.y != 0) + (b.z != 0) + (b.w != 0);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) nz += __shfl_xor_sync(0xFFFFFFFFu, nz, k);
    if (lane == 0) {
        dense_nz[r] = nz;
        total[r] = nz + total_j[t.slot_id[sorted_slot[r]]]; /* + its runs in the log */
    }
}

/* warp per junction: write its runs in ascending offset order -- dense row first (offsets < DENSE_W), then its
 * slice of the sorted log runs (offsets >= DENSE_W, or every offset when the table has no dense rows) */
__global__ void k_write_runs(Tables t, const uint32_t *sorted_slot, unsigned long long *d_count, uint64_t B, uint64_t R,
                             const unsigned long long *hist_off, const uint32_t *dense_nz, const unsigned long long *hash_lo,
                             const uint32_t *sorted_hcnt, uint32_t *hist_cnt) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r == 0 && lane == 0) {
        const unsigned long long n_runs = hist_off[B]; /* total[] is zero beyond the junctions */
        d_count[1] = n_runs;
        if (n_runs > R) atomicOr(&d_count[3], FLAG_R);
    }
    if (r >= d_count[0]) return;
    const unsigned long long o0 = hist_off[r], o1 = hist_off[r + 1];
    if (o1 > R) return; /* does not fit: the host grows R and replays */
    const uint32_t nz = dense_nz[r];
    if (t.dense && nz) {
        const uint4 *row = reinterpret_cast<const uint4 *>(t.dense + (uint64_t)sorted_slot[r] * DENSE_W);
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
            if (v[k] != 0) hist_cnt[o++] = v[k];
        }
    }
    const unsigned long long n_hash = (o1 - o0) - nz;
    if (n_hash) {
        const unsigned long long lo = hash_lo[t.slot_id[sorted_slot[r]]];
        for (unsigned long long j = lane; j < n_hash; j += 32) hist_cnt[o0 + nz + j] = sorted_hcnt[lo + j];
    }
}

#define IPX_CU(x)                          \
    do {                                   \
        cudaError_t e__ = (x);             \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

static inline unsigned blocks_for(uint64_t n, unsigned threads) {
    uint64_t b = (n + threads - 1) / threads;
    return (unsigned)(b ? b : 1);
}

static inline int bits_for64(uint64_t v) {
    int b = 1;
    while (b < 64 && (v >> b) != 0) b++;
    return b;
}

/* a bound for v with some headroom: v plus an eighth, rounded up to a multiple of 4096, at least floor_ (padding costs
 * sort work, so the bounds stay close to what was seen) */
static inline uint64_t pow2_ge(uint64_t v, uint64_t floor_) {
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
    cudaFree(w.sort_key); cudaFree(w.sort_key2); cudaFree(w.hash_lo); cudaFree(w.total); cudaFree(w.hk_a); cudaFree(w.hk_b);
    cudaFree(w.slot_a); cudaFree(w.slot_b); cudaFree(w.idx_b); cudaFree(w.dense_nz); cudaFree(w.hc_a); cudaFree(w.hc_b);
    cudaFree(w.tkey); cudaFree(w.tkey2); cudaFree(w.total_j); cudaFree(w.tidx_a); cudaFree(w.tidx_b); cudaFree(w.t_rank);
    cudaFree(w.cub_tmp); cudaFree(w.cub_tmp2); cudaFree(w.cub_tmp3); cudaFree(w.d_count);
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

/* (re)allocate everything whose size follows the bounds */
static cudaError_t finalize_allocate(FinalizeWork &w, uint64_t cap, uint64_t B, uint64_t L, uint64_t R) {
    if (!w.d_count) IPX_CU(cudaMalloc(&w.d_count, 8 * sizeof(unsigned long long)));
    if (!w.h_count) IPX_CU(cudaMallocHost(&w.h_count, 8 * sizeof(unsigned long long)));
    (void)cap;
    if (!w.s2) {
        IPX_CU(cudaStreamCreateWithFlags(&w.s2, cudaStreamNonBlocking));
        IPX_CU(cudaStreamCreateWithFlags(&w.s3, cudaStreamNonBlocking));
        IPX_CU(cudaEventCreateWithFlags(&w.ev_fork, cudaEventDisableTiming));
        IPX_CU(cudaEventCreateWithFlags(&w.ev_join2, cudaEventDisableTiming));
        IPX_CU(cudaEventCreateWithFlags(&w.ev_join3, cudaEventDisableTiming));
    }
    if (B != w.B) {
        IPX_CU(regrow(&w.tkey, B)); IPX_CU(regrow(&w.tkey2, B)); IPX_CU(regrow(&w.total_j, B)); IPX_CU(regrow(&w.tidx_a, B));
        IPX_CU(regrow(&w.tidx_b, B)); IPX_CU(regrow(&w.t_rank, B));
        IPX_CU(regrow(&w.sort_key, B)); IPX_CU(regrow(&w.sort_key2, B)); IPX_CU(regrow(&w.slot_a, B)); IPX_CU(regrow(&w.slot_b, B));
        IPX_CU(regrow(&w.idx_b, B)); IPX_CU(regrow(&w.hash_lo, B)); IPX_CU(regrow(&w.total, B + 1)); IPX_CU(regrow(&w.dense_nz, B));
        DeviceTable &o = w.out;
        IPX_CU(regrow(&o.chrom, B)); IPX_CU(regrow(&o.i0, B)); IPX_CU(regrow(&o.i1, B)); IPX_CU(regrow(&o.a0, B));
        IPX_CU(regrow(&o.a1, B)); IPX_CU(regrow(&o.ijunc, B)); IPX_CU(regrow(&o.strand, B)); IPX_CU(regrow(&o.sites, B * 8));
        IPX_CU(regrow(&o.counts, B * 3)); IPX_CU(regrow(&o.T, B)); IPX_CU(regrow(&o.hist_off, B + 2));
        w.B = B;
    }
    if (L != w.L) {
        IPX_CU(regrow(&w.hk_a, L)); IPX_CU(regrow(&w.hk_b, L)); IPX_CU(regrow(&w.hc_a, L)); IPX_CU(regrow(&w.hc_b, L));
        w.L = L;
    }
    if (R != w.R) {
        IPX_CU(regrow(&w.out.hist_cnt, R));
        w.R = R;
    }
    return cudaSuccess;
}

/* the sequence itself, issued into `st` (under stream capture, or directly when capture is not possible) */
static cudaError_t finalize_issue(const Tables &t, const GenomeView &g, const uint32_t *d_name_rank, FinalizeWork &w, cudaStream_t st) {
    const uint64_t B = w.B, L = w.L, R = w.R;
    const int key_bits = w.key.key_bits, t_bits = w.key.t_bits, h_bits = 16 + bits_for64(B);
    DeviceTable &out = w.out;
    size_t tb = 0;
    k_resolve_T<<<148 * 4, 256, 0, st>>>(t, g); /* creation thresholds for anything still unresolved */
    k_finalize_begin<<<1, 1, 0, st>>>(t, w.d_count, B, L);
    /* three independent chains from here: key order (st), sjN ranking (s2), read-offset log (s3) */
    IPX_CU(cudaEventRecord(w.ev_fork, st));
    IPX_CU(cudaStreamWaitEvent(w.s2, w.ev_fork, 0));
    IPX_CU(cudaStreamWaitEvent(w.s3, w.ev_fork, 0));

    /* st: key order -- chromosome-name rank, intron_start, intron_end */
    k_list_keys<<<blocks_for(B, 256), 256, 0, st>>>(t, d_name_rank, B, key_bits, w.d_count, w.sort_key, w.slot_a);
    tb = w.cub_tmp_bytes;
    IPX_CU(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.sort_key, w.sort_key2, w.slot_a, w.slot_b, (int)B, 0,
                                           key_bits + 1 > 64 ? 64 : key_bits + 1, st));
    const uint32_t *sorted_slot = w.slot_b;

    /* s2: sjN = rank of the creating event in stream order */
    k_list_T<<<blocks_for(B, 256), 256, 0, w.s2>>>(t, B, t_bits, w.tkey, w.tidx_a);
    tb = w.cub_tmp_bytes;
    IPX_CU(cub::DeviceRadixSort::SortPairs(w.cub_tmp2, tb, w.tkey, w.tkey2, w.tidx_a, w.tidx_b, (int)B, 0, t_bits + 1, w.s2));
    k_scatter_trank<<<blocks_for(B, 256), 256, 0, w.s2>>>(w.tkey2, w.tidx_b, B, t_bits, w.t_rank);
    IPX_CU(cudaEventRecord(w.ev_join2, w.s2));

    /* s3: read-offset log ordered by (junction, offset) with a sort, equal pairs summed */
    IPX_CU(cudaMemsetAsync(w.total_j, 0, B * 8, w.s3));
    k_olog_pairs<<<blocks_for(L, 256), 256, 0, w.s3>>>(t, B, L, h_bits, w.hk_a, w.hc_a);
    tb = w.cub_tmp_bytes;
    IPX_CU(cub::DeviceRadixSort::SortPairs(w.cub_tmp3, tb, w.hk_a, w.hk_b, w.hc_a, w.hc_b, (int)L, 0, h_bits + 1, w.s3));
    uint32_t *n_unique = reinterpret_cast<uint32_t *>(w.d_count + 2);
    tb = w.cub_tmp_bytes;
    IPX_CU(cub::DeviceReduce::ReduceByKey(w.cub_tmp3, tb, w.hk_b, w.hk_a, w.hc_b, w.hc_a, n_unique, cub::Sum(), (int)L, w.s3));
    k_hash_runs<<<blocks_for(L, 256), 256, 0, w.s3>>>(w.hk_a, n_unique, h_bits, w.hash_lo, w.total_j);
    IPX_CU(cudaEventRecord(w.ev_join3, w.s3));

    /* join: rows in key order (motif, strand correction, sjN), then the histograms -- dense rows are already in
     * offset order, the log runs follow them */
    IPX_CU(cudaStreamWaitEvent(st, w.ev_join2, 0));
    k_gather_rows<<<blocks_for(B, 256), 256, 0, st>>>(t, g, sorted_slot, w.d_count, w.t_rank, out);
    IPX_CU(cudaStreamWaitEvent(st, w.ev_join3, 0));
    IPX_CU(cudaMemsetAsync(w.total, 0, (B + 1) * 8, st));
    k_dense_count<<<blocks_for(B * 32, 256), 256, 0, st>>>(t, sorted_slot, w.d_count, w.total_j, w.total, w.dense_nz);
    tb = w.cub_tmp_bytes;
    IPX_CU(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.total, out.hist_off, (int)(B + 1), st));
    k_write_runs<<<blocks_for(B * 32, 256), 256, 0, st>>>(t, sorted_slot, w.d_count, B, R, out.hist_off, w.dense_nz, w.hash_lo, w.hc_a,
                                                        out.hist_cnt);
    IPX_CU(cudaMemcpyAsync(w.h_count, w.d_count, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    w.launches_per_replay = 12;
    return cudaGetLastError();
}

cudaError_t finalize_enqueue(const Tables &t, const GenomeView &g, const uint32_t *d_name_rank, int32_t n_targets,
                             uint64_t seq_hi, FinalizeWork &w, cudaStream_t st) {
    const uint64_t cap = t.cap_mask + 1;
    uint64_t B = w.B ? w.B : w.first_B, L = w.L ? w.L : w.first_L, R = w.R ? w.R : w.first_R;
    if (w.wish_B > B) B = w.wish_B;
    if (w.wish_L > L) L = w.wish_L;
    if (w.wish_R > R) R = w.wish_R;
    FinalizeWork::Key key{};
    key.B = B; key.L = L; key.R = R;
    key.key_bits = t.kl.tid_shift + bits_for64((uint64_t)(n_targets > 0 ? n_targets : 1));
    key.t_bits = (SV_SEQ_SHIFT - 8) + bits_for64(seq_hi ? seq_hi : 1);
    key.tab_keys = t.keys; key.name_rank = d_name_rank; key.gseq = g.seq; key.target_seq = g.target_seq; key.dense = t.dense;
    key.n_targets = g.n_targets; key.have_genome = g.have_genome; key.skip_missing = g.skip_missing; key.filter_canon = g.filter_canon;
    if (!w.exec || !(key == w.key)) {
        if (w.exec) {
            IPX_CU(cudaStreamSynchronize(st)); /* a replay may still be running */
            cudaGraphExecDestroy(w.exec);
            w.exec = nullptr;
        }
        IPX_CU(cudaStreamSynchronize(st));
        IPX_CU(finalize_allocate(w, cap, B, L, R));
        w.key = key;
        /* temporary storage of the library calls, sized for the largest of them */
        size_t need = 0, tb = 0;
        IPX_CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, w.sort_key, w.sort_key2, w.slot_a, w.slot_b, (int)B, 0, 64, st));
        need = tb > need ? tb : need;
        IPX_CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, w.hk_a, w.hk_b, w.hc_a, w.hc_b, (int)L, 0, 64, st));
        need = tb > need ? tb : need;
        IPX_CU(cub::DeviceReduce::ReduceByKey(nullptr, tb, w.hk_b, w.hk_a, w.hc_b, w.hc_a, (uint32_t *)nullptr, cub::Sum(), (int)L, st));
        need = tb > need ? tb : need;
        IPX_CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, w.total, w.out.hist_off, (int)(B + 1), st));
        need = tb > need ? tb : need;
        if (need > w.cub_tmp_bytes) {
            cudaFree(w.cub_tmp); cudaFree(w.cub_tmp2); cudaFree(w.cub_tmp3);
            w.cub_tmp = w.cub_tmp2 = w.cub_tmp3 = nullptr;
            IPX_CU(cudaMalloc(&w.cub_tmp, need));
            IPX_CU(cudaMalloc(&w.cub_tmp2, need));
            IPX_CU(cudaMalloc(&w.cub_tmp3, need));
            w.cub_tmp_bytes = need;
        }
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
    return finalize_issue(t, g, d_name_rank, w, st);
}

bool finalize_check(FinalizeWork &w) {
    const unsigned long long flags = w.h_count[3];
    if (flags == 0ull) {
        w.out.n = w.h_count[0];
        w.out.n_hist = w.h_count[1];
        return true;
    }
    /* grow what was too small; the next finalize_enqueue reallocates and re-captures */
    if (flags & FLAG_B) w.wish_B = pow2_ge(w.h_count[4], w.B + 4096);
    if (flags & FLAG_L) w.wish_L = pow2_ge(w.h_count[5], w.L + 4096);
    if (flags & FLAG_R) w.wish_R = pow2_ge(w.h_count[1], w.R + 4096);
    w.out.n = 0;
    w.out.n_hist = 0;
    return false;
}

void free_device_table(DeviceTable &, cudaStream_t) {}

/* ------------------------------------------------------------------ multi-GPU exchange */
__device__ __forceinline__ int part_of(const Splitters &sp, uint32_t name_rank, uint32_t i0) {
    int p = 0;
    while (p < sp.n_parts - 1 && (name_rank > sp.rank[p] || (name_rank == sp.rank[p] && i0 >= sp.start[p]))) p++;
    return p;
}

/* warp per slot in use: its destination and the number of non-zero counters of its dense offset row.
 * SCATTER = false: count rows per destination; true: write them behind the destination's cursor. */
__device__ __forceinline__ uint32_t hist_rows_of(uint32_t v) { return v ? (v - 1u) / 0xFFFFu + 1u : 0u; }

template <bool SCATTER>
__global__ void __launch_bounds__(256) k_export_slots(Tables t, const uint32_t *name_rank, Splitters sp, uint64_t n_keys,
                                                      unsigned long long *counts, uint32_t *slot_jrow, JunctionRow *jr, HistRow *hr) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_keys || i >= *t.n_keys) return; /* n_keys is a launch bound; the table knows how many there are */
    const uint32_t slot = t.slot_list[i];
    const JStat s = t.stat[slot];
    if (s.T == EMPTY64) return; /* never created: nothing was counted into it */
    const unsigned long long key = t.keys[slot];
    const int32_t tid = key_tid(t.kl, key);
    const uint32_t i0 = key_start(t.kl, key), i1 = i0 + key_len(t.kl, key);
    const int dest = part_of(sp, name_rank[tid], i0);
    uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    if (t.dense) {
        const uint4 *row = reinterpret_cast<const uint4 *>(t.dense + (uint64_t)slot * DENSE_W);
        const uint4 a = row[lane * 2], b = row[lane * 2 + 1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
        for (int k = 0; k < 8; ++k) mine += hist_rows_of(v[k]);
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
    }
    const uint32_t nz = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long jpos = 0ull, hpos = 0ull;
    if (lane == 0) {
        jpos = atomicAdd(&counts[dest], 1ull);
        if (nz) hpos = atomicAdd(&counts[sp.n_parts + dest], (unsigned long long)nz);
    }
    if (!SCATTER) return;
    jpos = __shfl_sync(0xFFFFFFFFu, jpos, 0);
    hpos = __shfl_sync(0xFFFFFFFFu, hpos, 0);
    const uint32_t jrow = (uint32_t)jpos - sp.jbase[dest];
    if (lane == 0) {
        JunctionRow r;
        r.first[0] = s.first[0]; r.first[1] = s.first[1]; r.first[2] = s.first[2];
        r.cnt[0] = s.cnt[0]; r.cnt[1] = s.cnt[1]; r.cnt[2] = s.cnt[2];
        r.tid = tid; r.i0 = i0; r.i1 = i1; r.a0 = s.a0; r.a1 = s.a1; r.pad = 0;
        jr[jpos] = r;
        slot_jrow[slot] = jrow; /* for the log entries of this junction */
    }
    uint64_t o = hpos + (incl - mine);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t left = v[k];
        while (left) {
            const uint32_t piece = left > 0xFFFFu ? 0xFFFFu : left;
            HistRow h;
            h.jrow = jrow; h.offset = (uint16_t)(lane * 8 + k); h.cnt = (uint16_t)piece;
            hr[o++] = h;
            left -= piece;
        }
    }
}

/* thread per entry of the read-offset log: one offset row each (equal pairs are summed by the importer) */
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_export_olog(Tables t, const uint32_t *name_rank, Splitters sp, uint64_t n_olog,
                                                     unsigned long long *counts, const uint32_t *slot_jrow, HistRow *hr) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int dest = -1 - lane; /* idle lanes group with nobody */
    unsigned long long e = OLOG_INVALID;
    uint32_t slot = 0u;
    if (i < n_olog && i < *t.n_olog) {
        e = t.olog[i];
        if (e != OLOG_INVALID) {
            slot = (uint32_t)(e >> 32);
            if (t.stat[slot].T != EMPTY64) {
                const unsigned long long key = t.keys[slot];
                dest = part_of(sp, name_rank[key_tid(t.kl, key)], key_start(t.kl, key));
            }
        }
    }
    /* one atomic per (warp, destination) */
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, dest);
    const int leader = __ffs(grp) - 1;
    unsigned long long base = 0ull;
    if (dest >= 0 && lane == leader) base = atomicAdd(&counts[sp.n_parts + dest], (unsigned long long)__popc(grp));
    if (!SCATTER) return;
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (dest >= 0) {
        HistRow h;
        h.jrow = slot_jrow[slot]; h.offset = (uint16_t)((e >> 16) & 0xFFFFull); h.cnt = (uint16_t)(e & 0xFFFFull);
        hr[base + __popc(grp & ((1u << lane) - 1u))] = h;
    }
}

void launch_export_count(const Tables &t, const uint32_t *name_rank, const Splitters &sp, uint64_t n_keys, uint64_t n_olog,
                         unsigned long long *counts, cudaStream_t st) {
    if (n_keys) k_export_slots<false><<<blocks_for(n_keys * 32, 256), 256, 0, st>>>(t, name_rank, sp, n_keys, counts, nullptr, nullptr, nullptr);
    if (n_olog) k_export_olog<false><<<blocks_for(n_olog, 256), 256, 0, st>>>(t, name_rank, sp, n_olog, counts, nullptr, nullptr);
}

void launch_export_scatter(const Tables &t, const uint32_t *name_rank, const Splitters &sp, uint64_t n_keys, uint64_t n_olog,
                           unsigned long long *cursors, uint32_t *slot_jrow, JunctionRow *jr, HistRow *hr, cudaStream_t st) {
    if (n_keys) k_export_slots<true><<<blocks_for(n_keys * 32, 256), 256, 0, st>>>(t, name_rank, sp, n_keys, cursors, slot_jrow, jr, hr);
    if (n_olog) k_export_olog<true><<<blocks_for(n_olog, 256), 256, 0, st>>>(t, name_rank, sp, n_olog, cursors, slot_jrow, hr);
}

__global__ void k_import_junction_rows(Tables t, const JunctionRow *jr, uint64_t n, uint32_t *slot_of) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const JunctionRow r = jr[i];
    const uint64_t key = make_key(t.kl, r.tid, r.i0, r.i1);
    const uint64_t slot = junction_find_or_insert(t, key);
    slot_of[i] = slot <= t.cap_mask ? (uint32_t)slot : 0xFFFFFFFFu;
    if (slot > t.cap_mask) return;
    JStat *s = &t.stat[slot];
    for (int c = 0; c < 3; ++c) {
        if (r.first[c] != EMPTY64) atomicMin(&s->first[c], r.first[c]);
        if (r.cnt[c]) atomicAdd(&s->cnt[c], r.cnt[c]);
    }
    atomicMin(&s->a0, r.a0);
    atomicMax(&s->a1, r.a1);
}

__global__ void k_import_hist_rows(Tables t, const HistRow *hr, uint64_t n, const uint32_t *slot_of, uint64_t nj) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    OlogWriter olog{0ull, 0u};
    uint32_t slot = 0xFFFFFFFFu, off = 0u, cnt = 0u;
    if (i < n) {
        const HistRow r = hr[i];
        if (r.jrow < nj) slot = slot_of[r.jrow];
        off = r.offset;
        cnt = r.cnt;
    }
    const bool ok = slot != 0xFFFFFFFFu && cnt != 0u;
    const bool in_row = t.dense != nullptr && off < DENSE_W;
    if (ok && in_row) atomicAdd(&t.dense[(uint64_t)slot * DENSE_W + off], cnt);
    olog_append(t, olog, ok && !in_row, olog_entry(slot, off, cnt));
    olog_finish(t, olog);
}

/* ---- the exchange over peer memory: export kernels that store straight into the owners' mailboxes ----
 * Every rank owns a mailbox (ipx_engine.h: Mailboxes) with one slot per source rank; the exporter's warps decide the
 * owner of a junction and write its rows behind a per-owner cursor directly into that owner's slot for this rank --
 * NVLink stores for the peers, plain stores for its own share.  One pass: a slot has room for the worst case, so
 * there is no counting pass, no size exchange and no send buffer; the row counts are published into the slot's header
 * by a last small kernel.  The owners fold their slots after a barrier the caller provides (any collective on the
 * streams involved: kernel completion makes the stores visible system-wide). */
__device__ __forceinline__ char *mail_slot(const Mailboxes &m, int owner, int source) {
    return m.base[owner] + (uint64_t)(m.parity * m.world + source) * m.slot_bytes;
}
__device__ __forceinline__ JunctionRow *mail_junction_rows(char *slot) { return reinterpret_cast<JunctionRow *>(slot + MAIL_HEADER_BYTES); }
__device__ __forceinline__ HistRow *mail_hist_rows(const Mailboxes &m, char *slot) {
    return reinterpret_cast<HistRow *>(slot + MAIL_HEADER_BYTES + m.hist_offset);
}

/* what a warp needs of one slot in use: owner, key fields, its dense offset row (8 counters per lane) */
struct MailSlot {
    int dest; /* -1: nothing to export */
    uint32_t slot, i0, i1, mine, incl, nz;
    int32_t tid;
    uint32_t v[8];
};
__device__ __forceinline__ MailSlot mail_read_slot(const Tables &t, const uint32_t *name_rank, const Splitters &sp, uint64_t i, int lane) {
    MailSlot r;
    r.dest = -1;
    r.slot = t.slot_list[i];
    r.mine = 0u;
    /* everything that hangs on the slot index is requested at once: the warp's time is round trips, not bytes */
    const unsigned long long T = t.stat[r.slot].T;
    const unsigned long long key = t.keys[r.slot];
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
    if (t.dense) {
        const uint4 *row = reinterpret_cast<const uint4 *>(t.dense + (uint64_t)r.slot * DENSE_W);
        a = row[lane * 2];
        b = row[lane * 2 + 1];
    }
    r.tid = key_tid(t.kl, key);
    r.i0 = key_start(t.kl, key);
    r.i1 = r.i0 + key_len(t.kl, key);
    const bool live = T != EMPTY64; /* else never created: nothing was counted into it */
    if (live) r.dest = part_of(sp, name_rank[r.tid], r.i0);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        r.v[k] = live ? r.v[k] : 0u;
        r.mine += hist_rows_of(r.v[k]);
    }
    r.incl = r.mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, r.incl, d);
        if (lane >= d) r.incl += y;
    }
    r.nz = __shfl_sync(0xFFFFFFFFu, r.incl, 31);
    return r;
}

/* A warp per slot in use.  The per-owner cursors are single addresses every block wants: a block first adds up what
 * its slots need per owner in shared memory, reserves it with one atomic per owner, and then hands the positions out
 * from shared memory while it writes the rows (the slots are read twice; the second time they come from L2). */
__global__ void __launch_bounds__(256) k_mail_slots(Tables t, const uint32_t *name_rank, Splitters sp, Mailboxes m,
                                                    unsigned long long *counts, uint32_t *slot_jrow) {
    __shared__ unsigned long long s_j[MAX_PARTS], s_h[MAX_PARTS]; /* needed, then next free position */
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_keys = *t.n_keys;
    if (threadIdx.x < MAX_PARTS) s_j[threadIdx.x] = s_h[threadIdx.x] = 0ull;
    __syncthreads();
    for (uint64_t i = warp0; i < n_keys; i += n_warps) {
        const MailSlot r = mail_read_slot(t, name_rank, sp, i, lane);
        if (lane == 0 && r.dest >= 0) {
            atomicAdd(&s_j[r.dest], 1ull);
            if (r.nz) atomicAdd(&s_h[r.dest], (unsigned long long)r.nz);
        }
    }
    __syncthreads();
    if (threadIdx.x < sp.n_parts) {
        const int d = threadIdx.x;
        const unsigned long long nj = s_j[d], nh = s_h[d];
        s_j[d] = nj ? atomicAdd(&counts[d], nj) : 0ull;
        s_h[d] = nh ? atomicAdd(&counts[sp.n_parts + d], nh) : 0ull;
    }
    __syncthreads();
    for (uint64_t i = warp0; i < n_keys; i += n_warps) {
        const MailSlot r = mail_read_slot(t, name_rank, sp, i, lane);
        if (r.dest < 0) continue;
        unsigned long long jpos = 0ull, hpos = 0ull;
        if (lane == 0) {
            jpos = atomicAdd(&s_j[r.dest], 1ull);
            if (r.nz) hpos = atomicAdd(&s_h[r.dest], (unsigned long long)r.nz);
        }
        jpos = __shfl_sync(0xFFFFFFFFu, jpos, 0);
        hpos = __shfl_sync(0xFFFFFFFFu, hpos, 0);
        /* a slot that is too small keeps counting: the owner sees counts beyond the capacity and reports it */
        if (jpos >= m.cap_j || hpos + r.nz > m.cap_h) continue;
        char *box = mail_slot(m, r.dest, m.me);
        if (lane == 0) {
            const JStat s = t.stat[r.slot];
            JunctionRow jr;
            jr.first[0] = s.first[0]; jr.first[1] = s.first[1]; jr.first[2] = s.first[2];
            jr.cnt[0] = s.cnt[0]; jr.cnt[1] = s.cnt[1]; jr.cnt[2] = s.cnt[2];
            jr.tid = r.tid; jr.i0 = r.i0; jr.i1 = r.i1; jr.a0 = s.a0; jr.a1 = s.a1; jr.pad = 0;
            mail_junction_rows(box)[jpos] = jr;
            slot_jrow[r.slot] = (uint32_t)jpos; /* for the log entries of this junction */
        }
        HistRow *hr = mail_hist_rows(m, box);
        uint64_t o = hpos + (r.incl - r.mine);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t left = r.v[k];
            while (left) {
                const uint32_t piece = left > 0xFFFFu ? 0xFFFFu : left;
                HistRow h;
                h.jrow = (uint32_t)jpos; h.offset = (uint16_t)(lane * 8 + k); h.cnt = (uint16_t)piece;
                hr[o++] = h;
                left -= piece;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_mail_olog(Tables t, const uint32_t *name_rank, Splitters sp, Mailboxes m,
                                                   unsigned long long *counts, const uint32_t *slot_jrow) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_olog = *t.n_olog, n_round = (n_olog + 31ull) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += n_threads) {
        int dest = -1 - lane; /* idle lanes group with nobody */
        unsigned long long e = OLOG_INVALID;
        uint32_t slot = 0u;
        if (i < n_olog) {
            e = t.olog[i];
            if (e != OLOG_INVALID) {
                slot = (uint32_t)(e >> 32);
                if (t.stat[slot].T != EMPTY64) {
                    const unsigned long long key = t.keys[slot];
                    dest = part_of(sp, name_rank[key_tid(t.kl, key)], key_start(t.kl, key));
                }
            }
        }
        /* one atomic per (warp, destination) */
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, dest);
        const int leader = __ffs(grp) - 1;
        unsigned long long base = 0ull;
        if (dest >= 0 && lane == leader) base = atomicAdd(&counts[sp.n_parts + dest], (unsigned long long)__popc(grp));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        const uint64_t at = base + __popc(grp & ((1u << lane) - 1u));
        if (dest >= 0 && at < m.cap_h) {
            HistRow h;
            h.jrow = slot_jrow[slot]; h.offset = (uint16_t)((e >> 16) & 0xFFFFull); h.cnt = (uint16_t)(e & 0xFFFFull);
            mail_hist_rows(m, mail_slot(m, dest, m.me))[at] = h;
        }
    }
}

__global__ void k_mail_publish(Mailboxes m, const unsigned long long *counts) {
    const int d = threadIdx.x;
    if (d >= m.world) return;
    unsigned long long *hdr = reinterpret_cast<unsigned long long *>(mail_slot(m, d, m.me));
    hdr[0] = counts[d];
    hdr[1] = counts[m.world + d];
}

void launch_mail_export(const Tables &t, const uint32_t *name_rank, const Splitters &sp, const Mailboxes &m, unsigned long long *counts,
                        uint32_t *slot_jrow, int sm_count, cudaStream_t st) {
    k_mail_slots<<<sm_count * 8, 256, 0, st>>>(t, name_rank, sp, m, counts, slot_jrow);
    k_mail_olog<<<sm_count * 8, 256, 0, st>>>(t, name_rank, sp, m, counts, slot_jrow);
    k_mail_publish<<<1, MAX_PARTS, 0, st>>>(m, counts);
}

/* fold the slot of one source out of this rank's own mailbox; the row counts come from the slot's header */
__global__ void __launch_bounds__(256) k_mail_import_junctions(Tables t, Mailboxes m, int source, uint32_t *slot_of) {
    char *box = mail_slot(m, m.me, source);
    const unsigned long long n_sent = *reinterpret_cast<const unsigned long long *>(box);
    if (n_sent > m.cap_j || reinterpret_cast<const unsigned long long *>(box)[1] > m.cap_h) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(t.overflow, 4u);
        return;
    }
    const JunctionRow *jr = mail_junction_rows(box);
    const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sent; i += n_threads) {
        const JunctionRow r = jr[i];
        const uint64_t key = make_key(t.kl, r.tid, r.i0, r.i1);
        const uint64_t slot = junction_find_or_insert(t, key);
        slot_of[i] = slot <= t.cap_mask ? (uint32_t)slot : 0xFFFFFFFFu;
        if (slot > t.cap_mask) continue;
        JStat *s = &t.stat[slot];
        for (int c = 0; c < 3; ++c) {
            if (r.first[c] != EMPTY64) atomicMin(&s->first[c], r.first[c]);
            if (r.cnt[c]) atomicAdd(&s->cnt[c], r.cnt[c]);
        }
        atomicMin(&s->a0, r.a0);
        atomicMax(&s->a1, r.a1);
    }
}

__global__ void __launch_bounds__(256) k_mail_import_hist(Tables t, Mailboxes m, int source, const uint32_t *slot_of) {
    char *box = mail_slot(m, m.me, source);
    const unsigned long long nj = reinterpret_cast<const unsigned long long *>(box)[0];
    const unsigned long long n_sent = reinterpret_cast<const unsigned long long *>(box)[1];
    if (nj > m.cap_j || n_sent > m.cap_h) return; /* reported by k_mail_import_junctions */
    const HistRow *hr = mail_hist_rows(m, box);
    const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n_sent + 31ull) & ~31ull;
    OlogWriter olog{0ull, 0u};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += n_threads) {
        uint32_t slot = 0xFFFFFFFFu, off = 0u, cnt = 0u;
        if (i < n_sent) {
            const HistRow r = hr[i];
            if (r.jrow < nj) slot = slot_of[r.jrow];
            off = r.offset;
            cnt = r.cnt;
        }
        const bool ok = slot != 0xFFFFFFFFu && cnt != 0u;
        const bool in_row = t.dense != nullptr && off < DENSE_W;
        if (ok && in_row) atomicAdd(&t.dense[(uint64_t)slot * DENSE_W + off], cnt);
        olog_append(t, olog, ok && !in_row, olog_entry(slot, off, cnt));
    }
    olog_finish(t, olog);
}

void launch_mail_import(const Tables &t, const Mailboxes &m, int source, uint32_t *slot_of, int sm_count, cudaStream_t st) {
    k_mail_import_junctions<<<sm_count * 4, 256, 0, st>>>(t, m, source, slot_of);
    k_mail_import_hist<<<sm_count * 8, 256, 0, st>>>(t, m, source, slot_of);
}

__global__ void k_export_thresholds(Tables t, ThresholdRow *rows, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = t.slot_list[i];
    const unsigned long long key = t.keys[slot];
    ThresholdRow r;
    r.tid = key_tid(t.kl, key);
    r.i0 = key_start(t.kl, key);
    r.i1 = r.i0 + key_len(t.kl, key);
    r.pad = 0u;
    r.T = t.stat[slot].T;
    rows[i] = r;
}

/* a key known here takes the earliest threshold any rank found (its own included) */
__global__ void k_import_thresholds(Tables t, const ThresholdRow *rows, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ThresholdRow r = rows[i];
    if (r.T == EMPTY64) return;
    const uint64_t slot = junction_find(t, make_key(t.kl, r.tid, r.i0, r.i1));
    if (slot <= t.cap_mask) atomicMin(&t.stat[slot].T, r.T);
}

void launch_export_thresholds(const Tables &t, ThresholdRow *rows, uint64_t n, cudaStream_t st) {
    if (n) k_export_thresholds<<<blocks_for(n, 256), 256, 0, st>>>(t, rows, n);
}
void launch_import_thresholds(const Tables &t, const ThresholdRow *rows, uint64_t n, cudaStream_t st) {
    if (n) k_import_thresholds<<<blocks_for(n, 256), 256, 0, st>>>(t, rows, n);
}

void launch_import_rows(const Tables &t, const JunctionRow *jr, uint64_t nj, const HistRow *hr, uint64_t nh, uint32_t *slot_of,
                        cudaStream_t st) {
    if (nj) k_import_junction_rows<<<blocks_for(nj, 256), 256, 0, st>>>(t, jr, nj, slot_of);
    if (nh) k_import_hist_rows<<<blocks_for(nh, 256), 256, 0, st>>>(t, hr, nh, slot_of, nj);
}

}  // namespace ipx
