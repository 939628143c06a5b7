/* ipx_extract.cu -- K1, the extraction kernels (sm_100a), and the expansion of compact batches.
 *
 *  k_extract_tma     the hot kernel: persistent blocks (two per SM: a producer warp + 15 worker warps) stream
 *                    tiles of EX_TILE (2048) consecutive records through a 2-stage shared-memory ring filled by TMA bulk
 *                    copies (cp.async.bulk + mbarrier), so every record column and CIGAR word is read from HBM
 *                    exactly once, coalesced, with the latency hidden behind the previous tiles' work.  The workers
 *                    (a) classify the tile's records and compact the live ones (process_alignment's skip test) into
 *                    two lists -- CIGARs of at most three operations, which can only be "aligned, N, aligned", and
 *                    longer ones -- then (b) walk the listed records with full warps (ipx_walk.cuh) and write each
 *                    junction event straight into the tile's private region of the event buffer (the emitting lanes
 *                    of a chunk reserve their slots with one shared-memory atomic).
 *  k_extract_sorted  K1 for long reads: records taken longest first, lean 32-bit automaton, 32-byte CIGAR groups;
 *                    k_seg_* walk outliers (tens of thousands of operations) segment by segment.  Long-read batches
 *                    are cut into tiles of LONG_TILE (256) records: a tile is K2's unit of work.
 *  k_extract_simple  K1 for what is left: the last partial tile, columns that are not 16-byte aligned.
 *
 * Reference semantics: SURVEY.md section 8(a) A1-A5 (junctions_extractor.cc:49-95,124-319,322-401,533-549).
 */
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "ipx_engine.h"
#include "ipx_scan.cuh"

namespace ipx {

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
    if (!tmp) {
        *tmp_bytes = scan_state_words(n + 1) * sizeof(unsigned long long);
        return cudaSuccess;
    }
    /* cigar_off[0 .. n] = exclusive sum of the operation counts (the item after the last record counts 0) */
    cudaError_t e = scan_exclusive(OpsOrZero{n_ops, n}, cigar_off, n + 1, (unsigned long long *)tmp, st);
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
            atomicOr(t.overflow, OVF_START); /* outside what a BAM position can hold */
            return;
        }
        const unsigned long long key = make_key(t.kl, tid, c.intron_start, c.intron_end);
        const unsigned long long sv = make_sv(record_seq, c.ordinal, strand);
        const uint32_t off_cat = ((uint32_t)(c.intron_start - pos) & 0xFFFFu) | (cat << 16) | ((uint32_t)cls << 18);
        const uint32_t idx = atomicAdd(region_fill, 1u);
        /* one 32-byte event = one DRAM sector: two 16-byte stores from the emitting lane */
        uint4 *dst = reinterpret_cast<uint4 *>(region + idx);
        __stcs(dst, make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)sv, (uint32_t)(sv >> 32)));
        __stcs(dst + 1, make_uint4((uint32_t)c.anchor_start, (uint32_t)c.anchor_end, off_cat, (uint32_t)tid));
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
    if (i % LONG_TILE == 0) tile_info[i / LONG_TILE] = TileInfo{0u, c0, 0u, 0u};
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
        const uint64_t tile = r / LONG_TILE;
        Event *region = out + tile_event_base(__ldg(&b.cigar_off[tile * LONG_TILE]));
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
static const int EX_STAGES = 2;
static const int EX_BLOCKS = 2; /* resident blocks per SM */
static const int EX_WORKER_WARPS = 15;
static const int EX_LOOKAHEAD = 1; /* workers classify up to this many tiles ahead of the one they walk */
static const int EX_ROUNDS = EX_TILE / 128;
static const int EX_CIG_CAP = 2 * EX_TILE; /* staged CIGAR words per tile (short reads average 1.5 per record); words beyond are
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
/* try_wait parks the thread on the barrier until the phase completes or the hint runs out */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try(bar, parity, 1000000u)) {
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
__device__ __forceinline__ void classify_round(ExStage &S, uint32_t r, int lane, unsigned long long &live_words) {
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
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t rl = 4 * q + j;
        if (kind[j] == 1u) {
            S.list[ps++] = (uint16_t)rl;
        } else if (kind[j] == 2u) {
            S.list[EX_TILE - 1 - (pl++)] = (uint16_t)rl;
        }
    }
}

/* take classification rounds of a stage until none is left */
__device__ __forceinline__ void take_rounds(ExStage &S, int stage, uint64_t *classified, int lane, unsigned long long &live_words) {
    for (;;) {
        uint32_t r = 0u;
        /* (the plain read spares the atomic when nothing is left; the decision stays lane 0's, so it is warp-uniform) */
        if (lane == 0) r = *((volatile uint32_t *)&S.next_round) >= (uint32_t)EX_ROUNDS ? 0xFFFFFFFFu : atomicAdd(&S.next_round, 1u);
        r = __shfl_sync(0xFFFFFFFFu, r, 0);
        if (r >= (uint32_t)EX_ROUNDS) break;
        classify_round(S, r, lane, live_words);
        __syncwarp();
        if (lane == 0) mbar_arrive(&classified[stage]); /* release: the lists are visible to whoever sees the phase complete */
    }
}

/* the short-CIGAR chunk: "aligned, N, aligned" is the only shape of two or three operations that emits a junction
 * (ReadJunctionExtractor, junctions_extractor.cc:208-254,278-282).  Called by all 32 lanes of a chunk (`valid`: the lane
 * has a record): the lanes that emit reserve their slots of the tile's region together, with one shared-memory atomic
 * per chunk instead of one per event (32 lanes adding to one address are served one after the other). */
__device__ __forceinline__ void walk_short(const BatchView &b, const ExtractParams &p, const Tables &t, ExStage &S, Event *region,
                                           bool valid, uint32_t rl, int lane, unsigned long long &events) {
    bool emit = false, general = false;
    uint32_t pos = 0u, i0 = 0u, l0 = 0u, l1 = 0u, l2 = 0u, flag = 0u, nh = 0u;
    if (valid) {
        const uint32_t c0 = S.off[rl], n_cigar = S.off[rl + 1] - c0;
        const uint32_t w0 = cigar_word(S, b.cigar, c0), w1 = cigar_word(S, b.cigar, c0 + 1);
        const uint32_t w2 = n_cigar == 3u ? cigar_word(S, b.cigar, c0 + 2) : 0u;
        const uint32_t o0 = w0 & 15u, o1 = w1 & 15u, o2 = w2 & 15u;
        const uint32_t known = 0x1BFu, aligned = 0x181u; /* MIDNSH=X ; M=X */
        if (!((known >> o0) & (known >> o1) & (known >> o2) & 1u)) {
            general = true; /* the full automaton reports the unknown operation exactly */
        } else if (n_cigar == 3u && o1 == OP_N && ((aligned >> o0) & (aligned >> o2) & 1u)) {
            /* the one candidate: anchor_start = pos, intron = [pos + l0, + l1), anchor_end = intron_end + l2, mismatches =
             * the X lengths, no indels.  Operation lengths are below 2^28 and pos is a BAM int32, so 32-bit arithmetic
             * gives what the reference's 64-bit arithmetic gives wherever the result is usable (a start outside
             * [0, 2^31) is reported). */
            l0 = w0 >> 4; l1 = w1 >> 4; l2 = w2 >> 4;
            flag = S.flag[rl]; nh = S.nh[rl];
            const uint32_t lbad = o0 == OP_X ? l0 : 0u, rbad = o2 == OP_X ? l2 : 0u;
            /* junction_qc (junctions_extractor.cc:384-401) */
            if (read_may_emit(p, flag, nh) && l1 >= p.min_intron_length && l1 <= p.max_intron_length &&
                l0 - lbad >= p.min_anchor_length && l2 - rbad >= p.min_anchor_length) {
                events++;
                pos = (uint32_t)S.pos[rl];
                i0 = pos + l0;
                if (i0 > 0x7FFFFFFFu) atomicOr(t.overflow, OVF_START); else emit = true;
            }
        }
    }
    const unsigned em = __ballot_sync(0xFFFFFFFFu, emit);
    if (em) {
        const int leader = __ffs(em) - 1;
        uint32_t idx = 0u;
        if (lane == leader) idx = atomicAdd(&S.region_fill, (uint32_t)__popc(em));
        idx = __shfl_sync(0xFFFFFFFFu, idx, leader) + (uint32_t)__popc(em & ((1u << lane) - 1u));
        if (emit) {
            const char strand = read_strand(p.strandness, flag, S.xs[rl]);
            const uint32_t tid = (uint32_t)S.tid[rl];
            const unsigned long long key = ((unsigned long long)(tid & t.kl.tid_mask) << t.kl.tid_shift) |
                                           ((unsigned long long)i0 << t.kl.start_shift) | l1;
            const unsigned long long sv = make_sv(b.seq_base + (uint64_t)S.tile * EX_TILE + rl, 0u, strand);
            const uint32_t off_cat = (l0 & 0xFFFFu) | (read_category(flag, nh) << 16) | ((uint32_t)strand_class(strand) << 18);
            uint4 *dst = reinterpret_cast<uint4 *>(region + idx);
            __stcs(dst, make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)sv, (uint32_t)(sv >> 32)));
            __stcs(dst + 1, make_uint4(pos, i0 + l1 + l2, off_cat, tid));
        }
    }
    if (general) walk_record(b, p, t, S, region, rl, events);
}

template <int LOOKAHEAD, int WORKERS, int STAGES, int BLOCKS>
__global__ void __launch_bounds__((WORKERS + 1) * 32, BLOCKS) k_extract_tma(BatchView b, uint32_t n_cigar_words, ExtractParams p, Tables t,
                                                               Event *out, TileInfo *tile_info, uint32_t n_full_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ExStage *stages = reinterpret_cast<ExStage *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + STAGES * sizeof(ExStage)); /* TMA bytes have landed */
    uint64_t *empty = full + STAGES;                                                    /* every worker has left */
    uint64_t *classified = empty + STAGES;                                              /* all rounds are done */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
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
        for (uint32_t it = 0; it < my_tiles + STAGES; ++it) {
            const int s = (int)(it % STAGES);
            ExStage &S = stages[s];
            if (it >= (uint32_t)STAGES) {
                /* the tile that occupied this stage has been left by every worker */
                mbar_wait(&empty[s], ((it / STAGES) & 1u) ^ 1u);
                tile_info[S.tile] = TileInfo{S.region_fill, S.off[0], 0u, 0u};
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
        const int s = (int)(it % STAGES);
        ExStage &S = stages[s];
        /* (1) rounds of the tiles ahead whose data is already here, nearest first */
#pragma unroll
        for (uint32_t d = 1; d <= (uint32_t)LOOKAHEAD; ++d) {
            if (it + d < my_tiles) {
                const int s1 = (int)((it + d) % STAGES);
                if (mbar_test(&full[s1], ((it + d) / STAGES) & 1u))
                    take_rounds(stages[s1], s1, classified, lane, live_words);
            }
        }
        /* (2) this tile must be classified before it can be walked: help with what is left, then wait for the rest */
        mbar_wait(&full[s], (it / STAGES) & 1u);
        take_rounds(S, s, classified, lane, live_words);
        mbar_wait(&classified[s], (it / STAGES) & 1u);
        /* (3) chunks of this tile: 32 listed records at a time, longest CIGARs first */
        const uint32_t n_lists = *((volatile uint32_t *)&S.n_lists);
        const uint32_t n_short = n_lists & 0xFFFFu, n_long = n_lists >> 16;
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
            walk_short(b, p, t, S, region, c + lane < n_short, c + lane < n_short ? S.list[c + lane] : 0u, lane, events);
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

static bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

cudaError_t launch_extract(const BatchView &b, uint64_t n_cigar, const ExtractParams &p, const Tables &t, Event *out,
                           TileInfo *tile_info, unsigned long long *next_record, int sm_count, cudaStream_t st, int *launches,
                           uint64_t *n_tiles_out) {
    if (launches) *launches = 0;
    if (n_tiles_out) *n_tiles_out = tiles_of(b.n);
    if (b.n == 0) return cudaSuccess;
    const uint64_t n_tiles = tiles_of(b.n), n_full = b.n / EX_TILE;
    /* the pipelined kernel takes whole tiles of batches whose columns can be bulk-copied and whose reads are short
     * enough for a tile's CIGAR words to fit its stage */
    const bool pipelined = n_full > 0 && n_cigar <= 8 * b.n && aligned16(b.tid) && aligned16(b.pos) && aligned16(b.flag) &&
                           aligned16(b.nh) && aligned16(b.xs) && aligned16(b.cigar_off) && aligned16(b.cigar);
    uint64_t first_simple = 0;
    if (pipelined) {
        /* measured on B200 (100 M short reads): 2048-record tiles with 15 workers, a 2-stage ring and
         * 2 blocks per SM 0.644 ms; 11 / 13 workers 0.666 / 0.649; one block of 31 workers with 3 / 4 stages 0.756 / 0.686;
         * 1024-record tiles: 15 workers x 4 stages x 2 blocks 0.716, 7 x 2 x 4 0.668, 5 x 2 x 4 0.727; look-ahead of
         * 2 / 3 tiles and named barriers were slower still */
        auto kernel = k_extract_tma<EX_LOOKAHEAD, EX_WORKER_WARPS, EX_STAGES, EX_BLOCKS>;
        const size_t smem = EX_STAGES * sizeof(ExStage) + 3 * EX_STAGES * sizeof(uint64_t);
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); /* per device: repeated */
        if (e != cudaSuccess) return e;
        const uint64_t max_blocks = (uint64_t)sm_count * EX_BLOCKS;
        const int blocks = (int)(n_full < max_blocks ? n_full : max_blocks);
        kernel<<<blocks, (EX_WORKER_WARPS + 1) * 32, smem, st>>>(b, (uint32_t)n_cigar, p, t, out, tile_info, (uint32_t)n_full);
        if (launches) (*launches)++;
        first_simple = n_full;
    }
    /* long reads (more than 8 operations per record on average): longest first, lean automaton; the longest of them
     * segment by segment */
    if (first_simple == 0 && n_cigar > 8 * b.n && aligned16(b.cigar) && next_record) {
        /* Records longer than this are walked segment by segment.  That is twice the work per operation, and on
         * log-normal read lengths (longest read 12x the mean) it measured slower at every threshold (B200, 1 M reads:
         * 3.6 / 2.9 / 2.6 / 2.6 / 2.4 ms for 1536 / 2048 / 3072 / 4096 / 6144 against 2.3 ms without): the kernel is
         * bound by throughput, not by its longest record.  So the default only catches outliers that would be -- a
         * single lane needs ~0.3 us per operation.  (Read per call: the tests lower it.) */
        const char *seg_env = getenv("IPX_K1_SEG_MIN");
        const uint32_t seg_min_ops = seg_env ? (uint32_t)atoi(seg_env) : 16384u;
        const bool segmented = seg_min_ops != 0u;
        if (n_tiles_out) *n_tiles_out = max_tiles_of(b.n); /* tiles of LONG_TILE records */
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

}  // namespace ipx
