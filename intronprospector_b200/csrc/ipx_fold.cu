/* ipx_fold.cu -- the junction table's kernels (sm_100a): initialisation / clear / rehash, K2 (keyed aggregation of the
 * junction events) and K3 (creation thresholds from the HBM genome).
 *
 *  k_fold<MODE>      K2: a warp takes the event regions of whole tiles, 32 events at a time.  match.any groups equal
 *                    keys; each group's leader looks its key up in the warp's write-back cache in shared memory
 *                    (FOLD_WAYS junctions: id, read counts, anchor extremes, first-seen values) and only on a miss
 *                    probes the hash table in HBM / L2.  Counts and extremes are accumulated in the cache and reach
 *                    the junction's 64-byte record once per residency (eviction or end of kernel) instead of once per
 *                    32 events; each read's offset is counted with one fire-and-forget atomic into the junction's
 *                    dense row, or staged in shared memory for the read-offset log (one reservation per LOG_STAGE
 *                    entries).  Events whose junction cannot be inserted because the table is full stay in the event
 *                    buffer (compacted to the front of their tile's region; tile_info.count = how many), so that the
 *                    same launches can simply be repeated after the host has grown the table.
 *                      MODE_SINGLE  no canon filter: insert + first-seen + counts + anchors + offsets
 *                      MODE_PASS1   canon filter:    insert + first-seen per strand class only
 *                      MODE_PASS2   canon filter:    count events at or after the creation threshold T
 *  k_resolve_T       K3: donor/acceptor gather from the HBM genome, per junction, and the choice of the
 *                    creating event T = first event whose (key, strand class) passes the filter
 *
 * Reference semantics: SURVEY.md section 8(a) A6-A8 (junctions_extractor.cc:405-530, junctions.hh:103-246).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipx_engine.h"
#include "ipx_motif.cuh"

namespace ipx {

/* ------------------------------------------------------------------ table initialisation */
__device__ __forceinline__ void reset_stat(JStat *s) {
    uint4 *p = reinterpret_cast<uint4 *>(s);
    p[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); /* first[0], first[1] */
    p[1] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); /* first[2], T */
    p[2] = make_uint4(0u, 0u, 0u, 0u);                                     /* cnt[0], cnt[1] */
    p[3] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);                            /* cnt[2], a0 = none, a1 = 0 */
}

__device__ __forceinline__ void reset_slot(Slot *s) {
    uint4 *p = reinterpret_cast<uint4 *>(s);
    p[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, JID_UNSET, BASE_UNSET); /* key, jid, base[0] */
    p[1] = make_uint4(BASE_UNSET, BASE_UNSET, 0u, 0u);                  /* base[1], base[2] */
}

/* hash slots [0, cap) */
__global__ void k_init_slots(Slot *slots, uint64_t cap) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) reset_slot(&slots[i]);
}

/* zero the dense rows of junction j that are in use: row 0, and the rows the slot's bases say were given out */
__device__ __forceinline__ void zero_dense_rows(const Tables &t, uint64_t j, const uint32_t base[DENSE_ROWS - 1], int lane) {
    if (!t.dense) return;
    uint4 *row = reinterpret_cast<uint4 *>(t.dense + j * (DENSE_ROWS * DENSE_W));
    row[lane] = make_uint4(0u, 0u, 0u, 0u);
    row[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r)
        if (base[r] != BASE_UNSET) {
            row[(r + 1u) * 64u + lane] = make_uint4(0u, 0u, 0u, 0u);
            row[(r + 1u) * 64u + lane + 32] = make_uint4(0u, 0u, 0u, 0u);
        }
}

/* junction rows [first, max_j): a warp per row (64-byte record, 1 KB offset row) */
__global__ void __launch_bounds__(256) k_init_rows(Tables t, uint64_t first) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = first + warp0; j < t.max_j; j += n_warps) {
        if (lane == 0) {
            reset_stat(&t.stat[j]);
            t.olog_cnt[j] = 0u;
            t.jslot[j] = SLOT_NONE;
        }
        if (t.dense) {
            uint4 *row = reinterpret_cast<uint4 *>(t.dense + j * (DENSE_ROWS * DENSE_W));
            for (uint32_t q = lane; q < DENSE_ROWS * DENSE_W / 4; q += 32u) row[q] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

/* <<<1, N_ID_POOLS>>> */
__global__ void k_reset_counters(Tables t) {
    t.id_pools[threadIdx.x] = (unsigned long long)ID_BLOCK; /* nothing to hand out: the first taker fetches a block */
    if (threadIdx.x != 0) return;
    *t.n_keys = 0ull;
    *t.n_olog = 0ull;
    *t.err_sv = EMPTY64;
    *t.n_unfolded = 0ull;
    *t.overflow = 0u;
}

/* clear(): undo exactly what is in use -- a warp per junction row (its hash slot, 64-byte record, 1 KB offset row);
 * the offset log is emptied by resetting its counter */
__global__ void __launch_bounds__(256) k_clear_tables(Tables t) {
    const uint64_t nk = junctions_in_use(t);
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t j = warp0; j < nk; j += n_warps) {
        const uint32_t slot = t.jslot[j];
        if (slot == SLOT_NONE) continue; /* an id no junction took: its row is in the reset state */
        Slot *s = &t.slots[slot];
        const uint32_t base[DENSE_ROWS - 1] = {s->base[0], s->base[1], s->base[2]};
        __syncwarp();
        zero_dense_rows(t, j, base, lane);
        if (lane == 0) {
            reset_slot(s);
            reset_stat(&t.stat[j]);
            t.olog_cnt[j] = 0u;
            t.jslot[j] = SLOT_NONE;
        }
    }
}

/* slots claimed while the rows were full (JID_FULL) are not undone by k_clear_tables: when the flag says there may be
 * some, sweep the slots (rare; otherwise the blocks leave at once) */
__global__ void k_sweep_full_slots(Tables t) {
    if (!(*((volatile uint32_t *)t.overflow) & OVF_TABLE)) return;
    const uint64_t cap = t.cap_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride)
        if (t.slots[i].jid == JID_FULL) reset_slot(&t.slots[i]);
}

void launch_init_tables(const Tables &t, cudaStream_t st) {
    k_init_slots<<<148 * 8, 256, 0, st>>>(t.slots, t.cap_mask + 1);
    k_init_rows<<<148 * 8, 256, 0, st>>>(t, 0);
    k_reset_counters<<<1, N_ID_POOLS, 0, st>>>(t);
}

void launch_init_rows(const Tables &t, uint64_t first, cudaStream_t st) { k_init_rows<<<148 * 8, 256, 0, st>>>(t, first); }

void launch_clear_tables(const Tables &t, cudaStream_t st) {
    k_clear_tables<<<148 * 4, 256, 0, st>>>(t);
    k_sweep_full_slots<<<148, 256, 0, st>>>(t);
    k_reset_counters<<<1, N_ID_POOLS, 0, st>>>(t);
}

/* Before the host grows anything: the counters that kept counting while a table was full go back to what exists,
 * and the flags of what is being repaired are lowered. */
__global__ void k_settle_counters(Tables t) {
    /* (every block of ids below max_j has been handed to a pool by the time one comes back short: the pools keep theirs) */
    *t.n_keys = junctions_in_use(t);
    *t.n_olog = log_entries_in_use(t);
    *t.n_unfolded = 0ull;
    atomicAnd(t.overflow, ~(OVF_TABLE | OVF_LOG));
}

/* rehash into a fresh slot array (already initialised): every junction row claims a slot for its key again and takes
 * the bases of its dense rows along; the rows themselves do not move */
__global__ void k_rehash(Tables t, const Slot *old_slots) {
    const uint64_t nk = junctions_in_use(t);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nk; j += stride) {
        const uint32_t was = t.jslot[j];
        if (was == SLOT_NONE) continue; /* an id no junction took */
        const unsigned long long key = t.jkey[j];
        uint64_t h = junction_hash(key, t.jtid[j]) & t.cap_mask;
        for (;;) {
            if (atomicCAS(&t.slots[h].key, EMPTY64, key) == EMPTY64) break;
            h = (h + 1) & t.cap_mask;
        }
        Slot *s = &t.slots[h];
        s->jid = (uint32_t)j;
        for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r) s->base[r] = old_slots[was].base[r];
        t.jslot[j] = (uint32_t)h;
    }
}

void launch_settle_counters(const Tables &t, cudaStream_t st) { k_settle_counters<<<1, 1, 0, st>>>(t); }
void launch_rehash(const Tables &t, const Slot *old_slots, cudaStream_t st) {
    k_init_slots<<<148 * 8, 256, 0, st>>>(t.slots, t.cap_mask + 1);
    k_rehash<<<148 * 4, 256, 0, st>>>(t, old_slots);
}

/* ------------------------------------------------------------------ K2: keyed aggregation */
__device__ __forceinline__ Event load_event(const Event *ev, uint64_t i, bool valid, int lane) {
    Event e;
    if (valid) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ev + i);
        const uint4 lo = __ldcs(src), hi = __ldcs(src + 1); /* streamed once: do not displace the tables in L2 */
        e.key = ((unsigned long long)lo.y << 32) | lo.x;
        e.sv = ((unsigned long long)lo.w << 32) | lo.z;
        e.a0 = hi.x; e.a1 = hi.y; e.off_cat = hi.z; e.tid = (int32_t)hi.w;
    } else {
        /* distinct keys that no real event has (bit 63 is never set in a key), so idle lanes group with nobody */
        e.key = EMPTY64 - 2 - (unsigned)lane; e.sv = EMPTY64; e.a0 = 0xFFFFFFFFu; e.a1 = 0u; e.off_cat = 0u; e.tid = 0;
    }
    return e;
}

__device__ __forceinline__ void store_event(Event *ev, uint64_t i, const Event &e) {
    uint4 *dst = reinterpret_cast<uint4 *>(ev + i);
    dst[0] = make_uint4((uint32_t)e.key, (uint32_t)(e.key >> 32), (uint32_t)e.sv, (uint32_t)(e.sv >> 32));
    dst[1] = make_uint4(e.a0, e.a1, e.off_cat, (uint32_t)e.tid);
}

/* The warp's write-back cache: FOLD_WAYS junctions, direct mapped, structure of arrays in shared memory.  An entry is
 * written by one lane at a time (the group leader that owns the way in this chunk) except for the extremes, which any
 * lane of the group may tighten with a shared-memory atomic after a plain read said it would improve them -- after a
 * junction's first chunk almost none does. */
#ifndef IPX_K2_BLOCKS
#define IPX_K2_BLOCKS 4 /* resident blocks per SM */
#endif
static const int FOLD_WARPS = 8;
#ifndef IPX_FOLD_WAYS
#define IPX_FOLD_WAYS 64
#endif
static const int FOLD_WAYS = IPX_FOLD_WAYS;

struct FoldCache {
    unsigned long long key[FOLD_WAYS];
    unsigned long long first[3][FOLD_WAYS]; /* MODE_PASS2: first[0] holds the junction's threshold T */
    uint32_t jid[FOLD_WAYS];
    int32_t tid[FOLD_WAYS];
    uint32_t cnt[3][FOLD_WAYS];
    uint32_t a0[FOLD_WAYS], a1[FOLD_WAYS];
    uint32_t slot[FOLD_WAYS];                   /* the junction's hash slot: the home of its row bases */
    uint32_t base[DENSE_ROWS - 1][FOLD_WAYS];   /* first offsets of the junction's dense rows 1 .. (copy of the slot's) */
    unsigned long long stage[LOG_STAGE];
};

/* write an entry's accumulated values into its junction's record */
template <int MODE>
__device__ __forceinline__ void fold_write_back(const Tables &t, FoldCache &c, int w) {
    JStat *s = &t.stat[c.jid[w]];
    if (MODE != MODE_PASS1) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (c.cnt[k][w]) atomicAdd(&s->cnt[k], (unsigned long long)c.cnt[k][w]);
        if (c.a0[w] != 0xFFFFFFFFu) atomicMin(&s->a0, c.a0[w]);
        if (c.a1[w] != 0u) atomicMax(&s->a1, c.a1[w]);
    }
    if (MODE != MODE_PASS2) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (c.first[k][w] != EMPTY64) atomicMin(&s->first[k], c.first[k][w]);
    }
}

template <int MODE>
__global__ void __launch_bounds__(FOLD_WARPS * 32, IPX_K2_BLOCKS) k_fold(Tables t, Event *ev, TileInfo *tile_info, uint64_t n_tiles) {
    __shared__ FoldCache s_cache[FOLD_WARPS];
    __shared__ IdPool s_pool;
    if (threadIdx.x == 0) id_pool_load(t, &s_pool);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    FoldCache &c = s_cache[threadIdx.x >> 5];
    c.key[lane] = EMPTY64;
    c.key[lane + 32] = EMPTY64;
    __syncwarp();
    LogStage log{c.stage, 0u};
    uint32_t spare = JID_NONE; /* see junction_probe */
    unsigned long long unfolded = 0ull;
    /* Tiles are handed out one at a time from a counter, in stream order: a warp that drew short tiles takes more of
     * them (with a fixed assignment a quarter of the warp time was spent waiting for the slowest warp at the end), and
     * the tiles in flight stay neighbours, whose junctions share the lines of the table that are hot in L2.  The next
     * tile is claimed, and its info requested, while the current one is folded. */
    auto claim = [&]() {
        unsigned long long v = 0ull;
        if (lane == 0) v = atomicAdd(t.next_tile, 1ull);
        return (uint64_t)__shfl_sync(0xFFFFFFFFu, v, 0);
    };
    uint64_t tile = claim(), tile_next = n_tiles;
    uint2 next = make_uint2(0u, 0u);
    if (tile < n_tiles) {
        next = __ldg(reinterpret_cast<const uint2 *>(tile_info + tile));
        tile_next = claim();
    }
    for (; tile < n_tiles; tile = tile_next, tile_next = tile_next < n_tiles ? claim() : n_tiles) {
        const uint32_t cnt = next.x;
        const uint64_t base = tile_event_base(next.y);
        if (tile_next < n_tiles) next = __ldg(reinterpret_cast<const uint2 *>(tile_info + tile_next));
        if (cnt == 0u) continue;
        uint32_t kept = 0u; /* events of this tile left unfolded (table full) */
        Event nxt = load_event(ev, base + lane, (uint32_t)lane < cnt, lane);
        for (uint32_t k0 = 0; k0 < cnt; k0 += 32u) {
            const Event e = nxt;
            const bool valid = k0 + lane < cnt;
            if (k0 + 32u < cnt) nxt = load_event(ev, base + k0 + 32u + lane, k0 + 32u + lane < cnt, lane);
            const uint32_t cat = (e.off_cat >> 16) & 3u;
            const int cls = (int)((e.off_cat >> 18) & 3u);
            unsigned grp = __match_any_sync(0xFFFFFFFFu, e.key);
            if (t.kl.wide) grp &= __match_any_sync(0xFFFFFFFFu, e.tid);
            const int leader = __ffs(grp) - 1;
            const bool is_leader = valid && lane == leader;
            const int way = (int)((junction_hash(e.key, e.tid) >> 40) & (FOLD_WAYS - 1));
            /* ---- the group's junction: from the cache, else from the table ---- */
            uint32_t jid = JID_NONE, gslot = 0u;
            uint32_t rb[DENSE_ROWS - 1] = {BASE_UNSET, BASE_UNSET, BASE_UNSET}; /* bases of the junction's dense rows 1 .. */
            unsigned long long T = EMPTY64;
            bool cached = false;
            const unsigned leaders = __ballot_sync(0xFFFFFFFFu, is_leader);
            if (is_leader) {
                /* two leaders of one chunk that map to the same way: the lower lane gets it, the other works uncached */
                const unsigned same = __match_any_sync(leaders, way);
                const bool owner = lane == __ffs(same) - 1;
                bool hit = false;
                if (owner) {
                    const unsigned long long ck = c.key[way];
                    if (ck == e.key && (!t.kl.wide || c.tid[way] == e.tid)) {
                        hit = cached = true;
                        jid = c.jid[way];
                        gslot = c.slot[way];
                        rb[0] = c.base[0][way]; rb[1] = c.base[1][way]; rb[2] = c.base[2][way];
                        if (MODE == MODE_PASS2) T = c.first[0][way];
                    } else if (ck != EMPTY64) {
                        fold_write_back<MODE>(t, c, way);
                        c.key[way] = EMPTY64;
                    }
                }
                if (!hit) {
                    /* one load of the junction's slot: its id and the bases of its dense rows */
                    const Found f = junction_probe<MODE != MODE_PASS2>(t, &s_pool, e.key, e.tid, spare);
                    jid = f.jid;
                    gslot = f.slot;
                    rb[0] = f.base[0]; rb[1] = f.base[1]; rb[2] = f.base[2];
                    if (MODE == MODE_PASS2 && jid != JID_NONE) T = *((volatile unsigned long long *)&t.stat[jid].T);
                    if (owner && jid != JID_NONE) {
                        c.key[way] = e.key;
                        c.tid[way] = e.tid;
                        c.jid[way] = jid;
                        c.slot[way] = gslot;
                        c.base[0][way] = rb[0]; c.base[1][way] = rb[1]; c.base[2][way] = rb[2];
                        c.cnt[0][way] = c.cnt[1][way] = c.cnt[2][way] = 0u;
                        c.a0[way] = 0xFFFFFFFFu;
                        c.a1[way] = 0u;
                        c.first[0][way] = MODE == MODE_PASS2 ? T : EMPTY64;
                        c.first[1][way] = c.first[2][way] = EMPTY64;
                        cached = true;
                    }
                }
            }
            __syncwarp();
            jid = __shfl_sync(0xFFFFFFFFu, jid, leader);
            cached = __shfl_sync(0xFFFFFFFFu, (int)cached, leader) != 0;
            const int gway = __shfl_sync(0xFFFFFFFFu, way, leader);
            const bool found = valid && jid != JID_NONE;
            /* An event whose junction has no row because the table is full stays in the buffer for the repeat of the
             * launch.  MODE_PASS1 keeps nothing: it is idempotent (minima), and MODE_PASS2 needs the events and the
             * tiles' counts as they are; a key the first pass could not insert is absent in the second, which keeps
             * its events. */
            const bool keep = MODE != MODE_PASS1 && valid && !found;
            bool counted = found;
            if (MODE == MODE_PASS2) {
                T = __shfl_sync(0xFFFFFFFFu, T, leader);
                counted = found && T != EMPTY64 && e.sv >= T; /* earlier events are dropped */
            }
            JStat *s = &t.stat[found ? jid : 0u];
            if (MODE != MODE_PASS2 && found) {
                /* first-seen sequence value of this lane's strand class */
                if (cached) {
                    if (e.sv < c.first[cls][gway]) atomicMin(&c.first[cls][gway], e.sv);
                } else {
                    stat_first(s, cls, e.sv);
                }
            }
            if (MODE != MODE_PASS1) {
                const unsigned c0 = __ballot_sync(0xFFFFFFFFu, counted && cat == 0u) & grp;
                const unsigned c1 = __ballot_sync(0xFFFFFFFFu, counted && cat == 1u) & grp;
                const unsigned c2 = __ballot_sync(0xFFFFFFFFu, counted && cat == 2u) & grp;
                if (lane == leader && found) {
                    if (cached) {
                        c.cnt[0][gway] += (uint32_t)__popc(c0);
                        c.cnt[1][gway] += (uint32_t)__popc(c1);
                        c.cnt[2][gway] += (uint32_t)__popc(c2);
                    } else {
                        if (c0) atomicAdd(&s->cnt[0], (unsigned long long)__popc(c0));
                        if (c1) atomicAdd(&s->cnt[1], (unsigned long long)__popc(c1));
                        if (c2) atomicAdd(&s->cnt[2], (unsigned long long)__popc(c2));
                    }
                }
                const uint32_t off = e.off_cat & 0xFFFFu;
                /* row 0 takes the offsets below DENSE_W: nothing else to look at */
                uint32_t place = (off < DENSE_W && t.dense != nullptr) ? off : DENSE_NONE;
                const bool later = counted && off >= DENSE_W && t.dense != nullptr && t.extra_rows;
                if (__any_sync(0xFFFFFFFFu, later)) {
                    /* a later junction of its read: the junction's extra rows (their bases travel from the leader) */
                    gslot = __shfl_sync(0xFFFFFFFFu, gslot, leader);
                    rb[0] = __shfl_sync(0xFFFFFFFFu, rb[0], leader);
                    rb[1] = __shfl_sync(0xFFFFFFFFu, rb[1], leader);
                    rb[2] = __shfl_sync(0xFFFFFFFFu, rb[2], leader);
                    if (later) {
                        const uint32_t was[DENSE_ROWS - 1] = {rb[0], rb[1], rb[2]};
                        place = dense_place(t, gslot, rb, off);
                        if (cached) {
                            /* rows given out just now: later chunks find them in the cache */
#pragma unroll
                            for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r)
                                if (rb[r] != was[r]) c.base[r][gway] = rb[r];
                        }
                    }
                }
                if (counted) {
                    if (cached) {
                        if (e.a0 < c.a0[gway]) atomicMin(&c.a0[gway], e.a0);
                        if (e.a1 > c.a1[gway]) atomicMax(&c.a1[gway], e.a1);
                    } else {
                        stat_anchor(s, e.a0, e.a1);
                    }
                    /* Junction::count_read's push_back (junctions.hh:163-167), run-length form */
                    if (place != DENSE_NONE) atomicAdd(&t.dense[(uint64_t)jid * (DENSE_ROWS * DENSE_W) + place], 1u);
                }
                log_append(t, log, counted && place == DENSE_NONE, olog_entry(jid, off, 1u));
            }
            /* events that could not be folded go back to the front of the region (behind what was read already) */
            const unsigned km = __ballot_sync(0xFFFFFFFFu, keep);
            if (km) {
                if (keep) store_event(ev, base + kept + (uint32_t)__popc(km & ((1u << lane) - 1u)), e);
                kept += (uint32_t)__popc(km);
            }
            __syncwarp();
        }
        if (MODE != MODE_PASS1 && lane == 0) tile_info[tile].count = kept;
        unfolded += kept;
    }
    /* end of kernel: everything the cache still holds, and the staged log entries */
#pragma unroll
    for (int w = lane; w < FOLD_WAYS; w += 32)
        if (c.key[w] != EMPTY64) fold_write_back<MODE>(t, c, w);
    if (MODE != MODE_PASS1) log_flush(t, log);
    if (lane == 0 && unfolded) atomicAdd(t.n_unfolded, unfolded);
    __syncthreads();
    if (threadIdx.x == 0) id_pool_store(t, &s_pool);
}

void launch_fold(int mode, const Tables &t, Event *ev, TileInfo *tile_info, uint64_t n_tiles, int sm_count, cudaStream_t st) {
    if (n_tiles == 0) return;
    cudaMemsetAsync(t.next_tile, 0, sizeof(unsigned long long), st);
    const uint64_t want = (n_tiles + FOLD_WARPS - 1) / FOLD_WARPS;
    const uint64_t max_blocks = (uint64_t)sm_count * IPX_K2_BLOCKS; /* what is resident at once */
    const int blocks = (int)(want < max_blocks ? want : max_blocks);
    switch (mode) {
        case MODE_SINGLE: k_fold<MODE_SINGLE><<<blocks, FOLD_WARPS * 32, 0, st>>>(t, ev, tile_info, n_tiles); break;
        case MODE_PASS1: k_fold<MODE_PASS1><<<blocks, FOLD_WARPS * 32, 0, st>>>(t, ev, tile_info, n_tiles); break;
        default: k_fold<MODE_PASS2><<<blocks, FOLD_WARPS * 32, 0, st>>>(t, ev, tile_info, n_tiles); break;
    }
}

/* ------------------------------------------------------------------ K3: creation thresholds */
/* For every junction without a creation threshold yet: T = first event (stream order) whose strand class
 * passes junction_filter_check (add_junction, junctions_extractor.cc:494-502; SURVEY.md 8a "A6"). */
__global__ void k_resolve_T(Tables t, GenomeView g) {
    const uint64_t nk = junctions_in_use(t);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nk; j += stride) {
        JStat *s = &t.stat[j];
        if (s->T != EMPTY64) continue;
        const unsigned long long f0 = s->first[CLS_PLUS], f1 = s->first[CLS_MINUS], f2 = s->first[CLS_DOT];
        bool p0 = true, p1 = true, p2 = true;
        const unsigned long long key = t.jkey[j];
        const int32_t tid = t.jtid[j];
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

void launch_resolve_T(const Tables &t, const GenomeView &g, cudaStream_t st) { k_resolve_T<<<148 * 4, 256, 0, st>>>(t, g); }

}  // namespace ipx
