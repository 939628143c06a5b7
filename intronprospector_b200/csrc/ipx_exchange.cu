/* ipx_exchange.cu -- multi-GPU exchange of partial junction tables by key range (sm_100a).
 *
 *  k_mail_*          the exchange over peer memory (NVLink / NVSwitch): every rank owns a mailbox with one slot per
 *                    source rank; the exporting kernels put the rows of the junctions another rank owns together in a
 *                    local outbox and disown them -- the rank's own share never moves -- k_mail_send copies every
 *                    outbox slot into its owner's mailbox (coalesced 16-byte stores over NVLink), and a last small
 *                    kernel publishes the row counts and the exchange's epoch into the slot headers; the importing
 *                    side starts with a one-warp kernel that waits for the epoch of every source (a device-side
 *                    barrier: no collective, no host round trip) and then folds all sources' slots in two launches.
 *  k_export_* / k_import_*   the same rows packed into send buffers for an all-to-all transport (hosts without peer
 *                    access; what the gloo test drives): everything is exported, the table cleared, every share imported.
 *  k_*_thresholds    canon filter across ranks: (key, creation threshold) rows
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipx_engine.h"

namespace ipx {

static inline unsigned blocks_for(uint64_t n, unsigned threads) {
    uint64_t b = (n + threads - 1) / threads;
    return (unsigned)(b ? b : 1);
}

__device__ __forceinline__ int part_of(const Splitters &sp, uint32_t name_rank, uint32_t i0) {
    int p = 0;
    while (p < sp.n_parts - 1 && (name_rank > sp.rank[p] || (name_rank == sp.rank[p] && i0 >= sp.start[p]))) p++;
    return p;
}

/* offset rows a counter needs: counts above 65535 are sent as several rows */
__device__ __forceinline__ uint32_t hist_rows_of(uint32_t v) { return v <= 0xFFFFu ? (v != 0u ? 1u : 0u) : (v - 1u) / 0xFFFFu + 1u; }

/* what a warp needs of one junction row: owner, key fields, the offset rows its dense counters turn into */
struct RowView {
    int dest; /* -1: nothing to export (never created: nothing was counted into it) */
    uint32_t i0, i1, nz; /* nz: offset rows of all its dense rows together */
    int32_t tid;
    uint32_t base[DENSE_ROWS - 1];
};

/* lane's 8 counters of dense row `row` of junction j */
__device__ __forceinline__ void load_counters(const Tables &t, uint64_t j, uint32_t row, int lane, uint32_t v[8]) {
    const uint4 *p = reinterpret_cast<const uint4 *>(t.dense + (j * DENSE_ROWS + row) * DENSE_W);
    const uint4 a = p[lane * 2], b = p[lane * 2 + 1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ RowView read_row(const Tables &t, const uint32_t *name_rank, const Splitters &sp, uint64_t j, int lane) {
    RowView r;
    r.dest = -1;
    r.nz = 0u;
    r.base[0] = r.base[1] = r.base[2] = BASE_UNSET;
    const unsigned long long T = t.stat[j].T;
    const unsigned long long key = t.jkey[j];
    r.tid = t.jtid[j];
    r.i0 = key_start(t.kl, key);
    r.i1 = r.i0 + key_len(t.kl, key);
    if (T == EMPTY64) return r; /* (a hole or a disowned row: warp-uniform) */
    r.dest = part_of(sp, name_rank[r.tid], r.i0);
    if (t.dense) {
        const Slot *s = &t.slots[t.jslot[j]];
        r.base[0] = s->base[0]; r.base[1] = s->base[1]; r.base[2] = s->base[2];
        uint32_t mine = 0u;
        /* (not unrolled, here and below: the export kernel spent 72 % of its stall samples waiting for instructions) */
#pragma unroll 1
        for (uint32_t row = 0; row < DENSE_ROWS; ++row) {
            if (row > 0u && (row == 1u ? r.base[0] : row == 2u ? r.base[1] : r.base[2]) == BASE_UNSET) continue;
            uint32_t v[8];
            load_counters(t, j, row, lane, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) mine += hist_rows_of(v[k]);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
        r.nz = mine;
    }
    return r;
}

__device__ __forceinline__ JunctionRow make_junction_row(const JStat &s, const RowView &r) {
    JunctionRow jr;
    jr.first[0] = s.first[0]; jr.first[1] = s.first[1]; jr.first[2] = s.first[2];
    jr.cnt[0] = s.cnt[0]; jr.cnt[1] = s.cnt[1]; jr.cnt[2] = s.cnt[2];
    jr.tid = r.tid; jr.i0 = r.i0; jr.i1 = r.i1; jr.a0 = s.a0; jr.a1 = s.a1; jr.pad = 0;
    return jr;
}

/* the non-zero dense counters of junction j as offset rows of junction row `jrow`, starting at hr[o] (r.nz of them);
 * ZERO: the counters are reset on the way (the junction is being disowned) */
template <bool ZERO>
__device__ __forceinline__ void write_dense_rows(const Tables &t, uint64_t j, HistRow *hr, uint64_t o, uint32_t jrow, const RowView &r,
                                                 int lane, bool write) {
#pragma unroll 1
    for (uint32_t row = 0; row < DENSE_ROWS; ++row) {
        const uint32_t row_base = row == 0u ? 0u : row == 1u ? r.base[0] : row == 2u ? r.base[1] : r.base[2];
        if (row > 0u && row_base == BASE_UNSET) continue;
        uint32_t v[8];
        load_counters(t, j, row, lane, v);
        uint32_t mine = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) mine += hist_rows_of(v[k]);
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += y;
        }
        const uint32_t first = row_base;
        uint64_t at = o + (incl - mine);
        if (write) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t left = v[k];
                while (left) {
                    const uint32_t piece = left > 0xFFFFu ? 0xFFFFu : left;
                    HistRow h;
                    h.jrow = jrow; h.offset = (uint16_t)(first + (uint32_t)(lane * 8 + k)); h.cnt = (uint16_t)piece;
                    hr[at++] = h;
                    left -= piece;
                }
            }
        }
        o += __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (ZERO) {
            uint4 *p = reinterpret_cast<uint4 *>(t.dense + (j * DENSE_ROWS + row) * DENSE_W);
            p[lane * 2] = make_uint4(0u, 0u, 0u, 0u);
            p[lane * 2 + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

/* ---- packed rows for an all-to-all transport ---- */
/* warp per junction row.  SCATTER = false: count rows per destination; true: write them behind the destination's cursor. */
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_export_rows(Tables t, const uint32_t *name_rank, Splitters sp, uint64_t n_keys,
                                                     unsigned long long *counts, uint32_t *jrow_of, JunctionRow *jr, HistRow *hr) {
    const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= n_keys || j >= junctions_in_use(t)) return; /* n_keys is a launch bound; the table knows how many there are */
    const RowView r = read_row(t, name_rank, sp, j, lane);
    if (r.dest < 0) return;
    unsigned long long jpos = 0ull, hpos = 0ull;
    if (lane == 0) {
        jpos = atomicAdd(&counts[r.dest], 1ull);
        if (r.nz) hpos = atomicAdd(&counts[sp.n_parts + r.dest], (unsigned long long)r.nz);
    }
    if (!SCATTER) return;
    jpos = __shfl_sync(0xFFFFFFFFu, jpos, 0);
    hpos = __shfl_sync(0xFFFFFFFFu, hpos, 0);
    const uint32_t jrow = (uint32_t)jpos - sp.jbase[r.dest];
    if (lane == 0) {
        jr[jpos] = make_junction_row(t.stat[j], r);
        jrow_of[j] = ((uint32_t)r.dest << 24) | jrow; /* for the log entries of this junction */
    }
    write_dense_rows<false>(t, j, hr, hpos, jrow, r, lane, true);
}

/* thread per entry of the read-offset log: one offset row each (equal pairs are summed by the importer) */
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_export_olog(Tables t, const uint32_t *name_rank, Splitters sp, uint64_t n_olog,
                                                     unsigned long long *counts, const uint32_t *jrow_of, HistRow *hr) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int dest = -1 - lane; /* idle lanes group with nobody */
    unsigned long long e = EMPTY64;
    uint32_t j = 0u;
    if (i < n_olog && i < log_entries_in_use(t)) {
        e = t.olog[i];
        j = (uint32_t)(e >> 32);
        if (j < junctions_in_use(t) && t.stat[j].T != EMPTY64)
            dest = part_of(sp, name_rank[t.jtid[j]], key_start(t.kl, t.jkey[j]));
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
        h.jrow = jrow_of[j] & 0xFFFFFFu; h.offset = (uint16_t)((e >> 16) & 0xFFFFull); h.cnt = (uint16_t)(e & 0xFFFFull);
        hr[base + __popc(grp & ((1u << lane) - 1u))] = h;
    }
}

void launch_export_count(const Tables &t, const uint32_t *name_rank, const Splitters &sp, uint64_t n_keys, uint64_t n_olog,
                         unsigned long long *counts, cudaStream_t st) {
    if (n_keys) k_export_rows<false><<<blocks_for(n_keys * 32, 256), 256, 0, st>>>(t, name_rank, sp, n_keys, counts, nullptr, nullptr, nullptr);
    if (n_olog) k_export_olog<false><<<blocks_for(n_olog, 256), 256, 0, st>>>(t, name_rank, sp, n_olog, counts, nullptr, nullptr);
}

void launch_export_scatter(const Tables &t, const uint32_t *name_rank, const Splitters &sp, uint64_t n_keys, uint64_t n_olog,
                           unsigned long long *cursors, uint32_t *jrow_of, JunctionRow *jr, HistRow *hr, cudaStream_t st) {
    if (n_keys) k_export_rows<true><<<blocks_for(n_keys * 32, 256), 256, 0, st>>>(t, name_rank, sp, n_keys, cursors, jrow_of, jr, hr);
    if (n_olog) k_export_olog<true><<<blocks_for(n_olog, 256), 256, 0, st>>>(t, name_rank, sp, n_olog, cursors, jrow_of, hr);
}

/* which dense counter of junction j takes an offset that arrives from another rank (rows are given out as in K2) */
__device__ __forceinline__ uint32_t place_offset(const Tables &t, uint32_t j, uint32_t off) {
    if (t.dense == nullptr) return DENSE_NONE;
    if (off < DENSE_W) return off;
    const uint32_t slot = t.jslot[j];
    const Slot *s = &t.slots[slot];
    uint32_t base[DENSE_ROWS - 1] = {*((volatile const uint32_t *)&s->base[0]), *((volatile const uint32_t *)&s->base[1]),
                                     *((volatile const uint32_t *)&s->base[2])};
    return dense_place(t, slot, base, off);
}

/* fold one received junction row; returns the junction id or JID_NONE */
__device__ __forceinline__ uint32_t import_junction_row(const Tables &t, IdPool *pool, const JunctionRow &r, uint32_t &spare) {
    const uint32_t j = junction_probe<true>(t, pool, make_key(t.kl, r.tid, r.i0, r.i1), r.tid, spare).jid;
    if (j == JID_NONE) return j;
    JStat *s = &t.stat[j];
    for (int c = 0; c < 3; ++c) {
        if (r.first[c] != EMPTY64) atomicMin(&s->first[c], r.first[c]);
        if (r.cnt[c]) atomicAdd(&s->cnt[c], r.cnt[c]);
    }
    atomicMin(&s->a0, r.a0);
    atomicMax(&s->a1, r.a1);
    /* the junction's creating event may now be one of another rank's: resolved again by the next k_resolve_T */
    *((volatile unsigned long long *)&s->T) = EMPTY64;
    return j;
}

__global__ void __launch_bounds__(256) k_import_junction_rows(Tables t, const JunctionRow *jr, uint64_t n, uint32_t *jid_of) {
    __shared__ IdPool s_pool;
    if (threadIdx.x == 0) id_pool_load(t, &s_pool);
    __syncthreads();
    const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
    uint32_t spare = JID_NONE;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += n_threads) jid_of[i] = import_junction_row(t, &s_pool, jr[i], spare);
    __syncthreads();
    if (threadIdx.x == 0) id_pool_store(t, &s_pool);
}

__global__ void __launch_bounds__(256) k_import_hist_rows(Tables t, const HistRow *hr, uint64_t n, const uint32_t *jid_of, uint64_t nj) {
    __shared__ unsigned long long s_stage[8][LOG_STAGE];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    LogStage log{s_stage[threadIdx.x >> 5], 0u};
    uint32_t j = JID_NONE, off = 0u, cnt = 0u;
    if (i < n) {
        const HistRow r = hr[i];
        if (r.jrow < nj) j = jid_of[r.jrow];
        off = r.offset;
        cnt = r.cnt;
    }
    const bool ok = j != JID_NONE && cnt != 0u;
    const uint32_t place = ok ? place_offset(t, j, off) : DENSE_NONE;
    if (place != DENSE_NONE) atomicAdd(&t.dense[(uint64_t)j * (DENSE_ROWS * DENSE_W) + place], cnt);
    log_append(t, log, ok && place == DENSE_NONE, olog_entry(j, off, cnt));
    log_flush(t, log);
}

void launch_import_rows(const Tables &t, const JunctionRow *jr, uint64_t nj, const HistRow *hr, uint64_t nh, uint32_t *jid_of,
                        cudaStream_t st) {
    if (nj) k_import_junction_rows<<<blocks_for(nj, 256) < N_ID_POOLS ? blocks_for(nj, 256) : N_ID_POOLS, 256, 0, st>>>(t, jr, nj, jid_of);
    if (nh) k_import_hist_rows<<<blocks_for(nh, 256), 256, 0, st>>>(t, hr, nh, jid_of, nj);
}

/* ---- the exchange over peer memory ---- */
__device__ __forceinline__ char *mail_slot(const Mailboxes &m, int owner, int source) {
    return m.base[owner] + (uint64_t)(m.parity * m.world + source) * m.slot_bytes;
}
/* where the rows for `owner` are put together before they cross the link in one coalesced copy (k_mail_send) */
__device__ __forceinline__ char *outbox_slot(const Mailboxes &m, int owner) { return m.outbox + (uint64_t)owner * m.slot_bytes; }
__device__ __forceinline__ JunctionRow *mail_junction_rows(char *slot) { return reinterpret_cast<JunctionRow *>(slot + MAIL_HEADER_BYTES); }
__device__ __forceinline__ HistRow *mail_hist_rows(const Mailboxes &m, char *slot) {
    return reinterpret_cast<HistRow *>(slot + MAIL_HEADER_BYTES + m.hist_offset);
}

static const uint32_t NOT_EXPORTED = 0xFFFFFFFFu;

/* A warp takes MAIL_BATCH consecutive junction rows at a time.  First it looks at all of them (owner, offset rows its
 * dense counters turn into), lane q keeping what it learnt about row q; then the lanes reserve room behind the owners'
 * cursors together -- junction ids follow the stream, so the rows of a batch mostly share their owner: one pair of
 * atomics per (batch, owner) -- and the rows go into the owner's slot of the local outbox (rows stored one by one
 * straight into the peer's memory made the export three times slower on NVLink than it is locally: 8- and 72-byte
 * stores; k_mail_send moves each outbox slot across the link as one coalesced copy).  Nothing is synchronised beyond
 * the warp.  An exported row is disowned here: its record and offset rows go back to the reset values, so that finalize
 * passes over it and later events would start it afresh. */
static const int MAIL_BATCH = 8;
__global__ void __launch_bounds__(256) k_mail_rows(Tables t, const uint32_t *name_rank, Splitters sp, Mailboxes m,
                                                   unsigned long long *counts, uint32_t *jrow_of) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_keys = junctions_in_use(t);
    const uint64_t n_batches = (n_keys + MAIL_BATCH - 1) / MAIL_BATCH;
    for (uint64_t bt = warp0; bt < n_batches; bt += n_warps) {
        const uint64_t j0 = bt * MAIL_BATCH;
        RowView mine;
        mine.dest = -1; mine.nz = 0u; mine.i0 = mine.i1 = 0u; mine.tid = 0;
        mine.base[0] = mine.base[1] = mine.base[2] = BASE_UNSET;
#pragma unroll 1
        for (int q = 0; q < MAIL_BATCH; ++q) {
            if (j0 + q < n_keys) {
                const RowView r = read_row(t, name_rank, sp, j0 + q, lane);
                if (lane == q) mine = r;
            }
        }
        const bool foreign = mine.dest >= 0 && mine.dest != m.me; /* (lanes >= MAIL_BATCH: false) */
        /* room behind the owners' cursors: the lanes of one owner share a pair of atomics */
        const int gkey = foreign ? mine.dest : -1 - lane;
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, gkey);
        const int leader = __ffs(grp) - 1;
        uint32_t hpre = 0u, htot = 0u;
#pragma unroll
        for (int q = 0; q < MAIL_BATCH; ++q) {
            const int dq = __shfl_sync(0xFFFFFFFFu, gkey, q);
            const uint32_t nq = __shfl_sync(0xFFFFFFFFu, mine.nz, q);
            if (dq == gkey) {
                htot += nq;
                if (q < lane) hpre += nq;
            }
        }
        unsigned long long jpos = 0ull, hpos = 0ull;
        if (foreign && lane == leader) {
            jpos = atomicAdd(&counts[mine.dest], (unsigned long long)__popc(grp));
            if (htot) hpos = atomicAdd(&counts[sp.n_parts + mine.dest], (unsigned long long)htot);
        }
        jpos = __shfl_sync(0xFFFFFFFFu, jpos, leader) + (unsigned long long)__popc(grp & ((1u << lane) - 1u));
        hpos = __shfl_sync(0xFFFFFFFFu, hpos, leader) + hpre;
#pragma unroll 1
        for (int q = 0; q < MAIL_BATCH; ++q) {
            const uint64_t j = j0 + q;
            if (j >= n_keys) break;
            RowView r;
            r.dest = __shfl_sync(0xFFFFFFFFu, mine.dest, q);
            r.nz = __shfl_sync(0xFFFFFFFFu, mine.nz, q);
            const bool fq = __shfl_sync(0xFFFFFFFFu, (int)foreign, q) != 0;
            if (!fq) {
                if (lane == 0) jrow_of[j] = NOT_EXPORTED;
                continue;
            }
            r.i0 = __shfl_sync(0xFFFFFFFFu, mine.i0, q);
            r.i1 = __shfl_sync(0xFFFFFFFFu, mine.i1, q);
            r.tid = __shfl_sync(0xFFFFFFFFu, mine.tid, q);
            r.base[0] = __shfl_sync(0xFFFFFFFFu, mine.base[0], q);
            r.base[1] = __shfl_sync(0xFFFFFFFFu, mine.base[1], q);
            r.base[2] = __shfl_sync(0xFFFFFFFFu, mine.base[2], q);
            const unsigned long long jp = __shfl_sync(0xFFFFFFFFu, jpos, q), hp = __shfl_sync(0xFFFFFFFFu, hpos, q);
            /* a slot that is too small keeps counting: the owner sees counts beyond the capacity and reports it */
            const bool fits = jp < m.cap_j && hp + r.nz <= m.cap_h;
            char *box = outbox_slot(m, r.dest);
            if (lane == 0) {
                if (fits) mail_junction_rows(box)[jp] = make_junction_row(t.stat[j], r);
                jrow_of[j] = fits ? (((uint32_t)r.dest << 24) | (uint32_t)jp) : (((uint32_t)r.dest << 24) | 0xFFFFFFu);
                /* disown */
                uint4 *p = reinterpret_cast<uint4 *>(&t.stat[j]);
                p[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                p[1] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                p[2] = make_uint4(0u, 0u, 0u, 0u);
                p[3] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
                t.olog_cnt[j] = 0u;
            }
            /* its offset rows go out and its counters back to zero (the rows stay given out: their bases are as good
             * for the reads that may follow as any) */
            if (t.dense) write_dense_rows<true>(t, j, mail_hist_rows(m, box), hp, (uint32_t)jp, r, lane, fits);
        }
    }
}

/* the log entries of the exported junctions follow them; they are struck from the local log */
__global__ void __launch_bounds__(256) k_mail_olog(Tables t, Splitters sp, Mailboxes m, unsigned long long *counts, const uint32_t *jrow_of) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_olog = log_entries_in_use(t), n_round = (n_olog + 31ull) & ~31ull;
    const uint64_t n_keys = junctions_in_use(t);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += n_threads) {
        int dest = -1 - lane; /* idle lanes group with nobody */
        unsigned long long e = EMPTY64;
        uint32_t jrow = 0u;
        if (i < n_olog) {
            e = t.olog[i];
            const uint32_t j = (uint32_t)(e >> 32);
            if (j < n_keys) {
                const uint32_t v = jrow_of[j];
                if (v != NOT_EXPORTED) {
                    dest = (int)(v >> 24);
                    jrow = v & 0xFFFFFFu;
                    t.olog[i] = EMPTY64; /* names no junction any more */
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
        if (dest >= 0 && at < m.cap_h && jrow != 0xFFFFFFu) {
            HistRow h;
            h.jrow = jrow; h.offset = (uint16_t)((e >> 16) & 0xFFFFull); h.cnt = (uint16_t)(e & 0xFFFFull);
            mail_hist_rows(m, outbox_slot(m, dest))[at] = h;
        }
    }
}

/* The outbox slots cross the link: blockIdx.y = destination, the blocks of a destination copy its junction rows and its
 * offset rows (as many as the cursors say) with 16-byte loads and stores, consecutive lanes on consecutive addresses. */
__global__ void __launch_bounds__(256) k_mail_send(Mailboxes m, const unsigned long long *counts) {
    const int d = blockIdx.y;
    if (d == m.me) return;
    const unsigned long long nj = counts[d] < m.cap_j ? counts[d] : m.cap_j, nh = counts[m.world + d] < m.cap_h ? counts[m.world + d] : m.cap_h;
    const char *src = outbox_slot(m, d);
    char *dst = mail_slot(m, d, m.me);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t qj = (nj * sizeof(JunctionRow) + 15) / 16, qh = (nh * sizeof(HistRow) + 15) / 16; /* (slots are padded to 256 bytes) */
    const uint4 *sj = reinterpret_cast<const uint4 *>(src + MAIL_HEADER_BYTES);
    uint4 *dj = reinterpret_cast<uint4 *>(dst + MAIL_HEADER_BYTES);
    for (uint64_t i = first; i < qj; i += stride) dj[i] = sj[i];
    const uint4 *sh = reinterpret_cast<const uint4 *>(src + MAIL_HEADER_BYTES + m.hist_offset);
    uint4 *dh = reinterpret_cast<uint4 *>(dst + MAIL_HEADER_BYTES + m.hist_offset);
    for (uint64_t i = first; i < qh; i += stride) dh[i] = sh[i];
}

/* header of a slot: [0] junction rows, [1] offset rows, [2] epoch of the exchange they belong to */
__global__ void k_mail_publish(Mailboxes m, const unsigned long long *counts) {
    const int d = threadIdx.x;
    if (d >= m.world || d == m.me) return;
    volatile unsigned long long *hdr = reinterpret_cast<volatile unsigned long long *>(mail_slot(m, d, m.me));
    hdr[0] = counts[d];
    hdr[1] = counts[m.world + d];
    __threadfence_system(); /* rows (earlier kernels of this stream) and counts before the epoch */
    hdr[2] = m.epoch;
}

void launch_mail_export(const Tables &t, const uint32_t *name_rank, const Splitters &sp, const Mailboxes &m, unsigned long long *counts,
                        uint32_t *jrow_of, int sm_count, cudaStream_t st) {
    k_mail_rows<<<sm_count * 8, 256, 0, st>>>(t, name_rank, sp, m, counts, jrow_of);
    k_mail_olog<<<sm_count * 8, 256, 0, st>>>(t, sp, m, counts, jrow_of);
    if (m.world > 1) {
        unsigned per_dest = (unsigned)((sm_count * 4 + m.world - 2) / (m.world - 1));
        k_mail_send<<<dim3(per_dest ? per_dest : 1, m.world), 256, 0, st>>>(m, counts);
    }
    k_mail_publish<<<1, MAX_PARTS, 0, st>>>(m, counts);
}

/* One warp: lane s waits until source s has published this exchange's epoch into its slot here (every rank enqueues
 * its export before its import, so nobody waits for a kernel that waits for it), then decides whether everything that
 * was sent fits the slots and this rank's tables.  verdict[0] = 1 go / 2 refused; verdict[1..2] = rows arriving. */
__global__ void k_mail_wait(Tables t, Mailboxes m, unsigned long long *verdict, unsigned long long timeout_ns) {
    const int s = threadIdx.x;
    unsigned long long nj = 0ull, nh = 0ull;
    bool bad = false;
    if (s < m.world && s != m.me) {
        volatile unsigned long long *hdr = reinterpret_cast<volatile unsigned long long *>(mail_slot(m, m.me, s));
        unsigned long long t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (hdr[2] != m.epoch) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns) {
                bad = true;
                atomicOr(t.overflow, OVF_PEER);
                break;
            }
            __nanosleep(200);
        }
        __threadfence_system();
        nj = hdr[0];
        nh = hdr[1];
        if (nj > m.cap_j || nh > m.cap_h) {
            bad = true;
            atomicOr(t.overflow, OVF_MAIL);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        nj += __shfl_xor_sync(0xFFFFFFFFu, nj, d);
        nh += __shfl_xor_sync(0xFFFFFFFFu, nh, d);
    }
    const bool any_bad = __any_sync(0xFFFFFFFFu, bad);
    if (threadIdx.x == 0) {
        bool go = !any_bad;
        /* all or nothing: a refused import can be repeated after the host has grown the tables */
        if (go && (junctions_in_use(t) + nj > t.max_j || 2ull * (junctions_in_use(t) + nj) > t.cap_mask + 1ull)) {
            go = false;
            atomicOr(t.overflow, OVF_TABLE);
        }
        if (go && log_entries_in_use(t) + nh > t.olog_cap) {
            go = false;
            atomicOr(t.overflow, OVF_LOG);
        }
        verdict[1] = nj;
        verdict[2] = nh;
        verdict[0] = go ? 1ull : 2ull;
    }
}

/* fold the junction rows of every source (blockIdx.y) out of this rank's own mailbox */
__global__ void __launch_bounds__(256) k_mail_import_junctions(Tables t, Mailboxes m, const unsigned long long *verdict, uint32_t *jid_of) {
    __shared__ IdPool s_pool;
    const int source = blockIdx.y;
    if (source == m.me || verdict[0] != 1ull) return;
    /* (thread blocks of different sources share pools: every block of the grid gets its own) */
    const uint32_t pool_index = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        s_pool.word = t.id_pools[pool_index % N_ID_POOLS];
        s_pool.lock = 0u;
    }
    __syncthreads();
    char *box = mail_slot(m, m.me, source);
    const unsigned long long n_sent = *reinterpret_cast<const unsigned long long *>(box);
    const JunctionRow *jr = mail_junction_rows(box);
    uint32_t *ids = jid_of + (uint64_t)source * m.cap_j;
    const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
    uint32_t spare = JID_NONE;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sent; i += n_threads) ids[i] = import_junction_row(t, &s_pool, jr[i], spare);
    __syncthreads();
    if (threadIdx.x == 0) t.id_pools[pool_index % N_ID_POOLS] = *((volatile unsigned long long *)&s_pool.word);
}

__global__ void __launch_bounds__(256) k_mail_import_hist(Tables t, Mailboxes m, const unsigned long long *verdict, const uint32_t *jid_of) {
    __shared__ unsigned long long s_stage[8][LOG_STAGE];
    const int source = blockIdx.y;
    if (source == m.me || verdict[0] != 1ull) return;
    char *box = mail_slot(m, m.me, source);
    const unsigned long long nj = reinterpret_cast<const unsigned long long *>(box)[0];
    const unsigned long long n_sent = reinterpret_cast<const unsigned long long *>(box)[1];
    const HistRow *hr = mail_hist_rows(m, box);
    const uint32_t *ids = jid_of + (uint64_t)source * m.cap_j;
    const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n_sent + 31ull) & ~31ull;
    LogStage log{s_stage[threadIdx.x >> 5], 0u};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += n_threads) {
        uint32_t j = JID_NONE, off = 0u, cnt = 0u;
        if (i < n_sent) {
            const HistRow r = hr[i];
            if (r.jrow < nj) j = ids[r.jrow];
            off = r.offset;
            cnt = r.cnt;
        }
        const bool ok = j != JID_NONE && cnt != 0u;
        const uint32_t place = ok ? place_offset(t, j, off) : DENSE_NONE;
        if (place != DENSE_NONE) atomicAdd(&t.dense[(uint64_t)j * (DENSE_ROWS * DENSE_W) + place], cnt);
        log_append(t, log, ok && place == DENSE_NONE, olog_entry(j, off, cnt));
    }
    log_flush(t, log);
}

void launch_mail_import(const Tables &t, const Mailboxes &m, unsigned long long *verdict, uint32_t *jid_of, int sm_count,
                        unsigned long long timeout_ns, cudaStream_t st) {
    k_mail_wait<<<1, 32, 0, st>>>(t, m, verdict, timeout_ns);
    if (m.world < 2) return;
    unsigned per_source = (unsigned)((sm_count * 4 + m.world - 2) / (m.world - 1));
    if (per_source * (unsigned)m.world > N_ID_POOLS) per_source = N_ID_POOLS / (unsigned)m.world; /* a pool of junction ids per block */
    k_mail_import_junctions<<<dim3(per_source ? per_source : 1, m.world), 256, 0, st>>>(t, m, verdict, jid_of);
    k_mail_import_hist<<<dim3(per_source ? 2 * per_source : 1, m.world), 256, 0, st>>>(t, m, verdict, jid_of);
}

/* ---- canon filter across ranks ---- */
__global__ void k_export_thresholds(Tables t, ThresholdRow *rows, uint64_t n) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long key = t.jkey[j];
    ThresholdRow r;
    r.tid = t.jtid[j];
    r.i0 = key_start(t.kl, key);
    r.i1 = r.i0 + key_len(t.kl, key);
    r.pad = 0u;
    r.T = t.stat[j].T;
    rows[j] = r;
}

/* a key known here takes the earliest threshold any rank found (its own included) */
__global__ void k_import_thresholds(Tables t, const ThresholdRow *rows, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ThresholdRow r = rows[i];
    if (r.T == EMPTY64) return;
    const uint32_t j = junction_find(t, make_key(t.kl, r.tid, r.i0, r.i1), r.tid).jid;
    if (j != JID_NONE) atomicMin(&t.stat[j].T, r.T);
}

void launch_export_thresholds(const Tables &t, ThresholdRow *rows, uint64_t n, cudaStream_t st) {
    if (n) k_export_thresholds<<<blocks_for(n, 256), 256, 0, st>>>(t, rows, n);
}
void launch_import_thresholds(const Tables &t, const ThresholdRow *rows, uint64_t n, cudaStream_t st) {
    if (n) k_import_thresholds<<<blocks_for(n, 256), 256, 0, st>>>(t, rows, n);
}

/* ---- a sample of the junction keys, for load-balanced splitters ---- */
__global__ void k_sample_keys(Tables t, uint64_t stride, uint64_t max_out, int32_t *tid, uint32_t *i0, unsigned long long *n_out) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nk = junctions_in_use(t);
    const uint64_t n = (nk + stride - 1) / stride;
    if (q == 0) *n_out = n < max_out ? n : max_out;
    if (q >= n || q >= max_out) return;
    const uint64_t j = q * stride;
    const bool hole = t.jslot[j] == SLOT_NONE; /* an id no junction took */
    tid[q] = hole ? -1 : t.jtid[j];
    i0[q] = hole ? 0u : key_start(t.kl, t.jkey[j]);
}

void launch_sample_keys(const Tables &t, uint64_t stride, uint64_t max_out, int32_t *tid, uint32_t *i0, unsigned long long *n_out,
                        cudaStream_t st) {
    k_sample_keys<<<blocks_for(max_out, 256), 256, 0, st>>>(t, stride ? stride : 1, max_out, tid, i0, n_out);
}

}  // namespace ipx
