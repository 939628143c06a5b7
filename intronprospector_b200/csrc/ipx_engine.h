/* ipx_engine.h -- internal interface between the C ABI (ipx_api.cc) and the CUDA side (ipx_kernels.cu). */
#ifndef IPX_ENGINE_H
#define IPX_ENGINE_H
#include <stdint.h>
#include <cuda_runtime.h>
#include "ipx_table.cuh"

namespace ipx {

/* device pointers of one batch (see ipx_batch in include/ipx.h) */
struct BatchView {
    const int32_t *tid;
    const int32_t *pos;
    const uint16_t *flag;
    const uint8_t *nh;
    const uint8_t *xs;
    const uint32_t *cigar_off;
    const uint32_t *cigar;
    uint64_t n;
    uint64_t seq_base;
};

/* HBM-resident genome: raw FASTA characters per sequence */
struct GenomeView {
    const char *const *seq;     /* [n_seqs] device pointers */
    const uint64_t *len;        /* [n_seqs] */
    const int32_t *target_seq;  /* [n_targets] genome sequence of each target or -1 */
    int32_t n_targets;
    int32_t have_genome;
    int32_t skip_missing;
    int32_t filter_canon;       /* 1 when the canon filter is active */
};

/* finished table, device side, key order */
struct DeviceTable {
    uint64_t n;        /* junctions */
    uint64_t n_hist;   /* histogram runs */
    int32_t *chrom;
    uint32_t *i0, *i1, *a0, *a1, *ijunc;
    char *strand;
    char *sites;       /* [n][8] */
    unsigned long long *counts; /* [n][3] */
    unsigned long long *T;      /* [n] creation sequence value (for cross-flush numbering / export) */
    unsigned long long *hist_off; /* [n+1] */
    uint32_t *hist_cnt;           /* [n_hist] */
};

/* packed rows exchanged between ranks (ipx_partial_export / ipx_partial_import) */
struct JunctionRow {
    unsigned long long first[3];
    unsigned long long cnt[3];
    int32_t tid;
    uint32_t i0, i1, a0, a1;
    uint32_t pad;
};
/* one run of a junction's read-offset histogram.  `jrow` is the index of the junction's JunctionRow among the rows the
 * same source sends to the same destination, so the 8-byte row needs no key (there are ~50 of these per junction) */
struct HistRow {
    uint32_t jrow;
    uint16_t offset;
    uint16_t cnt; /* larger counts are sent as several rows */
};

/* a junction key and its creation threshold, exchanged between ranks (ipx_thresholds_export / _import) */
struct ThresholdRow {
    int32_t tid;
    uint32_t i0, i1;
    uint32_t pad;
    unsigned long long T; /* EMPTY64 = no read of this rank may create the junction */
};

enum ExtractMode { MODE_SINGLE = 0, MODE_PASS1 = 1, MODE_PASS2 = 2 };

/* K1 works on tiles of EX_TILE consecutive records.  The events of tile i go to a private region of the event
 * buffer starting at event index ceil(cigar_off[i * EX_TILE] / 2): a read needs at least one N and one following
 * aligned operation per emitted junction, so a tile cannot emit more events than half its CIGAR words and the
 * regions of consecutive tiles never overlap.  tile_info[i] says how many events tile i wrote.  No global
 * atomic is involved; the buffer holds n_cigar / 2 + 1 events. */
/* (measured on B200, 100 M short reads, K1 / K2: 512 records 0.78 / 0.65 ms, 1024: 0.67 / 0.49, 2048: 0.64 / 0.45, 4096: 0.67 /
 * 0.60 -- the hand-over between the warps of a tile costs per tile, and K2 needs enough tiles to balance its warps) */
#ifndef IPX_EX_TILE
#define IPX_EX_TILE 2048
#endif
static const int EX_TILE = IPX_EX_TILE;
struct __align__(16) TileInfo {
    uint32_t count;      /* events in the tile's region: written by K1, events K2 could not fold (table full) after K2 */
    uint32_t first_word; /* cigar_off of its first record: the region starts at event (first_word + 1) / 2 */
    uint32_t pad[2];
};
/* first event index of a tile's region in the event buffer */
__host__ __device__ inline uint64_t tile_event_base(uint32_t first_cigar_word) { return ((uint64_t)first_cigar_word + 1ull) >> 1; }
inline uint64_t tiles_of(uint64_t n_records) { return (n_records + EX_TILE - 1) / EX_TILE; }
/* The long-read kernels cut the batch into tiles of LONG_TILE records instead: a tile is also K2's unit of work, and a
 * million long reads in tiles of 2048 are 488 tiles for 4736 warps (K2 took 2.1 ms instead of 0.7). */
static const int LONG_TILE = 256;
inline uint64_t max_tiles_of(uint64_t n_records) { return (n_records + LONG_TILE - 1) / LONG_TILE; }

/* kernels' launchers; all asynchronous on `st` */
void launch_init_tables(const Tables &t, cudaStream_t st);  /* whole tables, once after allocation */
void launch_init_rows(const Tables &t, uint64_t first, cudaStream_t st); /* junction rows [first, max_j) after growth */
/* only what is in use (and, when the overflow flag is up, the slots that were claimed while no junction row was left) */
void launch_clear_tables(const Tables &t, cudaStream_t st);
void launch_settle_counters(const Tables &t, cudaStream_t st); /* counters back to what exists, before growth */
/* every junction row into fresh hash slots (t.slots: allocated, any content); old_slots: where the rows' bases are */
void launch_rehash(const Tables &t, const Slot *old_slots, cudaStream_t st);
/* returns a CUDA error (attribute set-up of the pipelined kernel can fail), launches counted in *launches; *n_tiles = the
 * tiles the batch was cut into (tile_info has room for max_tiles_of(b.n) + 1) */
cudaError_t launch_extract(const BatchView &b, uint64_t n_cigar, const ExtractParams &p, const Tables &t, Event *out,
                           TileInfo *tile_info, unsigned long long *next_record, int sm_count, cudaStream_t st, int *launches,
                           uint64_t *n_tiles);
/* K2 consumes the events: afterwards tile_info[].count is what it could not fold (0 unless a table was full) */
void launch_fold(int mode, const Tables &t, Event *ev, TileInfo *tile_info, uint64_t n_tiles, int sm_count, cudaStream_t st);
void launch_resolve_T(const Tables &t, const GenomeView &g, cudaStream_t st);
/* ipx_batch_compact -> ipx_batch columns on the device: cigar_off = exclusive scan of the operation counts (n + 1
 * entries), tid = the run each record falls into.  tmp: scratch of *tmp_bytes (queried when tmp == NULL). */
cudaError_t expand_compact(const uint16_t *n_ops, uint64_t n, uint32_t *cigar_off, const int32_t *run_tid,
                           const unsigned long long *run_start, uint64_t n_runs, int32_t *tid, void *tmp, size_t *tmp_bytes,
                           cudaStream_t st);
void launch_export_thresholds(const Tables &t, ThresholdRow *rows, uint64_t n, cudaStream_t st);
void launch_import_thresholds(const Tables &t, const ThresholdRow *rows, uint64_t n, cudaStream_t st);
/* rows of ONE source: the offset rows refer to the junction rows of the same call; jid_of: [nj] scratch */
void launch_import_rows(const Tables &t, const JunctionRow *jr, uint64_t nj, const HistRow *hr, uint64_t nh, uint32_t *jid_of,
                        cudaStream_t st);

/* Workspace of finalize.  Everything it needs lives in buffers sized by grow-only bounds (B junction ids, L read-offset
 * log entries, R histogram runs, C keys per sort bucket), so the whole sequence is one CUDA graph that is captured once
 * per set of bounds and replayed with a single launch.  The kernels read the actual counts from device memory and
 * raise a flag when a bound was too small; the host then grows the bound and replays (the graph only reads the
 * tables, so a replay is harmless). */
struct FinalizeWork {
    uint64_t B = 0, L = 0, R = 0, C = 0;
    uint64_t wish_B = 0, wish_L = 0, wish_R = 0, wish_C = 0; /* what finalize_check asked for */
    uint64_t first_B = 1ull << 16, first_L = 1ull << 16, first_R = 1ull << 20; /* bounds of the very first finalize */
    unsigned long long *sort_key = nullptr, *tkey = nullptr;                     /* [B] */
    uint32_t *sort_sub = nullptr, *order = nullptr, *t_order = nullptr, *t_rank = nullptr; /* [B] */
    uint32_t *seg_off = nullptr, *seg_fill = nullptr, *runs_j = nullptr, *big_list = nullptr; /* [B] */
    uint16_t *nz4 = nullptr;     /* [4 B] non-zero counters per dense row */
    uint32_t *below3 = nullptr;  /* [3 B] log runs below each extra dense row */
    uint32_t *runs_r = nullptr;  /* [B + 1] */
    uint32_t *seg = nullptr;     /* [L] log entries grouped by junction, then the ordered runs' counts */
    uint32_t *big_scratch = nullptr;
    unsigned long long *scan_state = nullptr;
    /* the two sample sorts (key order, sjN) run side by side */
    unsigned long long *ss_splitters[2] = {nullptr, nullptr}, *ss_bk[2] = {nullptr, nullptr};
    uint32_t *ss_bucket_cnt[2] = {nullptr, nullptr}, *ss_bs[2] = {nullptr, nullptr}, *ss_bv[2] = {nullptr, nullptr};
    cudaStream_t s2 = nullptr, s3 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join2 = nullptr, ev_join3 = nullptr;
    /* [8]: 0 junctions, 1 histogram runs, 2 log entries in segments, 3 flags (1: B too small, 2: L, 4: R, 8: C),
     *      4 junction ids in use, 5 log entries in use, 6 junctions ordered by the block pass, 7 largest bucket */
    unsigned long long *d_count = nullptr;
    unsigned long long *h_count = nullptr; /* pinned mirror, filled by the graph's last node */
    DeviceTable out{};                     /* arrays of B rows (hist_cnt: R runs) */
    cudaGraphExec_t exec = nullptr;
    /* what the captured graph depends on */
    struct Key {
        uint64_t B, L, R, C, max_j, olog_cap;
        const void *tab_keys, *tab_rows, *tab_log, *tab_base, *name_rank, *gseq, *target_seq, *dense;
        int32_t n_targets, have_genome, skip_missing, filter_canon, wide, len_bits;
        bool operator==(const Key &o) const {
            return B == o.B && L == o.L && R == o.R && C == o.C && max_j == o.max_j && olog_cap == o.olog_cap &&
                   tab_keys == o.tab_keys && tab_rows == o.tab_rows && tab_log == o.tab_log && tab_base == o.tab_base && name_rank == o.name_rank &&
                   gseq == o.gseq && target_seq == o.target_seq && dense == o.dense && n_targets == o.n_targets &&
                   have_genome == o.have_genome && skip_missing == o.skip_missing && filter_canon == o.filter_canon &&
                   wide == o.wide && len_bits == o.len_bits;
        }
    } key{};
    int launches_per_replay = 0;
};

/* key order, sjN ranking, histogram ordering, row gather: enqueues one graph launch on `st` and the copy of the
 * counters to w.h_count; nothing is synchronised.  finalize_check (after a stream synchronise) says whether a bound
 * was too small -- it then grows the bounds and the caller enqueues again.  `name_rank[tid]` orders targets like
 * std::string::operator< on their names. */
cudaError_t finalize_enqueue(const Tables &t, const GenomeView &g, const uint32_t *d_name_rank, FinalizeWork &w, cudaStream_t st);
bool finalize_check(FinalizeWork &w); /* true = results valid (w.out.n / n_hist set); false = bounds grown, run again */
void finalize_release(FinalizeWork &w);

/* Sample sort of (u64 key, u32 sub-key) pairs -> the values 0 .. n-1 in key order (ipx_finalize.cu).  Keys equal to
 * EMPTY64 are left out; pairs must be unique.  `n_ptr`: the number of keys, on the device (at most `bound`). */
struct SampleSort {
    unsigned long long *splitters; /* [sample_sort_max_buckets()] */
    uint32_t *bucket_cnt;          /* [sample_sort_max_buckets() + 1]: counts, then the number of buckets */
    unsigned long long *bk;        /* [sample_sort_buckets(bound) * C] */
    uint32_t *bs, *bv;             /* [sample_sort_buckets(bound) * C] */
    uint64_t C;                    /* capacity of a bucket in bk / bs / bv */
};
uint64_t sample_sort_buckets(uint64_t bound);
uint64_t sample_sort_max_buckets();
/* d_count[3] |= 8 and d_count[7] = size when a bucket held more than C keys (grow C and run again) */
cudaError_t sample_sort_enqueue(const unsigned long long *keys, const uint32_t *sub, const unsigned long long *n_ptr, uint64_t bound,
                                const SampleSort &ss, uint32_t *out_vals, uint32_t *out_rank, unsigned long long *d_count, cudaStream_t st);

/* stream-ordered allocation from the device's default pool (kept resident, see ipx_create): no implicit
 * device synchronisation and no driver round trip once the pool is warm */
template <class T>
inline cudaError_t dmalloc(cudaStream_t st, T **p, size_t bytes) {
    return cudaMallocAsync((void **)p, bytes ? bytes : 1, st);
}
template <class T>
inline void dfree(cudaStream_t st, T *p) {
    if (p) cudaFreeAsync((void *)p, st);
}

/* Export of the partial table for the key-range exchange, straight from the tables (no finalize, no sort): the
 * destination of a key is found by comparing (chromosome-name rank, intron_start) with the splitters; pass one counts
 * rows per destination, pass two scatters them behind per-destination cursors. */
static const int MAX_PARTS = 64;
struct Splitters {
    int32_t n_parts;
    uint32_t rank[MAX_PARTS - 1];  /* first key of parts 1 .. n_parts-1: chromosome-name rank, */
    uint32_t start[MAX_PARTS - 1]; /* intron_start */
    uint32_t jbase[MAX_PARTS];     /* scatter pass: index of each part's first junction row in the send buffer */
};
/* counts / cursors: [2 * n_parts] unsigned long long on the device: junction rows per part, then offset rows per part;
 * jrow_of: [max_j] owner << 24 | index of the junction's row among the rows sent to that owner */
void launch_export_count(const Tables &t, const uint32_t *name_rank, const Splitters &sp, uint64_t n_keys, uint64_t n_olog,
                         unsigned long long *counts, cudaStream_t st);
void launch_export_scatter(const Tables &t, const uint32_t *name_rank, const Splitters &sp, uint64_t n_keys, uint64_t n_olog,
                           unsigned long long *cursors, uint32_t *jrow_of, JunctionRow *jr, HistRow *hr, cudaStream_t st);

/* Exchange over peer memory.  Every rank owns a mailbox of 2 (parity) x world slots, all ranks with the same layout;
 * slot (parity, source) = [header: junction rows, offset rows, epoch (3 x u64)] [cap_j junction rows] [cap_h offset rows]. */
static const uint64_t MAIL_HEADER_BYTES = 128;
struct Mailboxes {
    char *base[MAX_PARTS]; /* mailbox of every rank as mapped into this process (own included) */
    char *outbox;          /* local staging: one slot per destination, laid out like the slot it is copied into */
    uint64_t slot_bytes, hist_offset, cap_j, cap_h;
    unsigned long long epoch; /* number of the exchange in flight (same on every rank) */
    int32_t world, me, parity;
};
/* export the junctions other ranks own into their slots, disown them here, publish counts + epoch */
void launch_mail_export(const Tables &t, const uint32_t *name_rank, const Splitters &sp, const Mailboxes &m, unsigned long long *counts,
                        uint32_t *jrow_of, int sm_count, cudaStream_t st);
/* wait (on the device) for every source's epoch, then fold all slots; verdict: [3] words on the device, jid_of: [world * cap_j] */
void launch_mail_import(const Tables &t, const Mailboxes &m, unsigned long long *verdict, uint32_t *jid_of, int sm_count,
                        unsigned long long timeout_ns, cudaStream_t st);
/* every stride-th junction key (target index, intron start), at most max_out; *n_out = how many */
void launch_sample_keys(const Tables &t, uint64_t stride, uint64_t max_out, int32_t *tid, uint32_t *i0, unsigned long long *n_out,
                        cudaStream_t st);

}  // namespace ipx
#endif
