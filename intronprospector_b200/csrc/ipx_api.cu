/* ipx_api.cu -- implementation of the C ABI in include/ipx.h on top of the CUDA engine.
 * CUDA-only: every compute entry point fails with IPX_ERR_CUDA when no device is usable.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/ipx.h"
#include "ipx_engine.h"
#include "ipx_host.h"
#include "ipx_merge.h"

using namespace ipx;

namespace {

/* Pinned host blocks that back the tables ipx_flush returns: the columns arrive from the device straight into the
 * block the table then owns (no staging copy), and ipx_table_free hands the block back for the next flush. */
struct PinnedPool {
    std::mutex mu;
    std::vector<std::pair<void *, size_t>> free_blocks;
    void *take(size_t bytes, size_t *cap) {
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t i = 0; i < free_blocks.size(); i++)
                if (free_blocks[i].second >= bytes) {
                    void *p = free_blocks[i].first;
                    *cap = free_blocks[i].second;
                    free_blocks.erase(free_blocks.begin() + i);
                    return p;
                }
        }
        void *p = nullptr;
        const size_t want = bytes + bytes / 4 + (1 << 20);
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        *cap = want;
        return p;
    }
    void give(void *p, size_t cap) {
        std::lock_guard<std::mutex> g(mu);
        if (free_blocks.size() < 4) {
            free_blocks.emplace_back(p, cap);
            return;
        }
        /* keep the larger blocks */
        size_t smallest = 0;
        for (size_t i = 1; i < free_blocks.size(); i++)
            if (free_blocks[i].second < free_blocks[smallest].second) smallest = i;
        if (free_blocks[smallest].second < cap) {
            std::swap(free_blocks[smallest].first, p);
            std::swap(free_blocks[smallest].second, cap);
        }
        cudaFreeHost(p);
    }
};
PinnedPool g_pinned_pool;

struct HostTable {
    ipx_table t;
    void *block = nullptr; /* pinned block holding the columns (tables from ipx_flush); else the vectors do */
    size_t block_cap = 0;
    ~HostTable() {
        if (block) g_pinned_pool.give(block, block_cap);
    }
    std::vector<int32_t> chrom, missing;
    std::vector<uint32_t> i0, i1, a0, a1, ijunc, hist_cnt;
    std::vector<char> strand, sites;
    std::vector<uint64_t> counts, hist_off;
    std::vector<float> conf;
    std::vector<uint64_t> creation;
    std::vector<std::string> site_text; /* merge: the input's splice_sites strings at full length */
    std::vector<const char *> site_ptrs;
    std::vector<uint32_t> site_ids;
    void bind() {
        if (block) { /* the columns live in the block: only what stays in vectors is rebound */
            t.missing_targets = missing.data();
            t.n_missing_targets = missing.size();
            if (!ijunc.empty()) t.ijunc = ijunc.data(); /* renumbered by ipx_table_set_ijunc */
            return;
        }
        t.n_junctions = chrom.size();
        t.chrom = chrom.data();
        t.intron_start = i0.data();
        t.intron_end = i1.data();
        t.anchor_start = a0.data();
        t.anchor_end = a1.data();
        t.strand = strand.data();
        t.splice_sites = sites.data();
        t.read_counts = counts.data();
        t.ijunc = ijunc.data();
        t.confidence = conf.data();
        t.hist_off = hist_off.data();
        t.hist_cnt = hist_cnt.data();
        t.n_hist = hist_cnt.size();
        t.missing_targets = missing.data();
        t.n_missing_targets = missing.size();
        t.creation_seq = creation.empty() ? nullptr : creation.data();
        if (!site_text.empty()) {
            site_ptrs.resize(site_text.size());
            for (size_t i = 0; i < site_text.size(); i++) site_ptrs[i] = site_text[i].c_str();
            t.site_strings = site_ptrs.data();
            t.site_ids = site_ids.data();
        }
    }
};

struct DeviceBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr;
    cudaError_t reserve(size_t bytes, cudaStream_t stream) {
        st = stream;
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocAsync(&p, want, st);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
        cap = 0;
    }
};

struct StageTimer {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
    double ms = 0.0;
    uint64_t launches = 0;
};

/* what ipx_merge_rows_fold keeps between calls: the intron table, its counters, the upload lanes */
struct MergeWork {
    Tables tab{};              /* slots, jkey / jtid / jslot, counters; the wide key layout */
    MStat *stat = nullptr;     /* [max_j] */
    uint8_t *slow_of_id = nullptr; /* [max_j] introns the host folds */
    unsigned long long *counters = nullptr; /* n_keys, overflow, n_unfolded, n_slow, d_count[8] (16 words apart), id pools from 128 */
    unsigned long long *h_state = nullptr;  /* pinned, 16 words */
    struct Lane {
        cudaStream_t st;
        void *stage[2];
        cudaEvent_t free_ev[2];
    };
    std::vector<Lane> lanes;
};

static const size_t COUNTER_WORDS = 16 * 16 + 1024; /* 16 cache lines of counters and flags, then the pools of junction ids */

}  // namespace

struct ipx_handle {
    ipx_params params;
    ExtractParams xp;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;

    /* tables */
    Tables tab{};
    bool tables_ready = false;
    unsigned long long *d_counters = nullptr; /* COUNTER_WORDS: one 128-byte line per counter */
    unsigned long long *h_flags = nullptr;    /* pinned: counters / flags read back from the device */
    bool dirty = false;                       /* events folded since the last clear */

    /* targets / genome */
    std::vector<std::string> targets;
    uint32_t *d_name_rank = nullptr;
    size_t d_name_rank_cap = 0;
    bool have_genome = false;
    std::vector<std::string> gnames;
    std::vector<uint64_t> glen;
    std::vector<char *> d_gseq;
    const char **d_gseq_ptrs = nullptr;
    uint64_t *d_glen = nullptr;
    int32_t *d_target_seq = nullptr;
    size_t d_target_seq_cap = 0;
    std::vector<int32_t> target_seq;
    std::unordered_set<std::string> warned_missing;

    /* staging for host batches */
    DeviceBuf s_tid, s_pos, s_flag, s_nh, s_xs, s_off, s_cig;
    DeviceBuf s_nops, s_runs, s_tmp; /* compact batches: operation counts, target runs, scan scratch */
    DeviceBuf ev_buf;   /* junction events of the batch in flight (K1 -> K2), one region per tile */
    DeviceBuf tile_cnt; /* TileInfo per tile */
    /* The last batch folded with the events still in ev_buf / tile_cnt: when a table turns out to have been full
     * (overflow flag, read back lazily) the host grows it and repeats the fold, which picks up exactly the events
     * that were left (see k_fold). */
    bool unsettled = false;
    uint64_t unsettled_tiles = 0;
    bool import_pending = false;         /* ipx_partial_import_mail enqueued, its verdict not read yet */
    bool exported_unsettled = false;     /* an export went out before the last batch's fold was known to be complete */
    bool late_fold = false;              /* ... and it was not: events were folded after the export (ipx_exchange_settle) */
    uint64_t log_bound = 0; /* upper bound of the log entries written since the last clear (sizes the log ahead of a batch) */
    /* host -> device copies of the last pushes: an event behind each, so that a caller that packs into two pinned
     * buffers in turn knows when one may be written again (ipx_sync_uploads) */
    static const int UPLOAD_RING = 4;
    cudaEvent_t upload_done[UPLOAD_RING] = {};
    uint64_t n_uploads = 0;
    /* deferred counting (multi-GPU with the canon filter): the events of every batch pushed since the last count */
    struct Deferred {
        void *events;
        void *tile_info;
        uint64_t n_tiles;
    };
    std::vector<Deferred> deferred;
    /* key-range export in flight (ipx_partial_sizes -> ipx_partial_export) */
    Splitters export_split{};
    /* exchange over peer memory */
    Mailboxes mail = {};
    char *mail_own = nullptr;
    bool mail_opened[MAX_PARTS] = {}; /* peers mapped with cudaIpcOpenMemHandle */
    bool mail_connected = false;
    unsigned long long *d_export = nullptr; /* [2 * MAX_PARTS] counts, then cursors */
    unsigned long long export_counts[2 * MAX_PARTS + 4] = {}; /* + keys, log entries, overflow flag */
    unsigned long long export_cursors[2 * MAX_PARTS] = {};
    uint64_t export_keys = 1 << 16, export_olog = 1 << 16; /* launch bounds of the export passes (grow-only) */
    DeviceBuf jrow_of;  /* [max_j] export: owner << 24 | junction-row index of each junction */
    DeviceBuf jid_of;   /* import: junction id of each received junction row */
    unsigned long long mail_timeout_ns = 20ull * 1000 * 1000 * 1000; /* a peer that never publishes: give up */
    StageTimer tm_export, tm_import;
    bool export_ready = false;
    uint64_t seq_hi = 0; /* one past the largest record stream index pushed so far */

    /* finalize */
    FinalizeWork fw;
    bool finalized = false;        /* a finalize has been enqueued since the last change of the tables */
    bool finalize_settled = false; /* ... and its bounds were checked (fw.out.n / n_hist are valid) */
    bool bounds_proven = false;    /* the last checked finalize fitted its buffers: the next one need not be waited for */
    uint32_t ijunc_base = 0;

    /* merge */
    MergeWork mw;

    /* timings */
    StageTimer tm_extract, tm_aggregate, tm_motif, tm_finalize, tm_merge;
    ipx_timings totals{};
    unsigned long long events_base = 0, live_base = 0;
};

namespace {

int fail(ipx_handle *h, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CU(h, call)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail((h), IPX_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), \
                        __FILE__, __LINE__, cudaGetErrorString(e__));                            \
    } while (0)

struct Span {
    ipx_handle *h;
    StageTimer *t;
    cudaEvent_t a = nullptr, b = nullptr;
    Span(ipx_handle *h_, StageTimer *t_, uint64_t launches) : h(h_), t(t_) {
        t->launches += launches;
        if (h->params.profile) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, h->stream);
        }
    }
    ~Span() {
        if (a) {
            cudaEventRecord(b, h->stream);
            t->spans.emplace_back(a, b);
        }
    }
};

void drain(StageTimer &t) {
    for (auto &s : t.spans) {
        float ms = 0.f;
        cudaEventSynchronize(s.second);
        if (cudaEventElapsedTime(&ms, s.first, s.second) == cudaSuccess) t.ms += ms;
        cudaEventDestroy(s.first);
        cudaEventDestroy(s.second);
    }
    t.spans.clear();
}

void drop_deferred(ipx_handle *h) {
    for (auto &d : h->deferred) {
        cudaFreeAsync(d.events, h->stream);
        cudaFreeAsync(d.tile_info, h->stream);
    }
    h->deferred.clear();
}

int bits_for(uint32_t v) {
    int b = 1;
    while (b < 32 && (v >> b) != 0) b++;
    return b;
}

uint64_t pow2_at_least(uint64_t v) {
    uint64_t p = 1024;
    while (p < v) p <<= 1;
    return p;
}

void set_counter_pointers(ipx_handle *h) {
    Tables &t = h->tab;
    t.n_keys = h->d_counters + 0;
    t.n_olog = h->d_counters + 16;
    t.n_events = h->d_counters + 32;
    t.n_live = h->d_counters + 48;
    t.err_sv = h->d_counters + 64;
    t.overflow = (uint32_t *)(h->d_counters + 80);
    t.n_unfolded = h->d_counters + 128;
    t.next_tile = h->d_counters + 144;
    t.id_pools = h->d_counters + 256;
}

int alloc_counters(ipx_handle *h) {
    if (h->d_counters) return IPX_OK;
    CU(h, cudaMalloc(&h->d_counters, COUNTER_WORDS * sizeof(unsigned long long)));
    CU(h, cudaMemsetAsync(h->d_counters, 0, COUNTER_WORDS * sizeof(unsigned long long), h->stream));
    if (!h->h_flags) CU(h, cudaMallocHost(&h->h_flags, 8 * sizeof(unsigned long long)));
    return IPX_OK;
}

/* the junction rows: everything indexed by junction id */
int alloc_rows(ipx_handle *h, Tables &t, uint64_t max_j, bool dense) {
    t.max_j = max_j;
    CU(h, cudaMalloc(&t.jkey, max_j * sizeof(unsigned long long)));
    CU(h, cudaMalloc(&t.jtid, max_j * sizeof(int32_t)));
    CU(h, cudaMalloc(&t.jslot, max_j * sizeof(uint32_t)));
    CU(h, cudaMalloc(&t.stat, max_j * sizeof(JStat)));
    CU(h, cudaMalloc(&t.olog_cnt, max_j * sizeof(uint32_t)));
    t.dense = nullptr;
    /* dense offset rows: DENSE_ROWS x 1 KB per junction */
    if (dense) CU(h, cudaMalloc(&t.dense, max_j * DENSE_ROWS * DENSE_W * sizeof(uint32_t)));
    return IPX_OK;
}

void free_rows(Tables &t) {
    cudaFree(t.jkey); cudaFree(t.jtid); cudaFree(t.jslot); cudaFree(t.stat); cudaFree(t.olog_cnt); cudaFree(t.dense);
    t.jkey = nullptr; t.jtid = nullptr; t.jslot = nullptr; t.stat = nullptr; t.olog_cnt = nullptr; t.dense = nullptr;
}

int alloc_tables(ipx_handle *h) {
    if (h->tables_ready) return IPX_OK;
    /* table_capacity = hash slots; at most half of them are used, so half as many junction rows.  Both grow on demand. */
    uint64_t want_cap = h->params.table_capacity ? h->params.table_capacity : (1ull << 22);
    uint64_t cap = pow2_at_least(want_cap); /* (a power of two >= 1024: the junction rows come in whole blocks of ids) */
    uint64_t hcap = h->params.hist_capacity ? std::max<uint64_t>(h->params.hist_capacity, 1024) : (1ull << 26);
    if (cap > (1ull << 31)) return fail(h, IPX_ERR_ARG, "table_capacity too large");
    int rc = alloc_counters(h);
    if (rc) return rc;
    Tables &t = h->tab;
    CU(h, cudaMalloc(&t.slots, cap * sizeof(Slot)));
    t.cap_mask = cap - 1;
    t.extra_rows = 1;
    t.dense_lead = DENSE_LEAD;
    rc = alloc_rows(h, t, cap / 2, !getenv("IPX_NO_DENSE"));
    if (rc) return rc;
    CU(h, cudaMalloc(&t.olog, hcap * sizeof(unsigned long long)));
    t.olog_cap = hcap;
    set_counter_pointers(h);
    t.kl = make_key_layout(bits_for(h->params.max_intron_length), h->targets.size());
    launch_init_tables(t, h->stream);
    CU(h, cudaGetLastError());
    h->tables_ready = true;
    return IPX_OK;
}

/* ---- growth: nothing here is on the steady-state path ---- */
/* twice the junction rows (prefix copy of what is in use; the new rows get the reset values) */
int grow_rows(ipx_handle *h, uint64_t in_use) {
    Tables old = h->tab, &t = h->tab;
    const uint64_t max_j = old.max_j * 2;
    if (max_j >> 31) return fail(h, IPX_ERR_CAPACITY, "more than 2^31 junctions");
    int rc = alloc_rows(h, t, max_j, old.dense != nullptr);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    CU(h, cudaMemcpyAsync(t.jkey, old.jkey, in_use * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    CU(h, cudaMemcpyAsync(t.jtid, old.jtid, in_use * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    CU(h, cudaMemcpyAsync(t.jslot, old.jslot, in_use * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    CU(h, cudaMemcpyAsync(t.stat, old.stat, in_use * sizeof(JStat), cudaMemcpyDeviceToDevice, st));
    CU(h, cudaMemcpyAsync(t.olog_cnt, old.olog_cnt, in_use * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    if (t.dense) CU(h, cudaMemcpyAsync(t.dense, old.dense, in_use * DENSE_ROWS * DENSE_W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    launch_init_rows(t, in_use, st);
    CU(h, cudaStreamSynchronize(st));
    free_rows(old);
    return IPX_OK;
}

/* fresh hash slots (twice as many when `twice`), every junction row inserted again */
int rebuild_slots(ipx_handle *h, bool twice) {
    Tables &t = h->tab;
    const uint64_t cap = (t.cap_mask + 1) * (twice ? 2 : 1);
    if (cap > (1ull << 32)) return fail(h, IPX_ERR_CAPACITY, "junction hash table beyond 2^32 slots");
    Slot *old = t.slots;
    t.slots = nullptr;
    CU(h, cudaMalloc(&t.slots, cap * sizeof(Slot)));
    t.cap_mask = cap - 1;
    launch_rehash(t, old, h->stream);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(old);
    return IPX_OK;
}

int grow_log(ipx_handle *h, uint64_t in_use, uint64_t want) {
    Tables &t = h->tab;
    uint64_t cap = t.olog_cap;
    while (cap < want) cap *= 2;
    if (cap >> 32) return fail(h, IPX_ERR_CAPACITY, "read-offset log beyond 2^32 entries");
    unsigned long long *fresh = nullptr;
    CU(h, cudaMalloc(&fresh, cap * sizeof(unsigned long long)));
    if (in_use) CU(h, cudaMemcpyAsync(fresh, t.olog, in_use * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(t.olog);
    t.olog = fresh;
    t.olog_cap = cap;
    return IPX_OK;
}

/* counters and flags of the tables, on the host (synchronises the stream) */
struct TableState {
    unsigned long long n_keys, n_olog, n_unfolded;
    uint32_t overflow;
};
int read_table_state(ipx_handle *h, TableState *out) {
    unsigned long long *f = h->h_flags;
    cudaStream_t st = h->stream;
    CU(h, cudaMemcpyAsync(&f[0], h->tab.n_keys, 8, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(&f[1], h->tab.n_olog, 8, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(&f[2], h->tab.n_unfolded, 8, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(&f[3], h->tab.overflow, 4, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    out->n_keys = f[0]; out->n_olog = f[1]; out->n_unfolded = f[2]; out->overflow = (uint32_t)f[3];
    return IPX_OK;
}

/* make room for `bound` more log entries (an upper bound of what the next launches can write) */
int ensure_log(ipx_handle *h, uint64_t bound) {
    if (h->log_bound + bound <= h->tab.olog_cap) {
        h->log_bound += bound;
        return IPX_OK;
    }
    /* the bound is loose (one entry per two CIGAR words): see what is really there before growing */
    TableState ts;
    int rc = read_table_state(h, &ts);
    if (rc) return rc;
    const uint64_t in_use = std::min<uint64_t>(ts.n_olog, h->tab.olog_cap);
    if (in_use + bound > h->tab.olog_cap) {
        rc = grow_log(h, in_use, in_use + bound);
        if (rc) return rc;
    }
    h->log_bound = in_use + bound;
    return IPX_OK;
}

GenomeView genome_view(const ipx_handle *h) {
    GenomeView g{};
    g.seq = h->d_gseq_ptrs;
    g.len = h->d_glen;
    g.target_seq = h->d_target_seq;
    g.n_targets = (int32_t)h->targets.size();
    g.have_genome = h->have_genome ? 1 : 0;
    g.skip_missing = h->params.skip_missing_targets ? 1 : 0;
    g.filter_canon = (h->have_genome && h->params.junction_filter == IPX_FILTER_CANON) ? 1 : 0;
    return g;
}

bool two_pass(const ipx_handle *h) {
    /* events before the first filter-passing event of a junction are dropped (add_junction,
     * junctions_extractor.cc:494-502); only the canon filter over a genome can reject */
    return h->have_genome && h->params.junction_filter == IPX_FILTER_CANON;
}

int upload_target_maps(ipx_handle *h) {
    const size_t n = h->targets.size();
    if (n == 0) return IPX_OK;
    /* rank of every name under std::string::operator< (JunctionKey operator<, junctions.hh:77-100) */
    std::vector<uint32_t> order(n), rank(n);
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h->targets[a] < h->targets[b]; });
    for (size_t i = 0; i < n; i++) {
        if (i > 0 && h->targets[order[i]] == h->targets[order[i - 1]])
            return fail(h, IPX_ERR_ARG, "duplicate target name '%s'", h->targets[order[i]].c_str());
        rank[order[i]] = (uint32_t)i;
    }
    if (h->d_name_rank_cap < n) {
        if (h->d_name_rank) cudaFree(h->d_name_rank);
        CU(h, cudaMalloc(&h->d_name_rank, n * sizeof(uint32_t)));
        h->d_name_rank_cap = n;
    }
    CU(h, cudaMemcpyAsync(h->d_name_rank, rank.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    /* target -> genome sequence (Genome::has_seq, genome.cc:58-60) */
    h->target_seq.assign(n, -1);
    if (h->have_genome) {
        std::unordered_map<std::string, int32_t> by_name;
        for (size_t g = 0; g < h->gnames.size(); g++) by_name.emplace(h->gnames[g], (int32_t)g);
        for (size_t i = 0; i < n; i++) {
            auto it = by_name.find(h->targets[i]);
            if (it != by_name.end()) h->target_seq[i] = it->second;
        }
    }
    if (h->d_target_seq_cap < n) {
        if (h->d_target_seq) cudaFree(h->d_target_seq);
        CU(h, cudaMalloc(&h->d_target_seq, n * sizeof(int32_t)));
        h->d_target_seq_cap = n;
    }
    CU(h, cudaMemcpyAsync(h->d_target_seq, h->target_seq.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return IPX_OK;
}

void free_genome(ipx_handle *h) {
    for (char *p : h->d_gseq) cudaFree(p);
    h->d_gseq.clear();
    if (h->d_gseq_ptrs) cudaFree((void *)h->d_gseq_ptrs);
    if (h->d_glen) cudaFree(h->d_glen);
    h->d_gseq_ptrs = nullptr;
    h->d_glen = nullptr;
    h->have_genome = false;
}

/* K2 (and K3 with the canon filter) over events that are on the device */
void fold_events(ipx_handle *h, Event *events, TileInfo *tile_info, uint64_t n_tiles) {
    if (two_pass(h)) {
        {
            Span s(h, &h->tm_aggregate, 1);
            launch_fold(MODE_PASS1, h->tab, events, tile_info, n_tiles, h->sm_count, h->stream);
        }
        {
            Span s(h, &h->tm_motif, 1);
            launch_resolve_T(h->tab, genome_view(h), h->stream);
        }
        {
            Span s(h, &h->tm_aggregate, 1);
            launch_fold(MODE_PASS2, h->tab, events, tile_info, n_tiles, h->sm_count, h->stream);
        }
    } else {
        Span s(h, &h->tm_aggregate, 1);
        launch_fold(MODE_SINGLE, h->tab, events, tile_info, n_tiles, h->sm_count, h->stream);
    }
}

/* Give the tables room and repeat `again` (the launches that met the full table) until nothing is left over.
 * Rows are doubled when the junction ids ran out; the hash slots are rebuilt either way (slots claimed while no row was
 * left are dropped by that) and doubled when more than half of them would be in use. */
template <class Again>
int grow_and_repeat(ipx_handle *h, TableState ts, Again again) {
    for (int round = 0; round < 40 && (ts.overflow & (OVF_TABLE | OVF_LOG)); round++) {
        launch_settle_counters(h->tab, h->stream);
        const uint64_t in_use = std::min<uint64_t>(ts.n_keys, h->tab.max_j);
        int rc = IPX_OK;
        if (ts.overflow & OVF_LOG) rc = grow_log(h, std::min<uint64_t>(ts.n_olog, h->tab.olog_cap), h->tab.olog_cap * 2);
        if (rc) return rc;
        if (ts.overflow & OVF_TABLE) {
            if (ts.n_keys >= h->tab.max_j) rc = grow_rows(h, in_use);
            if (rc) return rc;
            rc = rebuild_slots(h, 2 * h->tab.max_j > h->tab.cap_mask + 1 || 4 * in_use > h->tab.cap_mask + 1);
            if (rc) return rc;
        }
        rc = again();
        if (rc) return rc;
        rc = read_table_state(h, &ts);
        if (rc) return rc;
    }
    if (ts.overflow & (OVF_TABLE | OVF_LOG)) return fail(h, IPX_ERR_CAPACITY, "the junction table could not be grown far enough");
    return IPX_OK;
}

/* Was a table full while the last batch was folded?  Then grow it and fold what was left.  Called before the batch's
 * event buffer is used again and before results are read; costs one host round trip. */
int settle_batch(ipx_handle *h, bool *repeated = nullptr) {
    if (repeated) *repeated = false;
    if (!h->unsettled) return IPX_OK;
    TableState ts;
    int rc = read_table_state(h, &ts);
    if (rc) return rc;
    h->unsettled = false;
    if (!(ts.overflow & (OVF_TABLE | OVF_LOG))) {
        h->exported_unsettled = false;
        return IPX_OK;
    }
    if (repeated) *repeated = true;
    if (h->exported_unsettled) h->late_fold = true; /* what is folded now may belong to another rank */
    Event *events = (Event *)h->ev_buf.p;
    TileInfo *tile_info = (TileInfo *)h->tile_cnt.p;
    const uint64_t n_tiles = h->unsettled_tiles;
    return grow_and_repeat(h, ts, [&]() {
        fold_events(h, events, tile_info, n_tiles);
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? IPX_OK : fail(h, IPX_ERR_CUDA, "CUDA error %s while repeating a fold", cudaGetErrorName(e));
    });
}

int run_batch(ipx_handle *h, const BatchView &b, uint64_t n_cigar) {
    if (h->finalized) return fail(h, IPX_ERR_STATE, "ipx_push_batch after ipx_finalize_device without ipx_flush");
    if (h->targets.empty()) return fail(h, IPX_ERR_STATE, "ipx_set_targets must be called before ipx_push_batch");
    int rc = alloc_tables(h);
    if (rc) return rc;
    if (b.n >> 36) return fail(h, IPX_ERR_ARG, "batch too large");
    /* event buffer for the worst case: every emitted candidate consumes one N op and a following aligned op */
    const uint64_t ev_bound = n_cigar / 2 + 1;
    uint64_t n_tiles = tiles_of(b.n); /* (launch_extract says how many tiles it cut the batch into) */
    const uint64_t tiles_bound = max_tiles_of(b.n);
    rc = settle_batch(h); /* the previous batch's events are about to be overwritten */
    if (rc) return rc;
    /* long reads scatter their offsets over thousands of values: extra dense rows would be given out to stragglers and
     * then read by every finalize for nothing.  From the first such batch on, offsets >= DENSE_W go to the log. */
    if (n_cigar > 8 * b.n) h->tab.extra_rows = 0;
    rc = ensure_log(h, ev_bound);
    if (rc) return rc;
    if (h->params.deferred_counting && two_pass(h)) {
        /* K1 and the first-seen pass only; the events stay on the device until ipx_count_deferred */
        ipx_handle::Deferred d{nullptr, nullptr, n_tiles};
        CU(h, cudaMallocAsync(&d.events, ev_bound * sizeof(Event), h->stream));
        CU(h, cudaMallocAsync(&d.tile_info, (tiles_bound + 1) * sizeof(TileInfo), h->stream));
        h->deferred.push_back(d);
        {
            Span s(h, &h->tm_extract, 0);
            int launched = 0;
            CU(h, launch_extract(b, n_cigar, h->xp, h->tab, (Event *)d.events, (TileInfo *)d.tile_info, h->d_counters + 96,
                                 h->sm_count, h->stream, &launched, &n_tiles));
            h->tm_extract.launches += (uint64_t)launched;
            h->deferred.back().n_tiles = n_tiles;
        }
        {
            Span s(h, &h->tm_aggregate, 1);
            launch_fold(MODE_PASS1, h->tab, (Event *)d.events, (TileInfo *)d.tile_info, n_tiles, h->sm_count, h->stream);
        }
        CU(h, cudaGetLastError());
        h->totals.records += b.n;
        h->seq_hi = std::max<uint64_t>(h->seq_hi, b.seq_base + b.n);
        h->dirty = true;
        return IPX_OK;
    }
    CU(h, h->ev_buf.reserve(ev_bound * sizeof(Event), h->stream));
    CU(h, h->tile_cnt.reserve((tiles_bound + 1) * sizeof(TileInfo), h->stream));
    {
        Span s(h, &h->tm_extract, 0);
        int launched = 0;
        CU(h, launch_extract(b, n_cigar, h->xp, h->tab, (Event *)h->ev_buf.p, (TileInfo *)h->tile_cnt.p, h->d_counters + 96,
                             h->sm_count, h->stream, &launched, &n_tiles));
        h->tm_extract.launches += (uint64_t)launched;
    }
    fold_events(h, (Event *)h->ev_buf.p, (TileInfo *)h->tile_cnt.p, n_tiles);
    CU(h, cudaGetLastError());
    h->unsettled = true;
    h->unsettled_tiles = n_tiles;
    h->totals.records += b.n;
    h->seq_hi = std::max<uint64_t>(h->seq_hi, b.seq_base + b.n);
    h->dirty = true;
    return IPX_OK;
}

}  // namespace

/* ------------------------------------------------------------------ basic entry points */
extern "C" const char *ipx_version(void) { return "intronprospector_b200 0.1 (reference semantics: intronProspector 1.4.0)"; }
extern "C" int ipx_abi_version(void) { return IPX_ABI_VERSION; }

extern "C" int ipx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int ipx_create(const ipx_params *params, ipx_handle **out) {
    if (!params || !out) return IPX_ERR_ARG;
    *out = nullptr;
    int n = ipx_device_count();
    if (n <= 0 || params->device < 0 || params->device >= n) return IPX_ERR_CUDA;
    ipx_handle *h = new ipx_handle();
    h->params = *params;
    h->device = params->device;
    if (cudaSetDevice(h->device) != cudaSuccess) {
        delete h;
        return IPX_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("IPX_FIN_BOUNDS")) { /* tests: start the grow-only finalize bounds small, "B,L,R" */
        unsigned long long fb = 0, fl = 0, fr = 0;
        if (sscanf(e, "%llu,%llu,%llu", &fb, &fl, &fr) == 3 && fb && fl && fr) {
            h->fw.first_B = fb; h->fw.first_L = fl; h->fw.first_R = fr;
        }
    }
    if (params->stream) {
        h->stream = (cudaStream_t)params->stream;
    } else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete h;
            return IPX_ERR_CUDA;
        }
        h->own_stream = true;
    }
    {
        /* keep freed stream-ordered allocations in the pool instead of returning them to the driver */
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    h->xp.min_anchor_length = params->min_anchor_length;
    h->xp.max_anchor_indel_size = params->max_anchor_indel_size;
    h->xp.min_intron_length = params->min_intron_length;
    h->xp.max_intron_length = params->max_intron_length;
    h->xp.excludes = params->excludes;
    h->xp.strandness = params->strandness;
    *out = h;
    return IPX_OK;
}

static void mailbox_release(ipx_handle *h);
namespace { void merge_release(MergeWork &m); }

extern "C" void ipx_destroy(ipx_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    drain(h->tm_extract); drain(h->tm_aggregate); drain(h->tm_motif); drain(h->tm_finalize); drain(h->tm_merge); drain(h->tm_export); drain(h->tm_import);
    drop_deferred(h);
    cudaFree(h->tab.slots); cudaFree(h->tab.olog);
    free_rows(h->tab);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    cudaFree(h->d_counters); cudaFree(h->d_name_rank); cudaFree(h->d_target_seq); cudaFree(h->d_export);
    free_genome(h);
    h->s_tid.release(); h->s_pos.release(); h->s_flag.release(); h->s_nh.release(); h->s_xs.release();
    h->s_off.release(); h->s_cig.release(); h->s_nops.release(); h->s_runs.release(); h->s_tmp.release(); h->ev_buf.release(); h->tile_cnt.release(); h->jrow_of.release(); h->jid_of.release();
    finalize_release(h->fw);
    mailbox_release(h);
    merge_release(h->mw);
    for (auto &e : h->upload_done)
        if (e) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *ipx_last_error(const ipx_handle *h) { return h ? h->err.c_str() : "invalid handle"; }
extern "C" void *ipx_stream(ipx_handle *h) { return h ? (void *)h->stream : nullptr; }

extern "C" int ipx_set_targets(ipx_handle *h, int32_t n_targets, const char *const *names) {
    if (!h || n_targets < 0 || (n_targets && !names)) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    for (size_t i = 0; i < h->targets.size() && i < (size_t)n_targets; i++)
        if (h->targets[i] != names[i]) return fail(h, IPX_ERR_ARG, "target %zu changed name", i);
    /* the key holds as many bits of the target index as the intron length leaves; beyond that the tables compare the
     * full index stored with every junction (the "wide" layout, ipx_table.cuh) */
    if (h->tables_ready && make_key_layout(h->tab.kl.len_bits, (uint64_t)n_targets).wide) h->tab.kl.wide = 1;
    h->targets.assign(names, names + n_targets);
    return upload_target_maps(h);
}

extern "C" int ipx_set_genome(ipx_handle *h, int32_t n_seqs, const char *const *names, const uint64_t *lengths,
                              const char *const *seqs) {
    if (!h || n_seqs < 0) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    free_genome(h);
    h->gnames.assign(names, names + n_seqs);
    h->glen.assign(lengths, lengths + n_seqs);
    std::vector<const char *> ptrs(n_seqs ? n_seqs : 1, nullptr);
    for (int32_t i = 0; i < n_seqs; i++) {
        char *d = nullptr;
        CU(h, cudaMalloc(&d, lengths[i] ? lengths[i] : 1));
        h->d_gseq.push_back(d);
        CU(h, cudaMemcpyAsync(d, seqs[i], lengths[i], cudaMemcpyHostToDevice, h->stream));
        ptrs[i] = d;
    }
    CU(h, cudaMalloc((void **)&h->d_gseq_ptrs, ptrs.size() * sizeof(char *)));
    CU(h, cudaMalloc(&h->d_glen, ptrs.size() * sizeof(uint64_t)));
    CU(h, cudaMemcpyAsync((void *)h->d_gseq_ptrs, ptrs.data(), ptrs.size() * sizeof(char *), cudaMemcpyHostToDevice, h->stream));
    if (n_seqs) CU(h, cudaMemcpyAsync(h->d_glen, lengths, n_seqs * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_genome = true;
    return upload_target_maps(h);
}

namespace {
void mark_upload(ipx_handle *h) {
    cudaEvent_t &e = h->upload_done[h->n_uploads % ipx_handle::UPLOAD_RING];
    if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventRecord(e, h->stream);
    h->n_uploads++;
}
}  // namespace

extern "C" int ipx_push_batch_device(ipx_handle *h, const ipx_batch *b) {
    if (!h || !b) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    BatchView v{b->tid, b->pos, b->flag, b->nh_class, b->xs, b->cigar_off, b->cigar, b->n_records, b->seq_base};
    return run_batch(h, v, b->n_cigar);
}

extern "C" int ipx_push_batch(ipx_handle *h, const ipx_batch *b) {
    if (!h || !b) return IPX_ERR_ARG;
    if (b->n_records == 0) return IPX_OK;
    if (b->n_cigar >> 32) return fail(h, IPX_ERR_ARG, "more than 2^32 CIGAR words in one batch");
    cudaSetDevice(h->device);
    const uint64_t n = b->n_records;
    {
        /* before this batch's copies are enqueued, so that the wait is for work that finished long ago */
        int rcs = settle_batch(h);
        if (rcs) return rcs;
    }
    CU(h, h->s_tid.reserve(n * 4, h->stream));
    CU(h, h->s_pos.reserve(n * 4, h->stream));
    CU(h, h->s_flag.reserve(n * 2, h->stream));
    CU(h, h->s_nh.reserve(n, h->stream));
    CU(h, h->s_xs.reserve(n, h->stream));
    CU(h, h->s_off.reserve((n + 1) * 4, h->stream));
    CU(h, h->s_cig.reserve((b->n_cigar ? b->n_cigar : 1) * 4, h->stream));
    /* the staging buffers may still be read by the previous batch's kernels: same stream, so ordered */
    CU(h, cudaMemcpyAsync(h->s_tid.p, b->tid, n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->s_pos.p, b->pos, n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->s_flag.p, b->flag, n * 2, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->s_nh.p, b->nh_class, n, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->s_xs.p, b->xs, n, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->s_off.p, b->cigar_off, (n + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    if (b->n_cigar) CU(h, cudaMemcpyAsync(h->s_cig.p, b->cigar, b->n_cigar * 4, cudaMemcpyHostToDevice, h->stream));
    mark_upload(h);
    BatchView v{(const int32_t *)h->s_tid.p, (const int32_t *)h->s_pos.p, (const uint16_t *)h->s_flag.p,
                (const uint8_t *)h->s_nh.p, (const uint8_t *)h->s_xs.p, (const uint32_t *)h->s_off.p,
                (const uint32_t *)h->s_cig.p, n, b->seq_base};
    return run_batch(h, v, b->n_cigar);
}

namespace {
int enqueue_finalize(ipx_handle *h) {
    Span s(h, &h->tm_finalize, 0);
    cudaError_t e = finalize_enqueue(h->tab, genome_view(h), h->d_name_rank, h->fw, h->stream);
    h->tm_finalize.launches += (uint64_t)h->fw.launches_per_replay;
    if (e != cudaSuccess)
        return fail(h, IPX_ERR_CUDA, "CUDA error %s in finalize: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return IPX_OK;
}

int settle_import(ipx_handle *h, bool *repeated);

/* wait for the enqueued finalize and make sure its buffers were large enough (growing and replaying otherwise) */
int settle_finalize(ipx_handle *h) {
    if (h->finalize_settled) return IPX_OK;
    {
        /* a table that was full while the last batch was folded, or too small for what other ranks sent: the finalize
         * that was enqueued saw a partial table */
        bool repeated = false, repeated_import = false;
        int rc = settle_batch(h, &repeated);
        if (rc) return rc;
        rc = settle_import(h, &repeated_import);
        if (rc) return rc;
        if (repeated || repeated_import) {
            rc = enqueue_finalize(h);
            if (rc) return rc;
        }
    }
    for (int round = 0; round < 12; round++) {
        CU(h, cudaStreamSynchronize(h->stream));
        if (finalize_check(h->fw)) {
            h->finalize_settled = true;
            h->bounds_proven = true;
            return IPX_OK;
        }
        h->bounds_proven = false;
        int rc = enqueue_finalize(h);
        if (rc) return rc;
    }
    return fail(h, IPX_ERR_CUDA, "finalize did not settle");
}
}  // namespace

extern "C" int ipx_push_batch_compact(ipx_handle *h, const ipx_batch_compact *b) {
    if (!h || !b) return IPX_ERR_ARG;
    if (b->n_records == 0) return IPX_OK;
    if (b->n_cigar >> 32) return fail(h, IPX_ERR_ARG, "more than 2^32 CIGAR words in one batch");
    if (b->n_records >> 31) return fail(h, IPX_ERR_ARG, "more than 2^31 records in one compact batch");
    if (b->n_tid_runs == 0 || !b->run_tid || !b->run_start || b->run_start[0] != 0)
        return fail(h, IPX_ERR_ARG, "target runs must cover the batch from record 0");
    cudaSetDevice(h->device);
    const uint64_t n = b->n_records, nr = b->n_tid_runs;
    cudaStream_t st = h->stream;
    {
        /* before this batch's copies are enqueued, so that the wait is for work that finished long ago */
        int rcs = settle_batch(h);
        if (rcs) return rcs;
    }
    CU(h, h->s_tid.reserve(n * 4, st));
    CU(h, h->s_pos.reserve(n * 4, st));
    CU(h, h->s_flag.reserve(n * 2, st));
    CU(h, h->s_nh.reserve(n, st));
    CU(h, h->s_xs.reserve(n, st));
    CU(h, h->s_off.reserve((n + 1) * 4, st));
    CU(h, h->s_cig.reserve((b->n_cigar ? b->n_cigar : 1) * 4, st));
    CU(h, h->s_nops.reserve(n * 2, st));
    CU(h, h->s_runs.reserve(nr * 16, st));
    size_t tmp_bytes = 0;
    CU(h, expand_compact(nullptr, n, nullptr, nullptr, nullptr, 0, nullptr, nullptr, &tmp_bytes, st));
    CU(h, h->s_tmp.reserve(tmp_bytes, st));
    /* same stream as the previous batch's kernels, so the staging buffers are free again when these copies run */
    CU(h, cudaMemcpyAsync(h->s_nops.p, b->n_cigar_ops, n * 2, cudaMemcpyHostToDevice, st));
    unsigned long long *d_run_start = (unsigned long long *)h->s_runs.p;
    int32_t *d_run_tid = (int32_t *)((char *)h->s_runs.p + nr * 8);
    CU(h, cudaMemcpyAsync(d_run_start, b->run_start, nr * 8, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(d_run_tid, b->run_tid, nr * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->s_pos.p, b->pos, n * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->s_flag.p, b->flag, n * 2, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->s_nh.p, b->nh_class, n, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->s_xs.p, b->xs, n, cudaMemcpyHostToDevice, st));
    if (b->n_cigar) CU(h, cudaMemcpyAsync(h->s_cig.p, b->cigar, b->n_cigar * 4, cudaMemcpyHostToDevice, st));
    mark_upload(h);
    CU(h, expand_compact((const uint16_t *)h->s_nops.p, n, (uint32_t *)h->s_off.p, d_run_tid, d_run_start, nr, (int32_t *)h->s_tid.p,
                         h->s_tmp.p, &tmp_bytes, st));
    BatchView v{(const int32_t *)h->s_tid.p, (const int32_t *)h->s_pos.p, (const uint16_t *)h->s_flag.p,
                (const uint8_t *)h->s_nh.p, (const uint8_t *)h->s_xs.p, (const uint32_t *)h->s_off.p,
                (const uint32_t *)h->s_cig.p, n, b->seq_base};
    return run_batch(h, v, b->n_cigar);
}

extern "C" void *ipx_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void ipx_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" int ipx_sync_uploads(ipx_handle *h, uint32_t keep_in_flight) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->n_uploads <= keep_in_flight) return IPX_OK;
    if (keep_in_flight >= (uint32_t)ipx_handle::UPLOAD_RING) return fail(h, IPX_ERR_ARG, "at most %d uploads are tracked", ipx_handle::UPLOAD_RING - 1);
    const uint64_t last = h->n_uploads - 1 - keep_in_flight; /* this push and every earlier one must be done */
    CU(h, cudaEventSynchronize(h->upload_done[last % ipx_handle::UPLOAD_RING]));
    return IPX_OK;
}

extern "C" int ipx_finalize_device(ipx_handle *h) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->finalized) return IPX_OK;
    if (!h->deferred.empty()) return fail(h, IPX_ERR_STATE, "junction events are waiting to be counted: call ipx_count_deferred first");
    if (h->targets.empty()) return fail(h, IPX_ERR_STATE, "ipx_set_targets must be called first");
    int rc = alloc_tables(h);
    if (rc) return rc;
    rc = enqueue_finalize(h);
    if (rc) return rc;
    h->finalized = true;
    h->finalize_settled = false;
    /* asynchronous once the buffers have proven large enough; ipx_flush checks every time */
    if (!h->bounds_proven) return settle_finalize(h);
    return IPX_OK;
}

namespace {

template <class T>
cudaError_t d2h(std::vector<T> &dst, const void *src, size_t n, cudaStream_t st) {
    dst.resize(n);
    if (n == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, st);
}

/* text of the exception the reference throws for a deferred device-side error */
void describe_error(ipx_handle *h, unsigned long long err_sv, HostTable *ht) {
    const unsigned code = (unsigned)(err_sv & 0xFFull);
    ht->t.first_error_seq = err_sv >> SV_SEQ_SHIFT;
    ht->t.error_tid = -1;
    if (code < 0x80u) {
        ht->t.status = IPX_ERR_CIGAR;
        const char *ops = "MIDNSHP=XB";
        snprintf(ht->t.error_text, sizeof(ht->t.error_text), "Unknown cigar operation '%c'", code < 10 ? ops[code] : '?');
        return;
    }
    ht->t.status = IPX_ERR_GENOME;
    /* find the junction whose first event raised it: only its key is needed, scan the junction rows in use on the host
     * (errors are fatal, so this path runs at most once per run) */
    unsigned long long nk = 0;
    cudaMemcpy(&nk, h->tab.n_keys, sizeof(nk), cudaMemcpyDeviceToHost);
    const uint64_t n = std::min<uint64_t>(nk, h->tab.max_j);
    std::vector<unsigned long long> keys(n);
    std::vector<int32_t> tids(n);
    std::vector<JStat> stat(n);
    cudaMemcpy(keys.data(), h->tab.jkey, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(tids.data(), h->tab.jtid, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaMemcpy(stat.data(), h->tab.stat, n * sizeof(JStat), cudaMemcpyDeviceToHost);
    const int status = (int)(code & 0x7Fu);
    for (uint64_t i = 0; i < n; i++) {
        unsigned long long fm = std::min(stat[i].first[0], std::min(stat[i].first[1], stat[i].first[2]));
        if (fm == EMPTY64 || (fm & ~0xFFull) != (err_sv & ~0xFFull)) continue;
        const int32_t tid = tids[i];
        const int64_t i0 = key_start(h->tab.kl, keys[i]), i1 = i0 + key_len(h->tab.kl, keys[i]);
        ht->t.error_tid = tid;
        const char *chrom = h->targets[tid].c_str();
        auto clipped = [&](int64_t s, int64_t e) {
            const int32_t g = h->target_seq[tid];
            int64_t len = g >= 0 ? (int64_t)h->glen[g] : 0, b = s, en = e - 1;
            if (en < b - 1) en = b - 1;
            if (b < 0) b = 0;
            if (b >= len) b = len;
            if (en < 0) en = 0; else if (en >= len) en = len - 1;
            int64_t got = en - b + 1;
            return got < 0 ? (int64_t)0 : got;
        };
        if (status == 2) /* MOTIF_ERR_MISSING: Genome::fetch, genome.cc:64-67 */
            snprintf(ht->t.error_text, sizeof(ht->t.error_text), "can't load genome sequence for: %s:%d-%d", chrom, (int)i0, (int)(i0 + 2));
        else if (status == 3) /* genome.cc:68-70 */
            snprintf(ht->t.error_text, sizeof(ht->t.error_text), "unexpected length returned for: %s:%d-%d: %d", chrom, (int)i0,
                     (int)(i0 + 2), (int)clipped(i0, i0 + 2));
        else
            snprintf(ht->t.error_text, sizeof(ht->t.error_text), "unexpected length returned for: %s:%d-%d: %d", chrom,
                     (int)(i1 - 2), (int)i1, (int)clipped(i1 - 2, i1));
        return;
    }
    snprintf(ht->t.error_text, sizeof(ht->t.error_text), "genome sequence error");
}

}  // namespace

namespace {
struct Trace {
    bool on = getenv("IPX_TRACE") != nullptr;
    double t0 = now();
    static double now() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    void mark(const char *what) {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[ipx] %-28s %8.3f ms\n", what, t - t0);
        t0 = t;
    }
};
}  // namespace

extern "C" int ipx_flush(ipx_handle *h, ipx_table **out) {
    if (!h || !out) return IPX_ERR_ARG;
    *out = nullptr;
    Trace tr;
    int rc = ipx_finalize_device(h);
    if (rc) return rc;
    rc = settle_finalize(h);
    if (rc) return rc;
    if (h->late_fold)
        return fail(h, IPX_ERR_STATE, "a table was full when the partial tables were exchanged: what was folded afterwards may belong "
                    "to another rank -- call ipx_exchange_settle on every rank and repeat the exchange before ipx_flush");
    tr.mark("flush: wait for the device");
    cudaStream_t st = h->stream;
    HostTable *ht = new HostTable();
    memset(&ht->t, 0, sizeof(ht->t));
    const DeviceTable &d = h->fw.out;
    const uint64_t n = d.n;
    /* every column goes into one pinned block at link speed; the table then owns that block */
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    struct Col { const void *src; size_t bytes; size_t at; };
    enum { C_CHROM, C_I0, C_I1, C_A0, C_A1, C_IJUNC, C_STRAND, C_SITES, C_COUNTS, C_T, C_HOFF, C_HCNT, C_CONF, N_COLS };
    Col cols[N_COLS] = {{d.chrom, n * 4, 0}, {d.i0, n * 4, 0}, {d.i1, n * 4, 0}, {d.a0, n * 4, 0}, {d.a1, n * 4, 0}, {d.ijunc, n * 4, 0},
                        {d.strand, n, 0}, {d.sites, n * 8, 0}, {d.counts, n * 24, 0}, {d.T, n * 8, 0}, {d.hist_off, (n + 1) * 8, 0},
                        {d.hist_cnt, d.n_hist * 4, 0}, {nullptr, n * 4, 0}};
    size_t need = 64;
    for (auto &c : cols) {
        c.at = need;
        need += (c.bytes + 63) & ~(size_t)63;
    }
    ht->block = g_pinned_pool.take(need, &ht->block_cap);
    if (!ht->block) {
        delete ht;
        return fail(h, IPX_ERR_CUDA, "cannot allocate %zu bytes of pinned host memory for the table", need);
    }
    char *stage = (char *)ht->block;
    for (auto &c : cols)
        if (c.bytes && c.src) chk(cudaMemcpyAsync(stage + c.at, c.src, c.bytes, cudaMemcpyDeviceToHost, st));
    unsigned long long ctr[4] = {0, 0, 0, 0};
    unsigned long long err_sv = EMPTY64;
    uint32_t overflow = 0;
    chk(cudaMemcpyAsync(&ctr[0], h->tab.n_events, 8, cudaMemcpyDeviceToHost, st));
    chk(cudaMemcpyAsync(&ctr[1], h->tab.n_live, 8, cudaMemcpyDeviceToHost, st));
    chk(cudaMemcpyAsync(&err_sv, h->tab.err_sv, 8, cudaMemcpyDeviceToHost, st));
    chk(cudaMemcpyAsync(&overflow, h->tab.overflow, 4, cudaMemcpyDeviceToHost, st));
    chk(cudaStreamSynchronize(st));
    tr.mark("flush: table to host");
    if (e != cudaSuccess) {
        delete ht;
        return fail(h, IPX_ERR_CUDA, "CUDA error %s in flush: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    }
    if (overflow) {
        delete ht;
        if (overflow & OVF_START) return fail(h, IPX_ERR_CAPACITY, "intron start beyond 2^31 cannot be keyed");
        if (overflow & OVF_MAIL) return fail(h, IPX_ERR_CAPACITY, "a mailbox slot was too small for the rows sent to it (ipx_mailbox_create)");
        if (overflow & OVF_PEER) return fail(h, IPX_ERR_STATE, "a rank never published its share of the exchange (ipx_partial_export_mail)");
        /* (a full table is grown and the fold repeated before results are read: settle_batch) */
        return fail(h, IPX_ERR_CAPACITY, "%s is full", (overflow & OVF_TABLE) ? "junction table" : "read-offset log");
    }
    ipx_table &tt = ht->t;
    tt.n_junctions = n;
    tt.chrom = (const int32_t *)(stage + cols[C_CHROM].at);
    tt.intron_start = (const uint32_t *)(stage + cols[C_I0].at);
    tt.intron_end = (const uint32_t *)(stage + cols[C_I1].at);
    tt.anchor_start = (const uint32_t *)(stage + cols[C_A0].at);
    tt.anchor_end = (const uint32_t *)(stage + cols[C_A1].at);
    tt.ijunc = (const uint32_t *)(stage + cols[C_IJUNC].at);
    tt.strand = stage + cols[C_STRAND].at;
    tt.splice_sites = stage + cols[C_SITES].at;
    tt.read_counts = (const uint64_t *)(stage + cols[C_COUNTS].at);
    tt.creation_seq = (const uint64_t *)(stage + cols[C_T].at);
    tt.hist_off = (const uint64_t *)(stage + cols[C_HOFF].at);
    tt.hist_cnt = (const uint32_t *)(stage + cols[C_HCNT].at);
    tt.n_hist = d.n_hist;
    tt.confidence = (const float *)(stage + cols[C_CONF].at);
    if (n == 0) *(uint64_t *)(stage + cols[C_HOFF].at) = 0;
    unsigned long long *T = (unsigned long long *)(stage + cols[C_T].at);
    uint32_t *ijunc = (uint32_t *)(stage + cols[C_IJUNC].at);
    /* sjN continues across flushes (ijunc_, junctions_extractor.hh:129) */
    for (uint64_t i = 0; i < n; i++) ijunc[i] += h->ijunc_base;
    h->ijunc_base += (uint32_t)n;
    tr.mark("flush: host columns");
    compute_confidences(n, tt.read_counts, tt.hist_off, tt.hist_cnt, (float *)(stage + cols[C_CONF].at));
    tr.mark("flush: confidences");
    /* missing-genome warnings in first-miss order (junctions_extractor.cc:411-417) */
    if (h->have_genome && h->params.skip_missing_targets) {
        std::unordered_map<int32_t, unsigned long long> first_miss;
        for (uint64_t i = 0; i < n; i++) {
            const int32_t tid = tt.chrom[i];
            if (h->target_seq[tid] >= 0) continue;
            auto it = first_miss.find(tid);
            if (it == first_miss.end() || T[i] < it->second) first_miss[tid] = T[i];
        }
        std::vector<std::pair<unsigned long long, int32_t>> v;
        for (auto &kv : first_miss) v.emplace_back(kv.second, kv.first);
        std::sort(v.begin(), v.end());
        for (auto &kv : v)
            if (h->warned_missing.insert(h->targets[kv.second]).second) ht->missing.push_back(kv.second);
    }
    for (uint64_t i = 0; i < n; i++) T[i] >>= 8; /* sequence value -> creation_seq (drops the strand byte) */
    ht->bind();
    ht->t.status = IPX_OK;
    if (err_sv != EMPTY64) describe_error(h, err_sv, ht);
    h->totals.junctions += n;
    h->totals.hist_runs += d.n_hist;
    /* clear(), junctions.hh:240-245 */
    h->finalized = false;
    launch_clear_tables(h->tab, st);
    h->tab.extra_rows = 1;
    h->log_bound = 0;
    h->unsettled = false;
    h->exported_unsettled = false;
    h->dirty = false;
    *out = &ht->t;
    return IPX_OK;
}

/* JunctionsExtractor::clear (junctions_extractor.hh:222-224) without reading the table back */
extern "C" int ipx_clear(ipx_handle *h) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    h->finalized = false;
    if (h->tables_ready) {
        Span s(h, &h->tm_finalize, 2);
        /* (a table that was full while the last batch was folded is not repaired: everything is dropped anyway) */
        launch_clear_tables(h->tab, h->stream);
        CU(h, cudaGetLastError());
    }
    drop_deferred(h);
    h->tab.extra_rows = 1;
    h->log_bound = 0;
    h->unsettled = false;
    h->import_pending = false;
    h->exported_unsettled = h->late_fold = false;
    h->dirty = false;
    return IPX_OK;
}

extern "C" int ipx_table_set_ijunc(ipx_table *t, const uint32_t *ijunc) {
    if (!t || (t->n_junctions && !ijunc)) return IPX_ERR_ARG;
    HostTable *ht = reinterpret_cast<HostTable *>(t);
    ht->ijunc.assign(ijunc, ijunc + t->n_junctions);
    ht->bind();
    if (ht->block && t->n_junctions == 0) t->ijunc = ht->ijunc.data();
    return IPX_OK;
}

extern "C" void ipx_table_free(ipx_table *t) {
    if (t) delete reinterpret_cast<HostTable *>(t);
}

/* ------------------------------------------------------------------ formatters */
static int give_text(const std::string &s, ipx_text *out) {
    out->data = (char *)malloc(s.size() + 1);
    if (!out->data) return IPX_ERR_ARG;
    memcpy(out->data, s.data(), s.size());
    out->data[s.size()] = 0;
    out->len = s.size();
    return IPX_OK;
}

extern "C" int ipx_format(const ipx_table *t, const char *const *names, int format, float min_confidence, uint64_t begin,
                          uint64_t end, ipx_text *out) {
    if (!t || !names || !out) return IPX_ERR_ARG;
    std::string s;
    int rc = format_rows(t, names, format, min_confidence, begin, end, s);
    if (rc) return rc;
    return give_text(s, out);
}

extern "C" int ipx_format_tsv_header(ipx_text *out) {
    if (!out) return IPX_ERR_ARG;
    return give_text(tsv_header(), out);
}

extern "C" void ipx_text_free(ipx_text *t) {
    if (t && t->data) free(t->data);
    if (t) {
        t->data = nullptr;
        t->len = 0;
    }
}

/* ------------------------------------------------------------------ timings */
extern "C" int ipx_get_timings(ipx_handle *h, ipx_timings *out) {
    if (!h || !out) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    drain(h->tm_extract); drain(h->tm_aggregate); drain(h->tm_motif); drain(h->tm_finalize); drain(h->tm_merge); drain(h->tm_export); drain(h->tm_import);
    ipx_timings t = h->totals;
    t.extract_ms = h->tm_extract.ms;
    t.aggregate_ms = h->tm_aggregate.ms;
    t.motif_ms = h->tm_motif.ms;
    t.finalize_ms = h->tm_finalize.ms;
    t.merge_ms = h->tm_merge.ms;
    t.extract_launches = h->tm_extract.launches;
    t.aggregate_launches = h->tm_aggregate.launches;
    t.motif_launches = h->tm_motif.launches;
    t.finalize_launches = h->tm_finalize.launches;
    t.merge_launches = h->tm_merge.launches;
    t.export_ms = h->tm_export.ms;
    t.import_ms = h->tm_import.ms;
    t.exchange_launches = h->tm_export.launches + h->tm_import.launches;
    if (h->tables_ready) {
        /* the device counters are cumulative over the life of the handle */
        unsigned long long ev = 0, lw = 0;
        cudaMemcpy(&ev, h->tab.n_events, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&lw, h->tab.n_live, 8, cudaMemcpyDeviceToHost);
        t.events = ev - h->events_base;
        t.cigar_live = lw - h->live_base;
    }
    *out = t;
    return IPX_OK;
}

extern "C" int ipx_reset_timings(ipx_handle *h) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    drain(h->tm_extract); drain(h->tm_aggregate); drain(h->tm_motif); drain(h->tm_finalize); drain(h->tm_merge); drain(h->tm_export); drain(h->tm_import);
    h->tm_extract = StageTimer(); h->tm_aggregate = StageTimer(); h->tm_motif = StageTimer(); h->tm_finalize = StageTimer(); h->tm_merge = StageTimer();
    h->tm_export = StageTimer(); h->tm_import = StageTimer();
    h->totals = ipx_timings{};
    if (h->tables_ready) {
        cudaMemcpy(&h->events_base, h->tab.n_events, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&h->live_base, h->tab.n_live, 8, cudaMemcpyDeviceToHost);
    }
    return IPX_OK;
}

/* ------------------------------------------------------------------ multi-GPU exchange */
extern "C" int ipx_partial_row_bytes(uint32_t *jb, uint32_t *hb) {
    if (jb) *jb = (uint32_t)sizeof(JunctionRow);
    if (hb) *hb = (uint32_t)sizeof(HistRow);
    return IPX_OK;
}

extern "C" int ipx_partial_sizes(ipx_handle *h, int32_t n_parts, const int32_t *split_tid, const uint32_t *split_start,
                                 uint64_t *junction_rows, uint64_t *hist_rows) {
    if (!h || n_parts < 1 || n_parts > MAX_PARTS || !junction_rows || !hist_rows) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (two_pass(h) && !h->params.deferred_counting)
        return fail(h, IPX_ERR_STATE, "multi-GPU exchange with the canon filter needs ipx_params.deferred_counting");
    if (!h->deferred.empty()) return fail(h, IPX_ERR_STATE, "call ipx_count_deferred before the exchange");
    if (h->finalized) return fail(h, IPX_ERR_STATE, "the exchange works on the table before ipx_finalize_device");
    if (h->targets.empty()) return fail(h, IPX_ERR_STATE, "ipx_set_targets must be called first");
    int rc = alloc_tables(h);
    if (rc) return rc;
    rc = settle_batch(h);
    if (rc) return rc;
    /* splitters in key-order terms: chromosome-name rank and intron_start */
    std::vector<uint32_t> order(h->targets.size()), rank(h->targets.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h->targets[a] < h->targets[b]; });
    for (size_t i = 0; i < order.size(); i++) rank[order[i]] = (uint32_t)i;
    Splitters &sp = h->export_split;
    sp.n_parts = n_parts;
    for (int32_t p = 0; p + 1 < n_parts; p++) {
        if (split_tid[p] < 0 || (size_t)split_tid[p] >= rank.size()) return fail(h, IPX_ERR_ARG, "splitter %d names no target", p);
        sp.rank[p] = rank[split_tid[p]];
        sp.start[p] = split_start[p];
        if (p > 0 && (sp.rank[p] < sp.rank[p - 1] || (sp.rank[p] == sp.rank[p - 1] && sp.start[p] < sp.start[p - 1])))
            return fail(h, IPX_ERR_ARG, "splitters are not in key order");
    }
    cudaStream_t st = h->stream;
    if (!h->d_export) CU(h, cudaMalloc(&h->d_export, (2 * MAX_PARTS + 4) * sizeof(unsigned long long)));
    /* The two passes are launched for grow-only bounds on the number of keys and log entries (the kernels read the
     * real numbers on the device); one host round trip brings back the row counts together with the real numbers, and
     * only a bound that turned out too small costs a second one. */
    for (int round = 0; round < 2; round++) {
        CU(h, cudaMemsetAsync(h->d_export, 0, 2 * MAX_PARTS * sizeof(unsigned long long), st));
        {
            /* creation thresholds for anything still unresolved: keys that never got one are not exported */
            Span s(h, &h->tm_motif, 1);
            launch_resolve_T(h->tab, genome_view(h), st);
        }
        {
            Span s(h, &h->tm_finalize, 2);
            launch_export_count(h->tab, h->d_name_rank, sp, h->export_keys, h->export_olog, h->d_export, st);
        }
        CU(h, cudaGetLastError());
        CU(h, cudaMemcpyAsync(h->d_export + 2 * MAX_PARTS, h->tab.n_keys, 8, cudaMemcpyDeviceToDevice, st));
        CU(h, cudaMemcpyAsync(h->d_export + 2 * MAX_PARTS + 1, h->tab.n_olog, 8, cudaMemcpyDeviceToDevice, st));
        CU(h, cudaMemcpyAsync(h->d_export + 2 * MAX_PARTS + 2, h->tab.overflow, 4, cudaMemcpyDeviceToDevice, st));
        CU(h, cudaMemcpyAsync(h->export_counts, h->d_export, (2 * MAX_PARTS + 4) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        const unsigned long long nk = h->export_counts[2 * MAX_PARTS], nl = h->export_counts[2 * MAX_PARTS + 1];
        if ((uint32_t)h->export_counts[2 * MAX_PARTS + 2] & (OVF_TABLE | OVF_LOG))
            return fail(h, IPX_ERR_CAPACITY, "a table ran out of room before the exchange");
        if (nk <= h->export_keys && nl <= h->export_olog) break;
        h->export_keys = nk + nk / 8 + 4096;
        h->export_olog = nl + nl / 8 + 4096;
    }
    for (int32_t p = 0; p < n_parts; p++) {
        junction_rows[p] = h->export_counts[p];
        hist_rows[p] = h->export_counts[n_parts + p];
    }
    h->export_ready = true;
    return IPX_OK;
}

extern "C" int ipx_partial_export(ipx_handle *h, void *dev_junction_rows, void *dev_hist_rows) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (!h->export_ready) return fail(h, IPX_ERR_STATE, "call ipx_partial_sizes first");
    cudaStream_t st = h->stream;
    const int n_parts = h->export_split.n_parts;
    /* destination-major: the cursor of a part starts where the parts before it end */
    unsigned long long *cursors = h->export_cursors;
    unsigned long long j = 0, hh = 0;
    for (int p = 0; p < n_parts; p++) {
        cursors[p] = j;
        h->export_split.jbase[p] = (uint32_t)j;
        cursors[n_parts + p] = hh;
        j += h->export_counts[p];
        hh += h->export_counts[n_parts + p];
    }
    CU(h, h->jrow_of.reserve(h->tab.max_j * sizeof(uint32_t), st));
    CU(h, cudaMemcpyAsync(h->d_export, cursors, 2 * n_parts * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    {
        Span s(h, &h->tm_export, 2);
        launch_export_scatter(h->tab, h->d_name_rank, h->export_split, h->export_keys, h->export_olog, h->d_export,
                              (uint32_t *)h->jrow_of.p, (JunctionRow *)dev_junction_rows, (HistRow *)dev_hist_rows, st);
        /* the local table has been handed over: empty it so that imports (own share included) rebuild it */
        launch_clear_tables(h->tab, st);
        h->log_bound = 0;
    }
    CU(h, cudaGetLastError());
    h->export_ready = false;
    return IPX_OK;
}

extern "C" int ipx_partial_import(ipx_handle *h, const void *dev_junction_rows, uint64_t n_junction_rows,
                                  const void *dev_hist_rows, uint64_t n_hist_rows) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->finalized) return fail(h, IPX_ERR_STATE, "table already finalized");
    int rc = alloc_tables(h);
    if (rc) return rc;
    /* room for everything that arrives, ahead of the launches (the rows are J-sized: exact growth is cheap) */
    {
        TableState ts;
        rc = read_table_state(h, &ts);
        if (rc) return rc;
        uint64_t in_use = std::min<uint64_t>(ts.n_keys, h->tab.max_j);
        bool rebuilt = false;
        while (in_use + n_junction_rows > h->tab.max_j) {
            rc = grow_rows(h, in_use);
            if (rc) return rc;
            rebuilt = true;
        }
        while (2 * (in_use + n_junction_rows) > h->tab.cap_mask + 1) {
            rc = rebuild_slots(h, true);
            if (rc) return rc;
            rebuilt = false;
        }
        if (rebuilt) { /* (the rows moved, the slots did not: nothing to do) */ }
        rc = ensure_log(h, n_hist_rows);
        if (rc) return rc;
    }
    CU(h, h->jid_of.reserve((n_junction_rows + 1) * sizeof(uint32_t), h->stream));
    {
        Span s(h, &h->tm_import, 2);
        launch_import_rows(h->tab, (const JunctionRow *)dev_junction_rows, n_junction_rows, (const HistRow *)dev_hist_rows,
                           n_hist_rows, (uint32_t *)h->jid_of.p, h->stream);
    }
    CU(h, cudaGetLastError());
    h->dirty = true;
    return IPX_OK;
}

/* ------------------------------------------------------------------ multi-GPU exchange over peer memory */
static int splitters_in_key_order(ipx_handle *h, int32_t n_parts, const int32_t *split_tid, const uint32_t *split_start, Splitters &sp) {
    std::vector<uint32_t> order(h->targets.size()), rank(h->targets.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h->targets[a] < h->targets[b]; });
    for (size_t i = 0; i < order.size(); i++) rank[order[i]] = (uint32_t)i;
    sp.n_parts = n_parts;
    for (int32_t p = 0; p + 1 < n_parts; p++) {
        if (split_tid[p] < 0 || (size_t)split_tid[p] >= rank.size()) return fail(h, IPX_ERR_ARG, "splitter %d names no target", p);
        sp.rank[p] = rank[split_tid[p]];
        sp.start[p] = split_start[p];
        if (p > 0 && (sp.rank[p] < sp.rank[p - 1] || (sp.rank[p] == sp.rank[p - 1] && sp.start[p] < sp.start[p - 1])))
            return fail(h, IPX_ERR_ARG, "splitters are not in key order");
    }
    return IPX_OK;
}

static void mailbox_release(ipx_handle *h) {
    for (int r = 0; r < MAX_PARTS; r++) {
        if (h->mail_opened[r]) cudaIpcCloseMemHandle(h->mail.base[r]);
        h->mail_opened[r] = false;
        h->mail.base[r] = nullptr;
    }
    if (h->mail_own) cudaFree(h->mail_own);
    if (h->mail.outbox) cudaFree(h->mail.outbox);
    h->mail_own = nullptr;
    h->mail.outbox = nullptr;
    h->mail_connected = false;
}

extern "C" int ipx_mailbox_create(ipx_handle *h, int32_t world, int32_t rank, uint64_t junction_rows, uint64_t hist_rows) {
    if (!h || world < 1 || world > MAX_PARTS || rank < 0 || rank >= world || !junction_rows || !hist_rows) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    mailbox_release(h);
    Mailboxes &m = h->mail;
    m.world = world;
    m.me = rank;
    m.parity = 0;
    m.epoch = 0;
    if (junction_rows >= (1ull << 24) - 1) return fail(h, IPX_ERR_ARG, "at most 2^24 - 2 junction rows per mailbox slot");
    m.cap_j = junction_rows;
    m.cap_h = hist_rows;
    m.hist_offset = (junction_rows * sizeof(JunctionRow) + 255) & ~(uint64_t)255;
    m.slot_bytes = MAIL_HEADER_BYTES + m.hist_offset + ((hist_rows * sizeof(HistRow) + 255) & ~(uint64_t)255);
    m.slot_bytes = (m.slot_bytes + 255) & ~(uint64_t)255;
    /* a plain cudaMalloc: memory from a stream-ordered pool cannot be shared through cudaIpcGetMemHandle */
    CU(h, cudaMalloc(&h->mail_own, 2 * (size_t)world * m.slot_bytes));
    CU(h, cudaMemset(h->mail_own, 0, 2 * (size_t)world * m.slot_bytes));
    CU(h, cudaMalloc(&m.outbox, (size_t)world * m.slot_bytes));
    m.base[rank] = h->mail_own;
    return IPX_OK;
}

extern "C" int ipx_mailbox_ipc_handle(ipx_handle *h, void *handle_64_bytes) {
    if (!h || !handle_64_bytes) return IPX_ERR_ARG;
    if (!h->mail_own) return fail(h, IPX_ERR_STATE, "call ipx_mailbox_create first");
    cudaSetDevice(h->device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t ipc;
    CU(h, cudaIpcGetMemHandle(&ipc, h->mail_own));
    memcpy(handle_64_bytes, &ipc, sizeof(ipc));
    return IPX_OK;
}

extern "C" void *ipx_mailbox_ptr(ipx_handle *h) { return h ? h->mail_own : nullptr; }

extern "C" int ipx_mailbox_connect(ipx_handle *h, int32_t ipc, const void *peers) {
    if (!h || !peers) return IPX_ERR_ARG;
    if (!h->mail_own) return fail(h, IPX_ERR_STATE, "call ipx_mailbox_create first");
    cudaSetDevice(h->device);
    Mailboxes &m = h->mail;
    for (int r = 0; r < m.world; r++) {
        if (r == m.me) continue;
        if (ipc) {
            cudaIpcMemHandle_t hd;
            memcpy(&hd, (const char *)peers + 64 * (size_t)r, sizeof(hd));
            void *p = nullptr;
            CU(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
            m.base[r] = (char *)p;
            h->mail_opened[r] = true;
        } else {
            m.base[r] = (char *)((void *const *)peers)[r];
            if (!m.base[r]) return fail(h, IPX_ERR_ARG, "mailbox pointer of rank %d is null", r);
        }
    }
    h->mail_connected = true;
    return IPX_OK;
}

extern "C" int ipx_partial_export_mail(ipx_handle *h, const int32_t *split_tid, const uint32_t *split_start) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (!h->mail_connected) return fail(h, IPX_ERR_STATE, "call ipx_mailbox_connect first");
    if (two_pass(h) && !h->params.deferred_counting)
        return fail(h, IPX_ERR_STATE, "multi-GPU exchange with the canon filter needs ipx_params.deferred_counting");
    if (!h->deferred.empty()) return fail(h, IPX_ERR_STATE, "call ipx_count_deferred before the exchange");
    if (h->finalized) return fail(h, IPX_ERR_STATE, "the exchange works on the table before ipx_finalize_device");
    if (h->targets.empty()) return fail(h, IPX_ERR_STATE, "ipx_set_targets must be called first");
    const int world = h->mail.world;
    if (world > 1 && (!split_tid || !split_start)) return IPX_ERR_ARG;
    int rc = alloc_tables(h);
    if (rc) return rc;
    Splitters &sp = h->export_split;
    rc = splitters_in_key_order(h, world, split_tid, split_start, sp);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    if (!h->d_export) CU(h, cudaMalloc(&h->d_export, (2 * MAX_PARTS + 8) * sizeof(unsigned long long)));
    CU(h, h->jrow_of.reserve(h->tab.max_j * sizeof(uint32_t), st));
    CU(h, cudaMemsetAsync(h->d_export, 0, (2 * MAX_PARTS + 8) * sizeof(unsigned long long), st));
    h->mail.parity ^= 1; /* the slots of the exchange before this one may still be read by a slower owner */
    h->mail.epoch += 1;
    /* no host round trip here: whether the last batch's fold was complete (no table was full) is checked later, by
     * ipx_exchange_settle or ipx_flush */
    h->exported_unsettled = h->unsettled;
    {
        /* creation thresholds for anything still unresolved: keys that never got one are not exported */
        Span s(h, &h->tm_motif, 1);
        launch_resolve_T(h->tab, genome_view(h), st);
    }
    {
        /* only what other ranks own moves; this rank's own share stays where it is */
        Span s(h, &h->tm_export, 4);
        launch_mail_export(h->tab, h->d_name_rank, sp, h->mail, h->d_export, (uint32_t *)h->jrow_of.p, h->sm_count, st);
    }
    CU(h, cudaGetLastError());
    return IPX_OK;
}

namespace {
int enqueue_mail_import(ipx_handle *h);
}
extern "C" int ipx_exchange_settle(ipx_handle *h, int32_t *repeat_exchange) {
    if (!h || !repeat_exchange) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    *repeat_exchange = 0;
    bool repeated = false;
    int rc = settle_batch(h, &repeated);
    if (rc) return rc;
    rc = settle_import(h, &repeated);
    if (rc) return rc;
    if (h->late_fold) {
        *repeat_exchange = 1;
        h->late_fold = false;
        h->finalized = false; /* whatever was finalized saw a table that is about to change */
        h->finalize_settled = false;
    }
    return IPX_OK;
}

namespace {
int enqueue_mail_import(ipx_handle *h) {
    Span s(h, &h->tm_import, 3);
    unsigned long long *verdict = h->d_export + 2 * MAX_PARTS + 4;
    launch_mail_import(h->tab, h->mail, verdict, (uint32_t *)h->jid_of.p, h->sm_count, h->mail_timeout_ns, h->stream);
    CU(h, cudaMemcpyAsync(h->h_flags + 4, verdict, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaGetLastError());
    return IPX_OK;
}

/* the import checks on the device that everything fits before it folds anything; when it refused, grow and repeat */
int settle_import(ipx_handle *h, bool *repeated) {
    if (!h->import_pending) return IPX_OK;
    for (int round = 0; round < 40; round++) {
        CU(h, cudaStreamSynchronize(h->stream));
        if (h->h_flags[4] != 2ull) break;
        TableState ts;
        int rc = read_table_state(h, &ts);
        if (rc) return rc;
        if (ts.overflow & (OVF_MAIL | OVF_PEER)) break; /* not a matter of room: reported by ipx_flush */
        *repeated = true;
        launch_settle_counters(h->tab, h->stream);
        const uint64_t nj = h->h_flags[5], nh = h->h_flags[6];
        uint64_t in_use = std::min<uint64_t>(ts.n_keys, h->tab.max_j);
        while (in_use + nj > h->tab.max_j) {
            rc = grow_rows(h, in_use);
            if (rc) return rc;
        }
        while (2 * (in_use + nj) > h->tab.cap_mask + 1) {
            rc = rebuild_slots(h, true);
            if (rc) return rc;
        }
        const uint64_t log_in_use = std::min<uint64_t>(ts.n_olog, h->tab.olog_cap);
        if (log_in_use + nh > h->tab.olog_cap) {
            rc = grow_log(h, log_in_use, log_in_use + nh);
            if (rc) return rc;
        }
        h->log_bound = log_in_use + nh;
        rc = enqueue_mail_import(h);
        if (rc) return rc;
    }
    h->import_pending = false;
    return IPX_OK;
}
}  // namespace

extern "C" int ipx_partial_import_mail(ipx_handle *h) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (!h->mail_connected) return fail(h, IPX_ERR_STATE, "call ipx_mailbox_connect first");
    if (h->finalized) return fail(h, IPX_ERR_STATE, "table already finalized");
    int rc = alloc_tables(h);
    if (rc) return rc;
    if (!h->d_export) return fail(h, IPX_ERR_STATE, "call ipx_partial_export_mail first");
    CU(h, h->jid_of.reserve(((uint64_t)h->mail.world * h->mail.cap_j + 1) * sizeof(uint32_t), h->stream));
    rc = enqueue_mail_import(h);
    if (rc) return rc;
    h->import_pending = true;
    h->dirty = true;
    return IPX_OK;
}

extern "C" int ipx_sample_keys(ipx_handle *h, uint64_t stride, uint64_t max_keys, int32_t *tid, uint32_t *intron_start, uint64_t *n_keys) {
    if (!h || !n_keys || (max_keys && (!tid || !intron_start))) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    *n_keys = 0;
    if (!h->tables_ready || max_keys == 0) return IPX_OK;
    int rc = settle_batch(h);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    DeviceBuf d_tid, d_i0, d_n;
    CU(h, d_tid.reserve(max_keys * 4, st));
    CU(h, d_i0.reserve(max_keys * 4, st));
    CU(h, d_n.reserve(8, st));
    launch_sample_keys(h->tab, stride, max_keys, (int32_t *)d_tid.p, (uint32_t *)d_i0.p, (unsigned long long *)d_n.p, st);
    CU(h, cudaGetLastError());
    unsigned long long n = 0;
    CU(h, cudaMemcpyAsync(&n, d_n.p, 8, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    if (n) {
        CU(h, cudaMemcpyAsync(tid, d_tid.p, n * 4, cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(intron_start, d_i0.p, n * 4, cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
    }
    d_tid.release(); d_i0.release(); d_n.release();
    *n_keys = n;
    return IPX_OK;
}

/* ------------------------------------------------------------------ multi-GPU with the canon filter */
extern "C" int ipx_thresholds_rows(ipx_handle *h, uint64_t *n_rows, uint32_t *row_bytes) {
    if (!h || !n_rows) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (row_bytes) *row_bytes = (uint32_t)sizeof(ThresholdRow);
    *n_rows = 0;
    if (h->targets.empty()) return fail(h, IPX_ERR_STATE, "ipx_set_targets must be called first");
    int rc = alloc_tables(h);
    if (rc) return rc;
    TableState ts;
    rc = read_table_state(h, &ts);
    if (rc) return rc;
    if (ts.overflow & OVF_TABLE) {
        /* the first-seen pass met a full table: grow it and run the pass again over the kept events (it is idempotent) */
        rc = grow_and_repeat(h, ts, [&]() {
            for (auto &d : h->deferred)
                launch_fold(MODE_PASS1, h->tab, (Event *)d.events, (TileInfo *)d.tile_info, d.n_tiles, h->sm_count, h->stream);
            cudaError_t e = cudaGetLastError();
            return e == cudaSuccess ? IPX_OK : fail(h, IPX_ERR_CUDA, "CUDA error %s while repeating a fold", cudaGetErrorName(e));
        });
        if (rc) return rc;
        rc = read_table_state(h, &ts);
        if (rc) return rc;
    }
    {
        Span s(h, &h->tm_motif, 1);
        launch_resolve_T(h->tab, genome_view(h), h->stream);
    }
    *n_rows = std::min<uint64_t>(ts.n_keys, h->tab.max_j);
    return IPX_OK;
}

extern "C" int ipx_thresholds_export(ipx_handle *h, void *dev_rows) {
    if (!h || !dev_rows) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (!h->tables_ready) return fail(h, IPX_ERR_STATE, "call ipx_thresholds_rows first");
    unsigned long long nk = 0;
    CU(h, cudaMemcpyAsync(&nk, h->tab.n_keys, sizeof(nk), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    nk = std::min<uint64_t>(nk, h->tab.max_j);
    launch_export_thresholds(h->tab, (ThresholdRow *)dev_rows, nk, h->stream);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream)); /* the rows are complete when this returns: callers hand them to another stream */
    return IPX_OK;
}

extern "C" int ipx_thresholds_import(ipx_handle *h, const void *dev_rows, uint64_t n_rows) {
    if (!h || (n_rows && !dev_rows)) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    int rc = alloc_tables(h);
    if (rc) return rc;
    launch_import_thresholds(h->tab, (const ThresholdRow *)dev_rows, n_rows, h->stream);
    CU(h, cudaGetLastError());
    return IPX_OK;
}

extern "C" int ipx_count_deferred(ipx_handle *h) {
    if (!h) return IPX_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->finalized) return fail(h, IPX_ERR_STATE, "table already finalized");
    for (auto &d : h->deferred) {
        Span s(h, &h->tm_aggregate, 1);
        launch_fold(MODE_PASS2, h->tab, (Event *)d.events, (TileInfo *)d.tile_info, d.n_tiles, h->sm_count, h->stream);
    }
    CU(h, cudaGetLastError());
    drop_deferred(h);
    return IPX_OK;
}

/* ------------------------------------------------------------------ merge */
namespace {

std::string describe(const ipx_handle *h, const ipx_merge_rows *rows, int32_t chrom, uint32_t s, uint32_t e, char strand,
                     uint32_t site) {
    /* Junction::get_description (junctions.cc:72-74); mk_coords_str prints ints */
    char buf[512];
    snprintf(buf, sizeof(buf), "%s:%d-%d[%c/%s]", h->targets[chrom].c_str(), (int)s, (int)e, strand, rows->site_strings[site]);
    return buf;
}

/* bytes of one merge row over the link: chrom, start, end, anchors, sites, confidence (4 each), counts (12), strand (1) */
const uint64_t MERGE_ROW_BYTES = 41;
/* rows per upload chunk: 10.7 MB of pinned staging each (a tool pays for pinning them once per run: 0.45 ms per MB; with
 * 1 M-row chunks on 8 lanes that was 0.3 s) */
const uint64_t MERGE_CHUNK = 1ull << 18;

void merge_free_table(MergeWork &m) {
    cudaFree(m.tab.slots); cudaFree(m.tab.jkey); cudaFree(m.tab.jtid); cudaFree(m.tab.jslot); cudaFree(m.stat); cudaFree(m.slow_of_id);
    m.tab.slots = nullptr; m.tab.jkey = nullptr; m.tab.jtid = nullptr; m.tab.jslot = nullptr; m.stat = nullptr; m.slow_of_id = nullptr;
    m.tab.max_j = 0;
}

void merge_release(MergeWork &m) {
    merge_free_table(m);
    cudaFree(m.counters);
    m.counters = nullptr;
    for (auto &l : m.lanes) {
        if (l.stage[0]) cudaFreeHost(l.stage[0]);
        if (l.stage[1]) cudaFreeHost(l.stage[1]);
        if (l.free_ev[0]) cudaEventDestroy(l.free_ev[0]);
        if (l.free_ev[1]) cudaEventDestroy(l.free_ev[1]);
        if (l.st) cudaStreamDestroy(l.st);
    }
    m.lanes.clear();
    if (m.h_state) cudaFreeHost(m.h_state);
    m.h_state = nullptr;
}

/* the intron table of the merge fold: `max_j` intron rows, twice as many hash slots */
int merge_alloc_table(ipx_handle *h, uint64_t max_j) {
    MergeWork &m = h->mw;
    if (!m.counters) {
        CU(h, cudaMalloc(&m.counters, (128 + N_ID_POOLS) * sizeof(unsigned long long)));
        CU(h, cudaMemsetAsync(m.counters, 0, (128 + N_ID_POOLS) * sizeof(unsigned long long), h->stream));
        CU(h, cudaMallocHost(&m.h_state, 16 * sizeof(unsigned long long)));
    }
    if (m.tab.max_j >= max_j) return IPX_OK;
    merge_free_table(m);
    if (max_j >> 31) return fail(h, IPX_ERR_CAPACITY, "more than 2^31 introns");
    Tables &t = m.tab;
    CU(h, cudaMalloc(&t.slots, 2 * max_j * sizeof(Slot)));
    CU(h, cudaMalloc(&t.jkey, max_j * sizeof(unsigned long long)));
    CU(h, cudaMalloc(&t.jtid, max_j * sizeof(int32_t)));
    CU(h, cudaMalloc(&t.jslot, max_j * sizeof(uint32_t)));
    CU(h, cudaMalloc(&m.stat, max_j * sizeof(MStat)));
    CU(h, cudaMalloc(&m.slow_of_id, max_j));
    t.cap_mask = 2 * max_j - 1;
    t.max_j = max_j;
    t.n_keys = m.counters + 0;
    t.overflow = (uint32_t *)(m.counters + 16);
    t.id_pools = m.counters + 128;
    t.kl = make_key_layout(31, 0);
    t.kl.wide = 1; /* the chromosome is compared beside the (start, end) key */
    return IPX_OK;
}

/* Host rows -> device columns.  The rows are ordinary (pageable) memory: worker threads copy chunk after chunk into
 * pinned staging blocks and send them on streams of their own, so that staging, the link and the fold of the chunks
 * that have arrived overlap.  chunk_ready[c] is recorded behind chunk c's copies. */
struct MergeUpload {
    std::vector<cudaEvent_t> chunk_ready;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> issued; /* chunk c's copies are enqueued (its event may be waited for) */
    cudaError_t err = cudaSuccess;
    ~MergeUpload() {
        for (auto &w : workers) w.join();
        for (auto e : chunk_ready) cudaEventDestroy(e);
    }
};

struct MergeColumns {
    DeviceBuf chrom, i0, i1, a0, a1, strand, sites, counts, conf;
    cudaError_t reserve(uint64_t n, cudaStream_t st) {
        cudaError_t e;
        for (DeviceBuf *b : {&chrom, &i0, &i1, &a0, &a1, &sites, &conf})
            if ((e = b->reserve(n * 4, st)) != cudaSuccess) return e;
        if ((e = strand.reserve(n, st)) != cudaSuccess) return e;
        return counts.reserve(n * 12, st);
    }
    void release() {
        for (DeviceBuf *b : {&chrom, &i0, &i1, &a0, &a1, &strand, &sites, &counts, &conf}) b->release();
    }
    MergeRowsView view(uint64_t first, uint64_t n) const {
        return MergeRowsView{(const int32_t *)chrom.p + first, (const uint32_t *)i0.p + first, (const uint32_t *)i1.p + first,
                             (const uint32_t *)a0.p + first, (const uint32_t *)a1.p + first, (const char *)strand.p + first,
                             (const uint32_t *)sites.p + first, (const uint32_t *)counts.p + 3 * first, (const float *)conf.p + first, n, first};
    }
};

int merge_lanes(ipx_handle *h, int want) {
    MergeWork &m = h->mw;
    while ((int)m.lanes.size() < want) {
        MergeWork::Lane l{};
        CU(h, cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CU(h, cudaMallocHost(&l.stage[k], MERGE_CHUNK * MERGE_ROW_BYTES));
            CU(h, cudaEventCreateWithFlags(&l.free_ev[k], cudaEventDisableTiming));
        }
        m.lanes.push_back(l);
    }
    return IPX_OK;
}

void merge_upload_worker(ipx_handle *h, const ipx_merge_rows *rows, const MergeColumns *cols, MergeUpload *up, int lane, int n_lanes) {
    cudaSetDevice(h->device);
    MergeWork::Lane &l = h->mw.lanes[lane];
    const uint64_t n = rows->n_rows, n_chunks = (n + MERGE_CHUNK - 1) / MERGE_CHUNK;
    int turn = 0;
    for (uint64_t c = (uint64_t)lane; c < n_chunks; c += (uint64_t)n_lanes, turn ^= 1) {
        const uint64_t first = c * MERGE_CHUNK, len = std::min<uint64_t>(MERGE_CHUNK, n - first);
        char *stage = (char *)l.stage[turn];
        cudaError_t e = cudaEventSynchronize(l.free_ev[turn]); /* the copies that last read this block are done */
        struct Col { const void *src; void *dst; uint64_t width; };
        const Col col[9] = {{rows->chrom, cols->chrom.p, 4}, {rows->intron_start, cols->i0.p, 4}, {rows->intron_end, cols->i1.p, 4},
                            {rows->anchor_start, cols->a0.p, 4}, {rows->anchor_end, cols->a1.p, 4}, {rows->sites, cols->sites.p, 4},
                            {rows->confidence, cols->conf.p, 4}, {rows->counts, cols->counts.p, 12}, {rows->strand, cols->strand.p, 1}};
        uint64_t at = 0;
        for (const Col &k : col) {
            memcpy(stage + at, (const char *)k.src + first * k.width, len * k.width);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync((char *)k.dst + first * k.width, stage + at, len * k.width, cudaMemcpyHostToDevice, l.st);
            at += len * k.width;
        }
        if (e == cudaSuccess) e = cudaEventRecord(l.free_ev[turn], l.st);
        if (e == cudaSuccess) e = cudaEventRecord(up->chunk_ready[c], l.st);
        {
            std::lock_guard<std::mutex> g(up->mu);
            if (e != cudaSuccess && up->err == cudaSuccess) up->err = e;
            up->issued[c] = 1;
        }
        up->cv.notify_all();
    }
}

}  // namespace

extern "C" void ipx_free(void *p) { free(p); }

extern "C" int ipx_merge_rows_fold(ipx_handle *h, const ipx_merge_rows *rows, ipx_table **out, ipx_text *conflict_log) {
    return ipx_merge_rows_fold_part(h, rows, nullptr, out, conflict_log, nullptr, nullptr, nullptr);
}

extern "C" int ipx_merge_rows_fold_part(ipx_handle *h, const ipx_merge_rows *rows, const uint64_t *row_ids, ipx_table **out,
                                        ipx_text *conflict_log, uint64_t **conflict_rows, uint64_t *n_conflict_rows, uint64_t *fail_row) {
    if (!h || !rows || !out) return IPX_ERR_ARG;
    if (conflict_rows) *conflict_rows = nullptr;
    if (n_conflict_rows) *n_conflict_rows = 0;
    if (fail_row) *fail_row = ~0ull;
    std::vector<uint64_t> log_rows;
    auto give_log_rows = [&]() {
        if (!conflict_rows || !n_conflict_rows || log_rows.empty()) return;
        *conflict_rows = (uint64_t *)malloc(log_rows.size() * sizeof(uint64_t));
        if (!*conflict_rows) return;
        memcpy(*conflict_rows, log_rows.data(), log_rows.size() * sizeof(uint64_t));
        *n_conflict_rows = log_rows.size();
    };
    *out = nullptr;
    if (conflict_log) { conflict_log->data = nullptr; conflict_log->len = 0; }
    cudaSetDevice(h->device);
    if (h->targets.empty()) return fail(h, IPX_ERR_STATE, "ipx_set_targets must be called first");
    cudaStream_t st = h->stream;
    MergeWork &m = h->mw;
    const uint64_t n = rows->n_rows;
    Trace tr;
    /* The table is sized by introns, not rows: a guess from the row count (tables of the same genes share most of their
     * introns), grown when the fold runs out of rows.  A table kept from an earlier call is used as it is. */
    uint64_t want_j = 1ull << 16;
    while (want_j < n / 16 && want_j < (1ull << 24)) want_j <<= 1;
    if (const char *e = getenv("IPX_MERGE_ROWS")) want_j = pow2_at_least(strtoull(e, nullptr, 10));
    int rc = merge_alloc_table(h, want_j);
    if (rc) return rc;
    unsigned long long *n_unfolded = m.counters + 32, *n_slow = m.counters + 48, *d_count = m.counters + 64;

    MergeColumns cols;
    CU(h, cols.reserve(n ? n : 1, st));
    CU(h, cudaStreamSynchronize(st)); /* the columns exist before other streams write them */
    const uint64_t n_chunks = (n + MERGE_CHUNK - 1) / MERGE_CHUNK;
    {
        /* the spans time the kernels alone: a chunk's span starts after the wait for its upload */
        {
            Span s(h, &h->tm_merge, 3);
            launch_merge_init(m.tab, m.stat, n_unfolded, st);
        }
        if (n) {
            int n_lanes = (int)std::min<uint64_t>(n_chunks, std::max(1u, std::min(4u, std::thread::hardware_concurrency() / 2)));
            if (const char *e = getenv("IPX_MERGE_UPLOAD_THREADS")) n_lanes = std::max(1, std::min(atoi(e), 32));
            rc = merge_lanes(h, n_lanes);
            if (rc) return rc;
            MergeUpload up;
            up.chunk_ready.resize(n_chunks);
            up.issued.assign(n_chunks, 0);
            for (auto &e : up.chunk_ready) CU(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            for (int l = 0; l < n_lanes; l++) up.workers.emplace_back(merge_upload_worker, h, rows, &cols, &up, l, n_lanes);
            for (uint64_t c = 0; c < n_chunks; c++) {
                {
                    std::unique_lock<std::mutex> g(up.mu);
                    up.cv.wait(g, [&] { return up.issued[c] != 0; });
                }
                cudaStreamWaitEvent(st, up.chunk_ready[c], 0);
                Span s(h, &h->tm_merge, 1);
                launch_merge_fold(cols.view(c * MERGE_CHUNK, std::min<uint64_t>(MERGE_CHUNK, n - c * MERGE_CHUNK)), m.tab, m.stat, n_unfolded,
                                  h->sm_count, st);
            }
            tr.mark("merge: folds enqueued");
            for (auto &w : up.workers) w.join();
            up.workers.clear();
            tr.mark("merge: upload threads joined");
            if (up.err != cudaSuccess) {
                cols.release();
                return fail(h, IPX_ERR_CUDA, "CUDA error %s while uploading merge rows", cudaGetErrorName(up.err));
            }
            cudaStreamSynchronize(st); /* before the chunk events are destroyed */
        }
    }
    tr.mark("merge: rows uploaded and folded");
    /* did every row find room?  otherwise: twice the introns, fold everything again (the rows are on the device) */
    unsigned long long *hs = m.h_state;
    for (int round = 0;; round++) {
        CU(h, cudaMemcpyAsync(&hs[0], m.tab.n_keys, 8, cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(&hs[1], n_unfolded, 8, cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        if (hs[1] == 0) break;
        if (round > 20) { cols.release(); return fail(h, IPX_ERR_CAPACITY, "the intron table could not be grown far enough"); }
        rc = merge_alloc_table(h, m.tab.max_j * 2);
        if (rc) { cols.release(); return rc; }
        Span s(h, &h->tm_merge, 4);
        launch_merge_init(m.tab, m.stat, n_unfolded, st);
        launch_merge_fold(cols.view(0, n), m.tab, m.stat, n_unfolded, h->sm_count, st);
    }
    const uint64_t B = std::max<uint64_t>(1, std::min<uint64_t>(hs[0], m.tab.max_j)); /* intron ids handed out, holes included */

    /* key order (chromosome-name rank, start, end): the extraction path's sample sort */
    DeviceBuf k_key, k_sub, order, ss_split, ss_cnt, ss_bk, ss_bs, ss_bv;
    auto release_sort = [&]() { for (DeviceBuf *b : {&k_key, &k_sub, &order, &ss_split, &ss_cnt, &ss_bk, &ss_bs, &ss_bv}) b->release(); };
    CU(h, k_key.reserve(B * 8, st)); CU(h, k_sub.reserve(B * 4, st)); CU(h, order.reserve(B * 4, st));
    CU(h, ss_split.reserve(sample_sort_max_buckets() * 8, st)); CU(h, ss_cnt.reserve((sample_sort_max_buckets() + 1) * 4, st));
    uint64_t J = 0;
    for (uint64_t C = 1024;;) {
        const uint64_t cells = sample_sort_buckets(B) * C;
        CU(h, ss_bk.reserve(cells * 8, st)); CU(h, ss_bs.reserve(cells * 4, st)); CU(h, ss_bv.reserve(cells * 4, st));
        SampleSort ss{(unsigned long long *)ss_split.p, (uint32_t *)ss_cnt.p, (unsigned long long *)ss_bk.p, (uint32_t *)ss_bs.p,
                      (uint32_t *)ss_bv.p, C};
        {
            Span s(h, &h->tm_merge, 4);
            CU(h, cudaMemsetAsync(d_count, 0, 8 * sizeof(unsigned long long), st));
            launch_merge_keys(m.tab, h->d_name_rank, B, d_count, (unsigned long long *)k_key.p, (uint32_t *)k_sub.p, st);
            CU(h, sample_sort_enqueue((const unsigned long long *)k_key.p, (const uint32_t *)k_sub.p, d_count + 4, B, ss, (uint32_t *)order.p,
                                      nullptr, d_count, st));
        }
        CU(h, cudaMemcpyAsync(hs, d_count, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        if (!(hs[3] & 8ull)) { J = hs[0]; break; }
        C = std::max<uint64_t>(2 * C, hs[7] + hs[7] / 4); /* a bucket was larger than its cells */
    }
    tr.mark("merge: key order");
    const uint64_t Ja = J ? J : 1;
    HostTable *ht = new HostTable();
    memset(&ht->t, 0, sizeof(ht->t));
    std::vector<uint32_t> site;
    std::vector<uint8_t> slow;
    std::vector<unsigned long long> counts;
    {
        DeviceBuf o_chrom, o_i0, o_i1, o_a0, o_a1, o_site, o_counts, o_conf, o_strand, o_slow;
        cudaError_t e = cudaSuccess;
        auto chk = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
        chk(o_chrom.reserve(Ja * 4, st)); chk(o_i0.reserve(Ja * 4, st)); chk(o_i1.reserve(Ja * 4, st)); chk(o_a0.reserve(Ja * 4, st));
        chk(o_a1.reserve(Ja * 4, st)); chk(o_site.reserve(Ja * 4, st)); chk(o_counts.reserve(Ja * 24, st)); chk(o_conf.reserve(Ja * 4, st));
        chk(o_strand.reserve(Ja, st)); chk(o_slow.reserve(Ja, st));
        if (e == cudaSuccess) {
            Span s(h, &h->tm_merge, 1);
            MergeOut mo{(int32_t *)o_chrom.p, (uint32_t *)o_i0.p, (uint32_t *)o_i1.p, (uint32_t *)o_a0.p, (uint32_t *)o_a1.p,
                        (uint32_t *)o_site.p, (unsigned long long *)o_counts.p, (float *)o_conf.p, (char *)o_strand.p,
                        (uint8_t *)o_slow.p};
            chk(cudaMemsetAsync(m.slow_of_id, 0, m.tab.max_j, st));
            launch_merge_gather(m.tab, m.stat, (const uint32_t *)order.p, J, mo, m.slow_of_id, st);
        }
        chk(d2h(ht->chrom, o_chrom.p, J, st)); chk(d2h(ht->i0, o_i0.p, J, st)); chk(d2h(ht->i1, o_i1.p, J, st));
        chk(d2h(ht->a0, o_a0.p, J, st)); chk(d2h(ht->a1, o_a1.p, J, st)); chk(d2h(site, o_site.p, J, st));
        chk(d2h(counts, o_counts.p, J * 3, st)); chk(d2h(ht->conf, o_conf.p, J, st)); chk(d2h(ht->strand, o_strand.p, J, st));
        chk(d2h(slow, o_slow.p, J, st));
        chk(cudaStreamSynchronize(st));
        chk(cudaGetLastError());
        for (DeviceBuf *b : {&o_chrom, &o_i0, &o_i1, &o_a0, &o_a1, &o_site, &o_counts, &o_conf, &o_strand, &o_slow}) b->release();
        release_sort();
        if (e != cudaSuccess) {
            delete ht;
            cols.release();
            return fail(h, IPX_ERR_CUDA, "CUDA error %s in merge: %s", cudaGetErrorName(e), cudaGetErrorString(e));
        }
    }
    ht->counts.assign(counts.begin(), counts.end());
    tr.mark("merge: table to host");

    /* Introns whose fold depends on row order: the device picks their rows out of the input, the host folds those in
     * row order exactly as Junction::merge does (junctions.cc:76-105). */
    std::string log;
    std::unordered_map<uint64_t, std::vector<uint64_t>> slow_by_key; /* (start << 32 | end) -> table rows */
    uint64_t n_slow_introns = 0;
    for (uint64_t i = 0; i < J; i++)
        if (slow[i]) {
            slow_by_key[((uint64_t)ht->i0[i] << 32) | ht->i1[i]].push_back(i);
            n_slow_introns++;
        }
    std::vector<unsigned long long> slow_rows;
    if (n_slow_introns) {
        DeviceBuf d_rows;
        for (uint64_t cap = std::min<uint64_t>(n, 1ull << 16);;) {
            cudaError_t e = d_rows.reserve(cap * 8, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(n_slow, 0, 8, st);
            if (e == cudaSuccess) {
                Span s(h, &h->tm_merge, 1);
                launch_merge_slow_rows(cols.view(0, n), m.tab, m.slow_of_id, n_slow, cap, (unsigned long long *)d_rows.p, h->sm_count, st);
                e = cudaMemcpyAsync(&hs[0], n_slow, 8, cudaMemcpyDeviceToHost, st);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e == cudaSuccess && hs[0] <= cap) {
                e = d2h(slow_rows, d_rows.p, hs[0], st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            }
            if (e != cudaSuccess) {
                d_rows.release(); cols.release();
                delete ht;
                return fail(h, IPX_ERR_CUDA, "CUDA error %s in merge: %s", cudaGetErrorName(e), cudaGetErrorString(e));
            }
            if (hs[0] <= cap) break;
            cap = hs[0];
        }
        d_rows.release();
        std::sort(slow_rows.begin(), slow_rows.end());
    }
    cols.release();
    if (n_slow_introns) {
        std::vector<char> started(J, 0);
        for (uint64_t r : slow_rows) {
            auto it = slow_by_key.find(((uint64_t)rows->intron_start[r] << 32) | rows->intron_end[r]);
            if (it == slow_by_key.end()) continue;
            uint64_t ti = ~0ull;
            for (uint64_t cand : it->second)
                if (ht->chrom[cand] == rows->chrom[r]) ti = cand;
            if (ti == ~0ull) continue;
            const char rstrand = rows->strand[r];
            const uint32_t rsite = rows->sites[r];
            if (!started[ti]) {
                started[ti] = 1;
                ht->a0[ti] = rows->anchor_start[r];
                ht->a1[ti] = rows->anchor_end[r];
                ht->strand[ti] = rstrand;
                site[ti] = rsite;
                for (int c = 0; c < 3; c++) ht->counts[3 * ti + c] = rows->counts[3 * r + c];
                ht->conf[ti] = rows->confidence[r];
                continue;
            }
            bool update_strand = false;
            if (rstrand != '.' && ht->strand[ti] == '.') {
                update_strand = true;
            } else if (rstrand == '.' && ht->strand[ti] != '.') {
            } else if (rstrand != ht->strand[ti]) {
                log += "intron records have different strand:" +
                       describe(h, rows, rows->chrom[r], rows->intron_start[r], rows->intron_end[r], rstrand, rsite) + " with " +
                       describe(h, rows, ht->chrom[ti], ht->i0[ti], ht->i1[ti], ht->strand[ti], site[ti]) + "\n";
                log_rows.push_back(row_ids ? row_ids[r] : r);
                continue;
            } else if (strcmp(rows->site_strings[rsite], rows->site_strings[site[ti]]) != 0) {
                std::string msg = "intron records have different splice sites:" +
                                  describe(h, rows, rows->chrom[r], rows->intron_start[r], rows->intron_end[r], rstrand, rsite) +
                                  " with " + describe(h, rows, ht->chrom[ti], ht->i0[ti], ht->i1[ti], ht->strand[ti], site[ti]);
                if (conflict_log) give_text(log, conflict_log);
                give_log_rows();
                if (fail_row) *fail_row = row_ids ? row_ids[r] : r;
                delete ht;
                return fail(h, IPX_ERR_MERGE, "%s", msg.c_str());
            }
            ht->a0[ti] = std::min(rows->anchor_start[r], ht->a0[ti]);
            ht->a1[ti] = std::max(rows->anchor_end[r], ht->a1[ti]);
            for (int c = 0; c < 3; c++) ht->counts[3 * ti + c] += rows->counts[3 * r + c];
            ht->conf[ti] = std::max(rows->confidence[r], ht->conf[ti]);
            if (update_strand) {
                ht->strand[ti] = rstrand;
                site[ti] = rsite;
            }
        }
    }
    tr.mark("merge: host fold of flagged introns");
    ht->sites.assign(J * 8, 0);
    for (uint64_t i = 0; i < J; i++) snprintf(ht->sites.data() + 8 * i, 8, "%s", rows->site_strings[site[i]]);
    ht->site_text.reserve(rows->n_site_strings);
    for (uint32_t k = 0; k < rows->n_site_strings; k++) ht->site_text.emplace_back(rows->site_strings[k]);
    ht->site_ids.assign(site.begin(), site.end());
    ht->ijunc.assign(J, 0);
    ht->hist_off.assign(J + 1, 0);
    ht->bind();
    /* sort_by_anchors + renumber_junctions (intronProspectorMerge.cc:201-223): same std::sort, same start order */
    std::vector<const char *> names(h->targets.size());
    for (size_t i = 0; i < names.size(); i++) names[i] = h->targets[i].c_str();
    std::vector<uint64_t> idx;
    sort_rows(&ht->t, names.data(), true, 0, J, idx);
    for (uint64_t k = 0; k < J; k++) ht->ijunc[idx[k]] = (uint32_t)k;
    tr.mark("merge: presentation order");
    h->totals.merge_rows += n;
    h->totals.junctions += J;
    if (conflict_log) give_text(log, conflict_log);
    give_log_rows();
    *out = &ht->t;
    return IPX_OK;
}
