/* ipx_table.cuh -- device-resident junction table: the replacement for the reference's
 * std::map<JunctionKey, Junction*> (junctions.hh:225-246) and Junction::read_offsets (junctions.hh:120).
 *
 * Layout in HBM (for real data a few tens of MB are live, i.e. mostly resident in the 126 MB L2):
 *   hash slots     slots[cap] (open addressing, linear probing): 32 bytes = one sector: key, the junction id of the key
 *                  and the first offsets of the junction's extra dense rows, so that one load answers a probe
 *   junction rows  indexed by junction id = order of first insertion, so everything a junction owns is COMPACT:
 *                  jkey / jtid / jslot (identity), JStat (64-byte record: per-strand-class first sequence numbers, the
 *                  creation threshold T, read counts by category, min anchor start, max anchor end), a dense row of
 *                  DENSE_W read-offset counters, and the number of its entries in the read-offset log.
 *                  clear / finalize / export walk [0, n_keys) with coalesced accesses and never touch the hash slots'
 *                  capacity; growing the hash (rehash) moves no junction data, growing the rows is a prefix copy.
 *   read-offset log  (junction id, uint16 read offset, count) entries for the offsets the dense rows do not take.
 *
 * key64 = target index (low bits that fit) | intron_start (31 bits) | intron length (L bits, L = bits(max_intron_length));
 * junction_qc guarantees length <= max_intron_length (junctions_extractor.cc:392-393) and BAM positions are int32, so
 * start and length are exact.  When the target index does not fit the bits that are left ("wide" layout: many contigs
 * or a large --max-intron-length) the slot's identity is (key64, full target index): a slot whose key64 matches is
 * compared by jtid[jid] as well.  The top bit is always 0, so ~0 is free as the empty marker.
 *
 * Order dependence of the reference is carried by "sequence values":
 *   sv = record stream index (36 bits) | candidate ordinal within the read (20 bits) | strand char (8 bits)
 * unsigned comparison of sv is stream order; atomicMin over sv finds the first event and brings its
 * strand along.  Three minima are kept per junction, one per strand class, because whether an event
 * may create the junction (canon filter) depends only on (key, strand class).
 */
#ifndef IPX_TABLE_CUH
#define IPX_TABLE_CUH
#include <stdint.h>
#include "ipx_walk.cuh"

namespace ipx {

static const uint64_t EMPTY64 = ~0ull;
static const uint32_t JID_UNSET = 0xFFFFFFFFu; /* the id field of an empty slot */
static const uint32_t JID_FULL = 0xFFFFFFFEu;  /* slot claimed when no junction row was left */
static const uint32_t JID_NONE = 0xFFFFFFFDu;  /* probe result: absent (lookup) or no room (insert) */
/* Read offsets (intron_start - read start) of short reads are < read length; counts for offsets below DENSE_W
 * live in a dense row of DENSE_W counters per junction (1 KB), where counting a read is a single fire-and-forget
 * atomic and the ordered run-length output needs no sort.  Larger offsets (second junction of a read, long reads)
 * go to the read-offset log, as does everything when the rows are disabled. */
static const uint32_t DENSE_W = 256;
static const uint32_t LOG_STAGE = 128; /* log entries a warp stages in shared memory between reservations */
/* Junction ids are handed out in blocks of ID_BLOCK to pools, one per thread block of an inserting kernel (kept in
 * global memory between launches): taking an id is a shared-memory atomic, and the table's one counter sees a block
 * per ID_BLOCK junctions instead of every insertion (a single address serves ~0.5 G atomics/s: at one per junction it
 * was the bottleneck of a cold fold).  Ids of a pool's unfinished block are "holes": rows nobody ever wrote, which
 * everything that walks the rows recognises by their reset state (T == EMPTY64, jslot == SLOT_NONE). */
static const uint32_t ID_BLOCK = 64;
static const uint32_t N_ID_POOLS = 1024; /* inserting kernels run at most this many thread blocks */
static const uint32_t SLOT_NONE = 0xFFFFFFFFu;
static const uint32_t BASE_UNSET = 0xFFFFFFFFu;
/* A junction owns DENSE_ROWS dense rows of DENSE_W counters.  Row 0 counts the offsets [0, DENSE_W): first junctions of
 * short reads (offset = length of the read's first aligned run).  The offsets of a read's later junctions include the
 * introns before them; the reads through one chain of junctions differ in their start by less than a read length, so
 * those offsets form a cluster narrower than a row per chain.  Rows 1 .. DENSE_ROWS - 1 are given out to such clusters
 * as they show up: the first offset no row takes yet centres the next free row on itself (slot.base[r - 1]), and from
 * then on the row takes [base, base + DENSE_W) less what earlier rows take.  What no row takes goes to the log. */
static const uint32_t DENSE_ROWS = 4;
static const uint32_t DENSE_LEAD = DENSE_W / 2; /* default of Tables::dense_lead */
static const int SV_ORD_BITS = 20;
static const int SV_SEQ_SHIFT = 28;

/* strand class of a read strand character: what get_splice_sites/correct_strand distinguish
 * (junctions_extractor.cc:421,434): '-' is reverse-complemented, '.' is corrected, anything else is kept */
enum : int { CLS_PLUS = 0, CLS_MINUS = 1, CLS_DOT = 2 };

IPX_HD int strand_class(char s) {
    return s == '-' ? CLS_MINUS : (s == '.' ? CLS_DOT : CLS_PLUS);
}

IPX_HD uint64_t make_sv(uint64_t record_seq, uint32_t ordinal, char strand) {
    return (record_seq << SV_SEQ_SHIFT) | ((uint64_t)(ordinal & ((1u << SV_ORD_BITS) - 1u)) << 8) | (uint8_t)strand;
}

/* Deferred fatal errors share the sv ordering so that the first one in stream order wins:
 *   CIGAR  error: after every candidate of its record (ordinal field all ones), low byte = op code (< 0x80)
 *   genome error: at the candidate that first needs the sequence, low byte = 0x80 | motif status */
IPX_HD unsigned long long make_err_cigar(uint64_t record_seq, uint32_t op) {
    return (record_seq << SV_SEQ_SHIFT) | (0xFFFFFull << 8) | (op & 0x7Fu);
}
IPX_HD unsigned long long make_err_genome(uint64_t first_sv, int status) {
    return (first_sv & ~0xFFull) | 0x80ull | (unsigned)(status & 0x7F);
}

#if defined(__CUDACC__)
struct __align__(64) JStat {
#else
struct alignas(64) JStat {
#endif
    unsigned long long first[3]; /* min sv per strand class, EMPTY64 = none */
    unsigned long long T;        /* sv of the event that creates the junction, EMPTY64 = not (yet) created */
    unsigned long long cnt[3];   /* read_counts[] (junctions.hh:117) */
    uint32_t a0;                 /* min anchor_start, 0xFFFFFFFF = none */
    uint32_t a1;                 /* max anchor_end */
};

/* one junction event between K1 (extract) and K2 (fold): 32 bytes */
#if defined(__CUDACC__)
struct __align__(16) Event {
#else
struct alignas(16) Event {
#endif
    unsigned long long key;
    unsigned long long sv;
    uint32_t a0, a1;  /* anchor_start, anchor_end */
    uint32_t off_cat; /* uint16 read offset | read category << 16 | strand class << 18 */
    int32_t tid;      /* target index (the key holds only its low bits in the wide layout) */
};

struct KeyLayout {
    int len_bits;    /* L */
    int start_shift; /* = L */
    int tid_shift;   /* = L + 31 */
    int tid_bits;    /* 63 - tid_shift: bits of the target index the key holds */
    uint32_t tid_mask; /* (1 << tid_bits) - 1 */
    int wide;        /* the target index does not fit tid_bits: compare jtid as well */
};

IPX_HD KeyLayout make_key_layout(int len_bits, uint64_t n_targets) {
    KeyLayout kl;
    kl.len_bits = len_bits;
    kl.start_shift = len_bits;
    kl.tid_shift = len_bits + 31;
    kl.tid_bits = 63 - kl.tid_shift;
    if (kl.tid_bits > 31) kl.tid_bits = 31;
    if (kl.tid_bits < 0) kl.tid_bits = 0;
    kl.tid_mask = kl.tid_bits ? ((1u << kl.tid_bits) - 1u) : 0u;
    kl.wide = (n_targets > (uint64_t)kl.tid_mask + 1ull) ? 1 : 0;
    return kl;
}

IPX_HD uint64_t make_key(const KeyLayout &kl, int32_t tid, int64_t i0, int64_t i1) {
    return ((uint64_t)((uint32_t)tid & kl.tid_mask) << kl.tid_shift) | ((uint64_t)i0 << kl.start_shift) | (uint64_t)(i1 - i0);
}
IPX_HD uint32_t key_start(const KeyLayout &kl, uint64_t k) { return (uint32_t)((k >> kl.start_shift) & 0x7FFFFFFFu); }
IPX_HD uint32_t key_len(const KeyLayout &kl, uint64_t k) { return (uint32_t)(k & ((1ull << kl.len_bits) - 1ull)); }

IPX_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ULL;
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ULL;
    x ^= x >> 32;
    return x;
}
/* hash of a junction's identity; the target index enters separately because the wide layout keeps only its low bits
 * in the key */
IPX_HD uint64_t junction_hash(uint64_t key, int32_t tid) { return mix64(key ^ ((uint64_t)(uint32_t)tid * 0x9E3779B97F4A7C15ull)); }

/* device-side view of the tables, passed to kernels by value */
/* one hash slot: a 32-byte sector */
#if defined(__CUDACC__)
struct __align__(32) Slot {
#else
struct alignas(32) Slot {
#endif
    unsigned long long key;         /* key64 or EMPTY64 */
    uint32_t jid;                   /* junction id of the key (JID_UNSET while its inserter publishes it) */
    uint32_t base[DENSE_ROWS - 1];  /* first offset of dense rows 1 .., BASE_UNSET = row not given out */
    uint32_t pad[2];
};

struct Tables {
    Slot *slots;              /* [cap] */
    uint64_t cap_mask;
    /* junction rows, [max_j] each, in order of first insertion */
    unsigned long long *jkey;
    int32_t *jtid;
    uint32_t *jslot;          /* hash slot of the junction's key (for clear) */
    JStat *stat;
    uint32_t *dense;          /* [max_j][DENSE_ROWS][DENSE_W] read counts per offset (see DENSE_ROWS), or NULL when disabled */
    uint32_t dense_lead;      /* a row given out to offset x takes [x - dense_lead, x - dense_lead + DENSE_W) */
    int32_t extra_rows;       /* rows 1 .. are given out (off for long reads, whose offsets scatter: everything >= DENSE_W
                                 goes to the log then) */
    uint32_t *olog_cnt;       /* entries of the read-offset log that name the junction */
    unsigned long long *id_pools; /* [N_ID_POOLS] first id of the pool's block << 32 | ids taken from it */
    uint64_t max_j;
    unsigned long long *olog; /* [olog_cap] junction id << 32 | read offset << 16 | count */
    uint64_t olog_cap;
    /* counters / flags, one cache line each to avoid false sharing */
    unsigned long long *n_keys;    /* junction ids handed out (may run past max_j when the rows are full) */
    unsigned long long *n_olog;    /* log entries reserved (may run past olog_cap when the log is full) */
    unsigned long long *n_events;  /* events that passed junction_qc */
    unsigned long long *n_live;    /* CIGAR words of live records */
    unsigned long long *err_sv;    /* min over failing records of (record seq << 28 | code); EMPTY64 = none */
    unsigned long long *next_tile;  /* K2: the next tile to hand out (reset by its launcher) */
    unsigned long long *n_unfolded; /* events K2 left in the event buffer because a table was full (see k_fold) */
    uint32_t *overflow;            /* bit set: 1 hash slots / junction rows full, 2 log full, 4 intron start beyond 2^31,
                                      8 mailbox slot too small, 16 a peer never published its export */
    KeyLayout kl;
};

enum : uint32_t { OVF_TABLE = 1u, OVF_LOG = 2u, OVF_START = 4u, OVF_MAIL = 8u, OVF_PEER = 16u };

#if defined(__CUDACC__)
static const uint32_t MAX_PROBE = 1u << 16;

/* number of junction ids in use (the counter keeps counting when the rows are full) */
__device__ __forceinline__ uint64_t junctions_in_use(const Tables &t) {
    const unsigned long long n = *t.n_keys;
    return n < t.max_j ? n : t.max_j;
}
__device__ __forceinline__ uint64_t log_entries_in_use(const Tables &t) {
    const unsigned long long n = *t.n_olog;
    return n < t.olog_cap ? n : t.olog_cap;
}

/* A thread block's pool of junction ids in shared memory (loaded from / stored back to t.id_pools by the kernel). */
struct IdPool {
    unsigned long long word; /* first id of the block << 32 | ids taken (>= ID_BLOCK: used up) */
    uint32_t lock;
};

__device__ __forceinline__ void id_pool_load(const Tables &t, IdPool *p) {
    p->word = t.id_pools[blockIdx.x % N_ID_POOLS];
    p->lock = 0u;
}
__device__ __forceinline__ void id_pool_store(const Tables &t, const IdPool *p) {
    t.id_pools[blockIdx.x % N_ID_POOLS] = *((volatile const unsigned long long *)&p->word);
}

/* the next junction id, or an id >= max_j when no row is left */
__device__ __forceinline__ uint32_t id_pool_take(const Tables &t, IdPool *p) {
    for (;;) {
        const unsigned long long old = atomicAdd(&p->word, 1ull);
        const uint32_t used = (uint32_t)old;
        if (used < ID_BLOCK) return (uint32_t)(old >> 32) + used;
        if (atomicCAS(&p->lock, 0u, 1u) == 0u) {
            /* this thread refills the pool unless somebody did in the meantime */
            bool full = false;
            if ((uint32_t)(*((volatile unsigned long long *)&p->word)) >= ID_BLOCK) {
                const unsigned long long nb = atomicAdd(t.n_keys, (unsigned long long)ID_BLOCK);
                if (nb + ID_BLOCK <= t.max_j) atomicExch(&p->word, nb << 32); else full = true;
            }
            __threadfence_block();
            atomicExch(&p->lock, 0u);
            if (full) return (uint32_t)t.max_j; /* the rows are used up (max_j < 2^31) */
        }
    }
}

/* what a probe learns in its one load of the slot */
struct Found {
    uint32_t jid;  /* JID_NONE: absent (lookup) or no room (insert) */
    uint32_t slot;
    uint32_t base[DENSE_ROWS - 1];
};

__device__ __forceinline__ void load_slot(const Slot *s, unsigned long long &key, uint32_t &jid, uint32_t base[DENSE_ROWS - 1]) {
    uint4 a, b;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(s));
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"((const char *)s + 16));
    key = ((unsigned long long)a.y << 32) | a.x;
    jid = a.z;
    base[0] = a.w;
    base[1] = b.x;
    base[2] = b.y;
}

/* 128-bit compare-and-swap on the first half of a slot -- (key, jid, base[0]) from the reset state to (key, jid,
 * BASE_UNSET) -- so that a slot's key and its junction id become visible together: nobody ever waits for another
 * thread to publish an id (threads that wait for each other's insertions across warps can wait in a circle).
 * Returns what the half held before. */
__device__ __forceinline__ bool slot_claim(Slot *s, unsigned long long key, uint32_t jid, uint32_t base0, unsigned long long &seen_key,
                                           uint32_t &seen_jid, uint32_t &seen_base0) {
    const unsigned long long c_lo = EMPTY64, c_hi = ((unsigned long long)BASE_UNSET << 32) | JID_UNSET;
    const unsigned long long n_hi = ((unsigned long long)base0 << 32) | jid;
    unsigned long long o_lo, o_hi;
    asm volatile(
        "{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
        "atom.relaxed.gpu.global.cas.b128 o, [%6], c, n;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(o_lo), "=l"(o_hi)
        : "l"(c_lo), "l"(c_hi), "l"(key), "l"(n_hi), "l"(s)
        : "memory");
    seen_key = o_lo;
    seen_jid = (uint32_t)o_hi;
    seen_base0 = (uint32_t)(o_hi >> 32);
    return o_lo == c_lo && o_hi == c_hi;
}

/* Find the junction of (key, tid); INSERT: claim a slot and a junction row when absent.  jid == JID_NONE when the key is
 * absent (lookup) or when no slot / row is left (insert: the overflow flag is raised, the caller keeps the event).
 * The junction rows of ids never handed out hold the reset values (init / clear), so an inserter has nothing to
 * initialise but the identity.
 * `spare` (INSERT; one per thread, JID_NONE at first, kept across calls): a junction id the thread took for an insertion
 * that another thread then won.  The id is taken BEFORE the slot is claimed and its identity (jkey, jtid) written and
 * fenced, so the claim publishes a complete junction in one atomic operation; a spare left over at the end of a kernel
 * is a hole like the unused ids of a pool's block. */
template <bool INSERT>
__device__ __forceinline__ Found junction_probe(const Tables &t, IdPool *pool, unsigned long long key, int32_t tid, uint32_t &spare) {
    Found f;
    f.jid = JID_NONE;
    f.slot = 0u;
    f.base[0] = f.base[1] = f.base[2] = BASE_UNSET;
    uint64_t h = junction_hash(key, tid) & t.cap_mask;
    for (uint32_t probe = 0; probe < MAX_PROBE; ++probe) {
        Slot *s = &t.slots[h];
        unsigned long long k;
        uint32_t j;
        load_slot(s, k, j, f.base);
        f.slot = (uint32_t)h;
        if (k == EMPTY64) {
            if (!INSERT) return f;
            if (spare == JID_NONE) {
                spare = id_pool_take(t, pool);
                if (spare >= t.max_j) spare = JID_NONE; /* no row left */
            }
            if (spare == JID_NONE) {
                /* The key is marked as one that met the full table, so that nobody with an id to spare inserts it later
                 * in the same launch: a key is in the table for all of its events of a pass or for none (the two-pass
                 * fold relies on that).  The slot is dropped by the rehash that follows the growth. */
                atomicOr(t.overflow, OVF_TABLE);
                if (slot_claim(s, key, JID_FULL, (uint32_t)tid, k, j, f.base[0])) return f; /* (the target index in place of a row base) */
            } else {
                t.jkey[spare] = key;
                t.jtid[spare] = tid;
                /* Whoever meets the id during the same launch reads jtid, and only in the wide layout: there the identity
                 * must be visible before the id is.  (A fence per insertion was a quarter of the fold's stall samples.) */
                if (t.kl.wide) __threadfence();
                if (slot_claim(s, key, spare, BASE_UNSET, k, j, f.base[0])) {
                    t.jslot[spare] = (uint32_t)h;
                    f.jid = spare;
                    spare = JID_NONE;
                    f.base[0] = f.base[1] = f.base[2] = BASE_UNSET;
                    return f;
                }
            }
            /* somebody else claimed the slot in the meantime: k / j / base[0] are what it holds now; a junction that
             * young has no further rows */
            f.base[1] = f.base[2] = BASE_UNSET;
        }
        if (k == key) {
            /* (key and id are written by one 128-bit operation; should a 16-byte load ever see them apart, the next
             * one does not -- this waits for nobody) */
            while (j == JID_UNSET) load_slot(s, k, j, f.base);
            if (j == JID_FULL) {
                if (INSERT && f.base[0] == (uint32_t)tid) return f; /* the same key met the full table before (flag already raised) */
            } else if (!t.kl.wide || *((volatile int32_t *)&t.jtid[j]) == tid) {
                f.jid = j;
                return f;
            }
        }
        h = (h + 1) & t.cap_mask;
    }
    if (INSERT) atomicOr(t.overflow, OVF_TABLE);
    return f;
}
/* lookup only */
__device__ __forceinline__ Found junction_find(const Tables &t, unsigned long long key, int32_t tid) {
    uint32_t none = JID_NONE;
    return junction_probe<false>(t, nullptr, key, tid, none);
}

/* Which dense counter takes `offset` for the junction in `slot`: row * DENSE_W + index, or DENSE_NONE (-> the log).
 * base[] is the caller's copy of the slot's row bases; a row given out here is written back to it. */
static const uint32_t DENSE_NONE = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t dense_place(const Tables &t, uint32_t slot, uint32_t base[DENSE_ROWS - 1], uint32_t offset) {
    if (t.dense == nullptr) return DENSE_NONE;
    if (offset < DENSE_W) return offset;
    if (!t.extra_rows) return DENSE_NONE;
#pragma unroll
    for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r) {
        if (base[r] == BASE_UNSET) {
            /* the next free row: centre it on this offset, unless somebody was faster */
            const uint32_t mine = offset - t.dense_lead;
            const uint32_t old = atomicCAS(&t.slots[slot].base[r], BASE_UNSET, mine);
            base[r] = old == BASE_UNSET ? mine : old;
        }
        if (offset >= base[r] && offset - base[r] < DENSE_W) return (r + 1u) * DENSE_W + (offset - base[r]);
    }
    return DENSE_NONE;
}
/* the same test against bases that are known to be complete (no row is given out) */
__device__ __forceinline__ uint32_t dense_find(const uint32_t base[DENSE_ROWS - 1], uint32_t offset) {
    if (offset < DENSE_W) return offset;
#pragma unroll
    for (uint32_t r = 0; r < DENSE_ROWS - 1; ++r)
        if (base[r] != BASE_UNSET && offset >= base[r] && offset - base[r] < DENSE_W) return (r + 1u) * DENSE_W + (offset - base[r]);
    return DENSE_NONE;
}

/* Per-warp writer of the read-offset log: entries are staged in shared memory (LOG_STAGE of them) and written out
 * with one reservation on the log's counter per stage, coalesced.  Every lane of the (converged) warp calls
 * log_append / log_flush together; `fill` is warp-uniform. */
struct LogStage {
    unsigned long long *buf; /* this warp's LOG_STAGE entries of shared memory */
    uint32_t fill;
};

__device__ __forceinline__ unsigned long long olog_entry(uint32_t jid, uint32_t offset16, uint32_t n) {
    return ((unsigned long long)jid << 32) | ((unsigned long long)(offset16 & 0xFFFFu) << 16) | (n & 0xFFFFu);
}

__device__ __forceinline__ void log_flush(const Tables &t, LogStage &s) {
    if (s.fill == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    __syncwarp();
    unsigned long long base = 0ull;
    if (lane == 0u) base = atomicAdd(t.n_olog, (unsigned long long)s.fill);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base + s.fill > t.olog_cap) {
        if (lane == 0u) atomicOr(t.overflow, OVF_LOG); /* the host sizes the log before every batch: not expected */
    } else {
        for (uint32_t i = lane; i < s.fill; i += 32u) {
            const unsigned long long e = s.buf[i];
            t.olog[base + i] = e;
            atomicAdd(&t.olog_cnt[(uint32_t)(e >> 32)], 1u);
        }
    }
    __syncwarp();
    s.fill = 0u;
}

__device__ __forceinline__ void log_append(const Tables &t, LogStage &s, bool want, unsigned long long entry) {
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, want);
    if (mask == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t m = (uint32_t)__popc(mask);
    if (s.fill + m > LOG_STAGE) log_flush(t, s);
    if (want) s.buf[s.fill + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = entry;
    s.fill += m;
}

/* min/max only ever tighten, so a stale read can only cause a redundant atomic, never a missed one */
__device__ __forceinline__ void stat_anchor(JStat *s, uint32_t a0, uint32_t a1) {
    if (a0 < *((volatile uint32_t *)&s->a0)) atomicMin(&s->a0, a0);
    if (a1 > *((volatile uint32_t *)&s->a1)) atomicMax(&s->a1, a1);
}
__device__ __forceinline__ void stat_first(JStat *s, int cls, uint64_t sv) {
    if (sv < *((volatile unsigned long long *)&s->first[cls])) atomicMin(&s->first[cls], (unsigned long long)sv);
}
#endif

}  // namespace ipx
#endif
