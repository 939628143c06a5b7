/* ipx_table.cuh -- device-resident junction table: the replacement for the reference's
 * std::map<JunctionKey, Junction*> (junctions.hh:225-246) and Junction::read_offsets (junctions.hh:120).
 *
 * Two open-addressing hash tables live in HBM for the whole run (and, being a few tens of MB for
 * real data, mostly in the 126 MB L2):
 *   junction table   key64 -> JStat (64-byte record: per-strand-class first sequence numbers, the
 *                    creation threshold T, read counts by category, min anchor start, max anchor end)
 *   offset table     (junction slot, uint16 read offset) -> count, the run-length form of read_offsets
 *
 * key64 = target index | intron_start (31 bits) | intron length (L bits, L = bits(max_intron_length));
 * junction_qc guarantees length <= max_intron_length (junctions_extractor.cc:392-393) and BAM positions
 * are int32, so the key is exact.  The top bit is always 0, so ~0 is free as the empty marker.
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
static const uint64_t TOMB64 = ~0ull - 1; /* offset-table entry folded into a dense row during finalize */
/* Read offsets (intron_start - read start) of short reads are < read length; counts for offsets below DENSE_W
 * live in a per-junction dense row (one 1 KB row per junction, allocated in insertion order), where counting a
 * read is a single fire-and-forget atomic and the ordered run-length output needs no sort.  Larger offsets
 * (long reads) and junctions beyond the pool go to the (junction, offset) hash table. */
static const uint32_t DENSE_W = 256;
static const uint32_t DENSE_NONE = 0xFFFFFFFFu;
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

/* offset-table entry: key and count share one 16-byte slot so that a probe and its count touch one sector */
#if defined(__CUDACC__)
struct __align__(16) HEntry {
#else
struct alignas(16) HEntry {
#endif
    unsigned long long key; /* junction slot << 16 | uint16 read offset, EMPTY64 = free */
    uint32_t cnt;
    uint32_t pad;
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
    uint32_t pad;
};

struct KeyLayout {
    int len_bits;   /* L */
    int start_shift; /* = L */
    int tid_shift;  /* = L + 31 */
    int tid_bits;   /* 63 - tid_shift */
};

IPX_HD uint64_t make_key(const KeyLayout &kl, int32_t tid, int64_t i0, int64_t i1) {
    return ((uint64_t)(uint32_t)tid << kl.tid_shift) | ((uint64_t)i0 << kl.start_shift) | (uint64_t)(i1 - i0);
}
IPX_HD int32_t key_tid(const KeyLayout &kl, uint64_t k) { return (int32_t)(k >> kl.tid_shift); }
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

/* device-side view of the tables, passed to kernels by value */
struct Tables {
    unsigned long long *keys; /* [cap] */
    JStat *stat;              /* [cap] */
    HEntry *hist;             /* [hcap] */
    uint32_t *dense_idx;      /* [cap] row of the junction in `dense`, DENSE_NONE = none */
    uint32_t *dense;          /* [dense_rows][DENSE_W] read counts for offsets < DENSE_W */
    uint64_t dense_rows;
    uint64_t cap_mask, hcap_mask;
    /* counters / flags, one cache line each to avoid false sharing */
    unsigned long long *n_keys;    /* distinct keys inserted */
    unsigned long long *n_hist;    /* distinct (slot, offset) pairs inserted */
    unsigned long long *n_events;  /* events that passed junction_qc */
    unsigned long long *n_live;    /* CIGAR words of live records */
    unsigned long long *err_sv;    /* min over failing records of (record seq << 28 | code); EMPTY64 = none */
    uint32_t *overflow;            /* set when a probe sequence ran out of room */
    KeyLayout kl;
};

#if defined(__CUDACC__)
static const uint32_t MAX_PROBE = 1u << 16;

/* find the slot of `key`, inserting it if absent; returns cap_mask + 1 on overflow.  When this call creates the
 * junction, *new_row receives its dense offset row (if the pool still had one). */
__device__ __forceinline__ uint64_t junction_find_or_insert(const Tables &t, uint64_t key, uint32_t *new_row = nullptr) {
    uint64_t h = mix64(key) & t.cap_mask;
    for (uint32_t probe = 0; probe < MAX_PROBE; ++probe) {
        unsigned long long k = *((volatile unsigned long long *)&t.keys[h]);
        if (k == key) return h;
        if (k == EMPTY64) {
            unsigned long long old = atomicCAS(&t.keys[h], EMPTY64, (unsigned long long)key);
            if (old == EMPTY64) {
                const unsigned long long id = atomicAdd(t.n_keys, 1ull);
                if (id < t.dense_rows) {
                    atomicExch(&t.dense_idx[h], (uint32_t)id);
                    if (new_row) *new_row = (uint32_t)id;
                }
                return h;
            }
            if (old == key) return h;
        }
        h = (h + 1) & t.cap_mask;
    }
    atomicExch(t.overflow, 1u);
    return t.cap_mask + 1;
}

/* lookup only; returns cap_mask + 1 when absent */
__device__ __forceinline__ uint64_t junction_find(const Tables &t, uint64_t key) {
    uint64_t h = mix64(key) & t.cap_mask;
    for (uint32_t probe = 0; probe < MAX_PROBE; ++probe) {
        unsigned long long k = *((volatile unsigned long long *)&t.keys[h]);
        if (k == key) return h;
        if (k == EMPTY64) break;
        h = (h + 1) & t.cap_mask;
    }
    return t.cap_mask + 1;
}

/* hash path of hist_count; returns 1 when a new (slot, offset) entry was created */
__device__ __forceinline__ uint32_t hist_add_hash(const Tables &t, uint64_t slot, uint32_t offset16, uint32_t n) {
    const unsigned long long hk = ((unsigned long long)slot << 16) | offset16;
    uint64_t h = mix64(hk) & t.hcap_mask;
    for (uint32_t probe = 0; probe < MAX_PROBE; ++probe) {
        unsigned long long k = *((volatile unsigned long long *)&t.hist[h].key);
        uint32_t inserted = 0u;
        if (k == EMPTY64) {
            const unsigned long long old = atomicCAS(&t.hist[h].key, EMPTY64, hk);
            inserted = (old == EMPTY64) ? 1u : 0u;
            k = inserted ? hk : old;
        }
        if (k == hk) {
            atomicAdd(&t.hist[h].cnt, n);
            return inserted;
        }
        h = (h + 1) & t.hcap_mask;
    }
    atomicExch(t.overflow, 2u);
    return 0u;
}

/* count n reads at (slot, offset): Junction::count_read's push_back (junctions.hh:163-167), run-length form.
 * Returns the number of hash entries created (callers sum it per block into n_hist). */
__device__ __forceinline__ uint32_t hist_count(const Tables &t, uint64_t slot, uint32_t offset16, uint32_t n) {
    if (offset16 < DENSE_W) {
        /* the row index is published right after the key; a reader that races ahead of it uses the hash instead */
        const uint32_t d = *((volatile uint32_t *)&t.dense_idx[slot]);
        if (d != DENSE_NONE) {
            atomicAdd(&t.dense[(uint64_t)d * DENSE_W + offset16], n);
            return 0u;
        }
    }
    return hist_add_hash(t, slot, offset16, n);
}

/* fold one counted read into a junction's statistics */
__device__ __forceinline__ void stat_accumulate(JStat *s, uint32_t cat, uint32_t a0, uint32_t a1) {
    atomicAdd(&s->cnt[cat], 1ull);
    /* min/max only ever tighten, so a stale read can only cause a redundant atomic, never a missed one */
    if (a0 < *((volatile uint32_t *)&s->a0)) atomicMin(&s->a0, a0);
    if (a1 > *((volatile uint32_t *)&s->a1)) atomicMax(&s->a1, a1);
}

__device__ __forceinline__ void stat_first(JStat *s, int cls, uint64_t sv) {
    if (sv < *((volatile unsigned long long *)&s->first[cls])) atomicMin(&s->first[cls], (unsigned long long)sv);
}
#endif

}  // namespace ipx
#endif
