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
/* Read offsets (intron_start - read start) of short reads are < read length; counts for offsets below DENSE_W
 * live in a dense row of DENSE_W counters per table slot (1 KB per slot: 4 GB of HBM for the default 4 M slots,
 * touched only where a junction lives), where counting a read is a single fire-and-forget atomic and the ordered
 * run-length output needs no sort.  Larger offsets (long reads) go to the (junction, offset) hash table, as does
 * everything when the table is too large to give every slot a row. */
static const uint32_t DENSE_W = 256;
/* log of (slot, offset, count) entries for what the dense rows do not take; sorted and run-length reduced at finalize */
static const uint64_t OLOG_INVALID = ~0ull;
static const uint32_t OLOG_BLOCK = 32; /* entries a warp reserves at a time */
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
    uint32_t *slot_list;      /* [cap] slots in use, in insertion order: clear and finalize walk this, never the table */
    uint32_t *slot_id;        /* [cap] position of a slot in slot_list (valid for slots in use) */
    unsigned long long *olog; /* [olog_cap] slot << 32 | read offset << 16 | count, or OLOG_INVALID */
    uint64_t olog_cap;
    uint32_t *dense;          /* [cap][DENSE_W] read counts for offsets < DENSE_W, or NULL when disabled */
    uint64_t cap_mask;
    /* counters / flags, one cache line each to avoid false sharing */
    unsigned long long *n_keys;    /* distinct keys inserted = entries of slot_list */
    unsigned long long *n_olog;    /* entries of olog handed out (whole blocks; unused ones hold OLOG_INVALID) */
    unsigned long long *n_events;  /* events that passed junction_qc */
    unsigned long long *n_live;    /* CIGAR words of live records */
    unsigned long long *err_sv;    /* min over failing records of (record seq << 28 | code); EMPTY64 = none */
    uint32_t *overflow;            /* set when a probe sequence ran out of room */
    KeyLayout kl;
};

#if defined(__CUDACC__)
static const uint32_t MAX_PROBE = 1u << 16;

/* find the slot of `key`, inserting it if absent; returns cap_mask + 1 on overflow */
__device__ __forceinline__ uint64_t junction_find_or_insert(const Tables &t, uint64_t key) {
    uint64_t h = mix64(key) & t.cap_mask;
    for (uint32_t probe = 0; probe < MAX_PROBE; ++probe) {
        unsigned long long k = *((volatile unsigned long long *)&t.keys[h]);
        if (k == key) return h;
        if (k == EMPTY64) {
            unsigned long long old = atomicCAS(&t.keys[h], EMPTY64, (unsigned long long)key);
            if (old == EMPTY64) {
                /* every slot is claimed at most once between clears, so the list cannot overflow */
                const unsigned long long id = atomicAdd(t.n_keys, 1ull);
                t.slot_list[id] = (uint32_t)h;
                t.slot_id[h] = (uint32_t)id;
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

/* Per-warp writer of the offset log.  base/left are warp-uniform; every lane of the (converged) warp calls
 * olog_append together.  A block of OLOG_BLOCK entries is reserved with one atomic and filled chunk by chunk; what is
 * left of a block when the next one is needed, or when the kernel ends (olog_finish), is marked invalid. */
struct OlogWriter {
    unsigned long long base;
    uint32_t left;
};

__device__ __forceinline__ void olog_append(const Tables &t, OlogWriter &w, bool want, unsigned long long entry) {
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, want);
    if (mask == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t m = (uint32_t)__popc(mask);
    if (m > w.left) {
        if (lane < w.left) t.olog[w.base + lane] = OLOG_INVALID;
        unsigned long long nb = 0ull;
        if (lane == 0u) nb = atomicAdd(t.n_olog, (unsigned long long)OLOG_BLOCK);
        nb = __shfl_sync(0xFFFFFFFFu, nb, 0);
        w.base = nb;
        w.left = OLOG_BLOCK;
        if (nb + OLOG_BLOCK > t.olog_cap) {
            atomicExch(t.overflow, 2u); /* the run is lost: ipx_flush reports it */
            w.left = 0u;
            return;
        }
    }
    if (want) t.olog[w.base + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = entry;
    w.base += m;
    w.left -= m;
}

__device__ __forceinline__ void olog_finish(const Tables &t, const OlogWriter &w) {
    const uint32_t lane = threadIdx.x & 31u;
    if (lane < w.left) t.olog[w.base + lane] = OLOG_INVALID;
}

__device__ __forceinline__ unsigned long long olog_entry(uint64_t slot, uint32_t offset16, uint32_t n) {
    return ((unsigned long long)slot << 32) | ((unsigned long long)(offset16 & 0xFFFFu) << 16) | (n & 0xFFFFu);
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
