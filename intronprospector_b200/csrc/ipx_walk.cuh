/* ipx_walk.cuh -- per-alignment CIGAR automaton and read classification, shared by every kernel.
 *
 * The functions are __host__ __device__ so that tests/native can exercise exactly the code the
 * kernels run against the oracle without a GPU; the product only ever calls them from device code.
 *
 * Reference semantics restated here (file:line into the reference's src/):
 *   ReadJunctionExtractor        junctions_extractor.cc:124-319   walk_cigar()
 *   get_junction_strand*         junctions_extractor.cc:322-371   read_strand()
 *   get_read_category            junctions_extractor.cc:59-95     read_category()
 *   junction_qc                  junctions_extractor.cc:384-401   junction_qc()
 */
#ifndef IPX_WALK_CUH
#define IPX_WALK_CUH
#include <stdint.h>

#if defined(__CUDACC__)
#define IPX_HD __host__ __device__ __forceinline__
#else
#define IPX_HD inline
#endif

namespace ipx {

/* automaton states (junctions_extractor.cc:114-119) */
enum : int { ST_LIMBO = 0, ST_LEFT = 1, ST_INTRON = 2, ST_RIGHT = 3 };

/* BAM CIGAR op codes, index into "MIDNSHP=XB" */
enum : uint32_t { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

struct ExtractParams {
    uint32_t min_anchor_length;
    uint32_t max_anchor_indel_size;
    uint32_t min_intron_length;
    uint32_t max_intron_length;
    uint32_t excludes;
    int32_t strandness; /* 0 UN, 1 RF, 2 FR */
};

/* One junction candidate as the automaton hands it to record_junction (junctions_extractor.cc:145-151);
 * only the sums of mismatch + indel counts are consumed downstream. */
struct Candidate {
    int64_t anchor_start, intron_start, intron_end, anchor_end;
    uint32_t left_bad, right_bad; /* mismatch + indel counts, uint32 wrap-around as in the reference */
    uint32_t ordinal;             /* position of this candidate among the read's candidates */
};

/* get_junction_strand (junctions_extractor.cc:362-371): flag rule when stranded, else XS:A, else '.' */
IPX_HD char read_strand(int32_t strandness, uint32_t flag, uint32_t xs) {
    if (strandness != 0) {
        /* get_junction_strand_flag (junctions_extractor.cc:333-359) */
        uint32_t reversed = (flag >> 4) & 1u, mate_reversed = (flag >> 5) & 1u;
        uint32_t first_strand = ((flag >> 6) & 1u) ^ reversed;
        uint32_t second_strand = ((flag >> 7) & 1u) ^ mate_reversed;
        uint32_t flip = (strandness == 1) ? 1u : 0u; /* "^= !(strandness_ - 1)": RF (1) flips, FR (2) keeps */
        first_strand ^= flip;
        second_strand ^= flip;
        if (first_strand != second_strand) return '.';
        return first_strand ? '+' : '-';
    }
    if (xs == 0u || xs == 0xFFu) return '.'; /* no tag, or a tag that is not a non-NUL 'A' value */
    return (char)xs;
}

/* get_read_category (junctions_extractor.cc:71-95); nh_class: 0 none/<=0, 1 ==1, 2 >1 */
IPX_HD uint32_t read_category(uint32_t flag, uint32_t nh_class) {
    uint32_t by_tag = (nh_class == 1u) ? 0u : ((nh_class == 2u) ? 1u : 2u);
    if (flag & (0x100u | 0x800u)) return 1u;
    if (flag & 0x1u) return ((flag & 0x2u) && !(flag & 0x4u)) ? by_tag : 2u;
    return by_tag;
}

/* process_alignment's skip test (junctions_extractor.cc:536) */
IPX_HD bool record_is_live(uint32_t n_cigar, int32_t tid, uint32_t flag) {
    return !(n_cigar <= 1u || tid < 0 || (flag & (0x200u | 0x400u)));
}

/* the per-read half of junction_qc (junctions_extractor.cc:388-391) */
IPX_HD bool read_may_emit(const ExtractParams &p, uint32_t flag, uint32_t nh_class) {
    if (flag & 0x100u) return false;
    if ((p.excludes & 1u) && nh_class == 2u) return false;
    return true;
}

/* the per-candidate half of junction_qc (junctions_extractor.cc:392-397), signed 64-bit like hts_pos_t */
IPX_HD bool junction_qc(const ExtractParams &p, const Candidate &c) {
    int64_t ilen = c.intron_end - c.intron_start;
    if (ilen < (int64_t)p.min_intron_length || ilen > (int64_t)p.max_intron_length) return false;
    int64_t left = (c.intron_start - c.anchor_start) - (int64_t)c.left_bad;
    int64_t right = (c.anchor_end - c.intron_end) - (int64_t)c.right_bad;
    return !(left < (int64_t)p.min_anchor_length || right < (int64_t)p.min_anchor_length);
}

/* Automaton registers; kept in a struct so a kernel can carry it across CIGAR tiles. */
struct WalkState {
    int64_t tpos, a0, i0, i1, a1;
    uint32_t lmm, rmm, lind, rind, cur;
    uint32_t n_emitted;
    int state;
    IPX_HD void init(int64_t pos) {
        tpos = pos;
        a0 = i0 = i1 = a1 = -1;
        lmm = rmm = lind = rind = cur = 0u;
        n_emitted = 0u;
        state = ST_LIMBO;
    }
};

#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class Emit>
IPX_HD void emit_candidate(WalkState &w, Emit &emit) {
    Candidate c;
    c.anchor_start = w.a0;
    c.intron_start = w.i0;
    c.intron_end = w.i1;
    c.anchor_end = w.a1;
    c.left_bad = w.lmm + w.lind;
    c.right_bad = w.rmm + w.rind;
    c.ordinal = w.n_emitted++;
    emit(c);
}

/* Feed one CIGAR word.  Returns 0, or 0x100 | op for an operation the reference rejects
 * ("Unknown cigar operation", junctions_extractor.cc:312-313). */
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class Emit>
IPX_HD uint32_t walk_op(WalkState &w, uint32_t word, uint32_t max_indel, Emit &emit) {
    const uint32_t op = word & 15u;
    const uint32_t len = word >> 4;
    if (op == OP_M || op == OP_EQ || op == OP_X) {
        /* enter_aligned (junctions_extractor.cc:232-254) */
        if (w.state == ST_LIMBO) {
            w.a0 = w.tpos;
            w.lmm = w.rmm = 0u; /* indel counters are deliberately not reset */
            w.state = ST_LEFT;
        } else if (w.state == ST_INTRON) {
            w.i1 = w.tpos;
            w.state = ST_RIGHT;
        }
        w.a1 = w.tpos + len;
        if (op == OP_X) {
            if (w.state == ST_LEFT) w.lmm += len; else w.rmm += len;
        }
        w.cur = 0u;
        w.tpos += len;
    } else if (op == OP_N) {
        /* enter_intron (junctions_extractor.cc:208-229) */
        if (w.state == ST_RIGHT) {
            emit_candidate(w, emit);
            w.a0 = w.i1;
            w.lmm = w.rmm; w.rmm = 0u;
            w.lind = w.rind; w.rind = 0u;
            w.state = ST_LEFT;
        }
        if (w.state == ST_LEFT) {
            w.i0 = w.tpos;
            w.i1 = w.tpos + len;
            w.state = ST_INTRON;
        }
        w.cur = 0u;
        w.tpos += len;
    } else if (op == OP_D || op == OP_I) {
        /* enter_query_delete / enter_query_insert (junctions_extractor.cc:257-275) */
        if (w.state == ST_LEFT) w.lind += len; else if (w.state == ST_RIGHT) w.rind += len;
        w.cur += len;
        if (w.cur > max_indel) {
            if (w.state == ST_RIGHT) emit_candidate(w, emit);
            w.state = ST_LIMBO;
        }
        if (op == OP_D) w.tpos += len;
    } else if (op == OP_S || op == OP_H) {
        /* enter_limbo (junctions_extractor.cc:153-158); cur is not reset */
        if (w.state == ST_RIGHT) emit_candidate(w, emit);
        w.state = ST_LIMBO;
    } else {
        return 0x100u | op;
    }
    return 0u;
}

/* The same transition function without data-dependent branches (selects only; the rare emission stays a branch).
 * In a warp that walks 32 different reads, walk_op executes every branch some lane takes -- with long reads that is
 * all of them, every step; this form executes one straight-line body.  tests/native/walk_fuzz.cc holds the two
 * equal on random CIGARs. */
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class Emit>
IPX_HD uint32_t walk_op_flat(WalkState &w, uint32_t word, uint32_t max_indel, Emit &emit) {
    const uint32_t op = word & 15u;
    const uint32_t len = word >> 4;
    const bool is_al = ((0x181u >> op) & 1u) != 0u; /* M = X */
    const bool is_x = op == OP_X;
    const bool is_n = op == OP_N;
    const bool is_di = op == OP_I || op == OP_D;
    const bool is_cl = op == OP_S || op == OP_H;
    if (!(is_al | is_n | is_di | is_cl)) return 0x100u | op;
    const int st = w.state;
    /* enter_query_delete / enter_query_insert: the indel is counted before the break test */
    w.lind += (is_di && st == ST_LEFT) ? len : 0u;
    w.rind += (is_di && st == ST_RIGHT) ? len : 0u;
    const uint32_t cur = is_di ? w.cur + len : (is_cl ? w.cur : 0u);
    const bool brk = is_di && cur > max_indel;
    if (st == ST_RIGHT && (is_n | is_cl | brk)) emit_candidate(w, emit);
    /* enter_intron on a right anchor: it becomes the next junction's left anchor */
    const bool shift = is_n && st == ST_RIGHT;
    w.a0 = shift ? w.i1 : w.a0;
    w.lmm = shift ? w.rmm : w.lmm;
    w.lind = shift ? w.rind : w.lind;
    w.rmm = shift ? 0u : w.rmm;
    w.rind = shift ? 0u : w.rind;
    const bool open_intron = is_n && (st == ST_RIGHT || st == ST_LEFT);
    w.i0 = open_intron ? w.tpos : w.i0;
    /* enter_aligned */
    const bool from_limbo = is_al && st == ST_LIMBO, from_intron = is_al && st == ST_INTRON;
    w.a0 = from_limbo ? w.tpos : w.a0;
    w.lmm = from_limbo ? 0u : w.lmm;
    w.rmm = from_limbo ? 0u : w.rmm;
    w.i1 = open_intron ? w.tpos + len : (from_intron ? w.tpos : w.i1);
    w.a1 = is_al ? w.tpos + len : w.a1;
    int ns = st;
    ns = (is_cl | brk) ? ST_LIMBO : ns;
    ns = open_intron ? ST_INTRON : ns;
    ns = from_limbo ? ST_LEFT : ns;
    ns = from_intron ? ST_RIGHT : ns;
    w.lmm += (is_x && ns == ST_LEFT) ? len : 0u;
    w.rmm += (is_x && ns != ST_LEFT) ? len : 0u;
    w.state = ns;
    w.cur = cur;
    w.tpos += (is_al | is_n | (op == OP_D)) ? (int64_t)len : 0;
    return 0u;
}

/* finish (junctions_extractor.cc:278-282) */
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class Emit>
IPX_HD void walk_finish(WalkState &w, Emit &emit) {
    if (w.state == ST_RIGHT) emit_candidate(w, emit);
}

/* ---- the same automaton for long CIGARs: two paths, 32-bit registers ----
 * In a long read nearly every operation is an aligned run or a short indel inside an anchor, which changes no state:
 * those take a path of a dozen instructions.  What makes it short:
 *   - positions are kept relative to the record's start (32 bits; the caller checks `tpos` against LEAN_TPOS_LIMIT
 *     every eighth operation and converts to a WalkState long before a wrap-around),
 *   - the mismatch / indel counters are kept as "anchor being extended" (cm, ci) and "the other one" (om, oi), so the
 *     common path does not ask which anchor it is in.  In ST_RIGHT the extended anchor is the right one; in every
 *     other state (c*) stands for the left counters, (o*) for the right ones.
 * Everything else (N, clips, an indel run that breaks the anchor, the first aligned run after limbo or an intron,
 * unknown operations) takes the general path.  tests/native/walk_fuzz.cc holds it equal to walk_op. */
struct LeanState {
    uint32_t tpos, a0, i0, i1, a1; /* relative to the record's pos */
    uint32_t cm, ci, om, oi, cur, n_emitted;
    int state;
    IPX_HD void init() {
        tpos = a0 = i0 = i1 = a1 = 0u;
        cm = ci = om = oi = cur = n_emitted = 0u;
        state = ST_LIMBO;
    }
    /* the equivalent WalkState, to continue with walk_op_flat */
    IPX_HD void widen(int64_t pos, WalkState &w) const {
        w.tpos = pos + (int64_t)tpos;
        /* registers the automaton has not set yet are never read before it sets them */
        w.a0 = pos + (int64_t)a0; w.i0 = pos + (int64_t)i0; w.i1 = pos + (int64_t)i1; w.a1 = pos + (int64_t)a1;
        const bool right = state == ST_RIGHT;
        w.lmm = right ? om : cm; w.lind = right ? oi : ci;
        w.rmm = right ? cm : om; w.rind = right ? ci : oi;
        w.cur = cur;
        w.n_emitted = n_emitted;
        w.state = state;
    }
};
/* tested before every operation whose index in the CIGAR buffer is a multiple of 8: from below the limit, the up to
 * 8 operations of < 2^28 each until the next test stay below 2^32 (a record's first test comes after at most 7) */
static const uint32_t LEAN_TPOS_LIMIT = 0x70000000u;

#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class Emit>
IPX_HD void lean_emit(LeanState &w, int64_t pos, Emit &emit) {
    Candidate c;
    c.anchor_start = pos + (int64_t)w.a0;
    c.intron_start = pos + (int64_t)w.i0;
    c.intron_end = pos + (int64_t)w.i1;
    c.anchor_end = pos + (int64_t)w.a1;
    c.left_bad = w.om + w.oi; /* only called in ST_RIGHT */
    c.right_bad = w.cm + w.ci;
    c.ordinal = w.n_emitted++;
    emit(c);
}

#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
/* `bad` receives 0x100 | op of the first operation the reference rejects; the caller may test it as rarely as it likes
 * (what is emitted after such an operation is moot: the run fails with the reference's error) */
template <class Emit>
IPX_HD void lean_op(LeanState &w, int64_t pos, uint32_t word, uint32_t max_indel, Emit &emit, uint32_t &bad) {
    const uint32_t op = word & 15u;
    const uint32_t len = word >> 4;
    const uint32_t bit = 1u << op; /* one shift, then every class test is one AND */
    const bool is_al = (bit & 0x181u) != 0u; /* M = X */
    const bool is_di = (bit & 0x006u) != 0u; /* I D */
    const bool in_anchor = (w.state & 1) != 0; /* ST_LEFT or ST_RIGHT */
    const uint32_t ncur = is_di ? w.cur + len : 0u;
    /* the common path, written as selects on `fast` so that it shares no register moves with the branch below */
    const bool fast = is_al ? in_anchor : (is_di && ncur <= max_indel);
    const uint32_t flen = fast ? len : 0u;
    w.ci += (is_di && in_anchor) ? flen : 0u;
    w.cm += (bit & 0x100u) ? flen : 0u; /* X */
    w.tpos += (bit & 0x002u) ? 0u : flen; /* everything on this path but I advances the reference */
    w.a1 = (fast && is_al) ? w.tpos : w.a1;
    w.cur = fast ? ncur : w.cur;
    if (fast) return;
    const bool was_right = w.state == ST_RIGHT;
    if (is_al) {
        if (w.state == ST_LIMBO) {
            w.a0 = w.tpos;
            w.cm = 0u; w.om = 0u; /* the indel counters are deliberately not reset */
            w.state = ST_LEFT;
        } else { /* ST_INTRON: the right anchor starts; it becomes the one being extended */
            w.i1 = w.tpos;
            uint32_t x = w.cm; w.cm = w.om; w.om = x;
            x = w.ci; w.ci = w.oi; w.oi = x;
            w.state = ST_RIGHT;
        }
        w.tpos += len;
        w.a1 = w.tpos;
        w.cm += (op == OP_X) ? len : 0u;
        w.cur = 0u;
    } else if (op == OP_N) {
        if (was_right) {
            lean_emit(w, pos, emit);
            w.a0 = w.i1; /* the right anchor is the next junction's left anchor: its counters stay in (c*) */
            w.om = 0u; w.oi = 0u;
        }
        if (in_anchor) {
            w.i0 = w.tpos;
            w.i1 = w.tpos + len;
            w.state = ST_INTRON;
        }
        w.cur = 0u;
        w.tpos += len;
    } else if (is_di || op == OP_S || op == OP_H) {
        /* an indel run longer than max_indel, or a clip: the anchor ends here */
        w.ci += (is_di && in_anchor) ? len : 0u; /* counted before the break test, as in the reference */
        w.cur = is_di ? ncur : w.cur;            /* a clip leaves the run length alone */
        if (was_right) {
            lean_emit(w, pos, emit);
            uint32_t x = w.cm; w.cm = w.om; w.om = x;
            x = w.ci; w.ci = w.oi; w.oi = x;
        }
        w.state = ST_LIMBO;
        w.tpos += (op == OP_D) ? len : 0u;
    } else if (!bad) {
        bad = 0x100u | op;
    }
}

#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class Emit>
IPX_HD void lean_finish(LeanState &w, int64_t pos, Emit &emit) {
    if (w.state == ST_RIGHT) lean_emit(w, pos, emit);
}

/* ---- one long CIGAR walked by many lanes ----
 * A record is cut into segments that start right after an aligned operation: there the automaton is in ST_LEFT or
 * ST_RIGHT with an empty indel run, and what it does next depends on nothing else -- only the VALUES of its registers
 * do, and those are carried along as they are (a register at a segment's end is either a value made inside the
 * segment, or one of the registers it was entered with, plus what the segment added).  So:
 *   pass A  a lane per segment walks its operations once and writes that mapping (SegSummary); the two possible entry
 *           states part ways at the first operation that changes state and run identically afterwards, so one walk
 *           carries both;
 *   stitch  a thread per record applies the mappings in order: every segment's entry registers, concrete;
 *   pass C  a lane per segment walks it again from those registers with lean_op and emits.
 * Positions are relative to the record's pos as in LeanState; a record whose span comes near 2^32, or with an unknown
 * operation, is flagged (bad) and walked by one lane as usual.  tests/native/walk_fuzz.cc holds the three passes
 * equal to walk_op for random segment lengths. */
enum : uint32_t { SRC_NONE = 0, SRC_CM, SRC_CI, SRC_OM, SRC_OI, SRC_A0, SRC_I0, SRC_I1, SRC_A1 };
struct SegForm {
    uint32_t val; /* counters: added to the source; positions: relative to the segment's first operation (no source) */
    uint32_t src;
};
struct SegSummary {
    uint32_t dtpos, cur, n_cand; /* n_cand: candidates the segment emits whatever state it is entered in */
    uint8_t synced, state, e0, bad; /* e0: entered in ST_RIGHT it emits one more, at its first state-changing operation */
    SegForm a1;
    SegForm a0[2], i0[2], i1[2], cm[2], ci[2], om[2], oi[2]; /* [0] entered in ST_LEFT (or at the record start), [1] in ST_RIGHT */
};

struct SegWalk {
    SegSummary s;
    int state;
    IPX_HD void init(bool record_start) {
        s.dtpos = s.cur = s.n_cand = 0u;
        s.e0 = s.bad = 0;
        s.a1.val = 0u; s.a1.src = SRC_A1;
        for (int v = 0; v < 2; ++v) {
            s.a0[v].val = s.i0[v].val = s.i1[v].val = s.cm[v].val = s.ci[v].val = s.om[v].val = s.oi[v].val = 0u;
            s.a0[v].src = SRC_A0; s.i0[v].src = SRC_I0; s.i1[v].src = SRC_I1;
            s.cm[v].src = SRC_CM; s.ci[v].src = SRC_CI; s.om[v].src = SRC_OM; s.oi[v].src = SRC_OI;
        }
        /* a record starts in limbo: nothing to leave open */
        s.synced = record_start ? 1 : 0;
        state = record_start ? ST_LIMBO : ST_LEFT; /* before the first state change: "in an anchor" */
    }
    IPX_HD void finish() { s.state = (uint8_t)state; }
};

IPX_HD void seg_swap(SegForm &a, SegForm &b) {
    const SegForm t = a;
    a = b;
    b = t;
}
IPX_HD SegForm seg_local(uint32_t v) {
    SegForm f;
    f.val = v;
    f.src = SRC_NONE;
    return f;
}

IPX_HD void seg_op(SegWalk &w, uint32_t word, uint32_t max_indel) {
    SegSummary &s = w.s;
    const uint32_t op = word & 15u;
    const uint32_t len = word >> 4;
    const bool is_al = ((0x181u >> op) & 1u) != 0u;
    const bool is_di = ((0x006u >> op) & 1u) != 0u;
    const bool is_n = op == OP_N, is_cl = op == OP_S || op == OP_H;
    if (!(is_al | is_di | is_n | is_cl) || s.dtpos >= LEAN_TPOS_LIMIT) {
        s.bad = 1;
        return;
    }
    const bool in_anchor = (w.state & 1) != 0;
    const uint32_t ncur = is_di ? s.cur + len : 0u;
    if (is_al ? in_anchor : (is_di && ncur <= max_indel)) {
        if (is_di && in_anchor) { s.ci[0].val += len; s.ci[1].val += len; }
        if (op == OP_X) { s.cm[0].val += len; s.cm[1].val += len; }
        if (op != OP_I) s.dtpos += len;
        if (is_al) s.a1 = seg_local(s.dtpos);
        s.cur = ncur;
        return;
    }
    if (!s.synced) {
        /* the first operation that changes state: N, a clip, or an indel run that breaks the anchor.  Entered in
         * ST_RIGHT the segment emits the junction it was handed, and the two entry states are one from here on */
        s.synced = 1;
        s.e0 = 1;
        if (is_di) { s.ci[0].val += len; s.ci[1].val += len; }
        if (is_n) {
            s.a0[1] = s.i1[1];
            s.om[1] = seg_local(0u); s.oi[1] = seg_local(0u);
            s.i0[0] = s.i0[1] = seg_local(s.dtpos);
            s.i1[0] = s.i1[1] = seg_local(s.dtpos + len);
            w.state = ST_INTRON;
            s.cur = 0u;
            s.dtpos += len;
        } else {
            seg_swap(s.cm[1], s.om[1]);
            seg_swap(s.ci[1], s.oi[1]);
            w.state = ST_LIMBO;
            s.cur = is_di ? ncur : s.cur;
            s.dtpos += (op == OP_D) ? len : 0u;
        }
        return;
    }
    const bool was_right = w.state == ST_RIGHT;
    if (is_al) {
        for (int v = 0; v < 2; ++v) {
            if (w.state == ST_LIMBO) {
                s.a0[v] = seg_local(s.dtpos);
                s.cm[v] = seg_local(0u); s.om[v] = seg_local(0u);
            } else {
                s.i1[v] = seg_local(s.dtpos);
                seg_swap(s.cm[v], s.om[v]);
                seg_swap(s.ci[v], s.oi[v]);
            }
            if (op == OP_X) s.cm[v].val += len;
        }
        w.state = w.state == ST_LIMBO ? ST_LEFT : ST_RIGHT;
        s.dtpos += len;
        s.a1 = seg_local(s.dtpos);
        s.cur = 0u;
    } else if (is_n) {
        if (was_right) {
            s.n_cand++;
            for (int v = 0; v < 2; ++v) {
                s.a0[v] = s.i1[v];
                s.om[v] = seg_local(0u); s.oi[v] = seg_local(0u);
            }
        }
        if (in_anchor) {
            s.i0[0] = s.i0[1] = seg_local(s.dtpos);
            s.i1[0] = s.i1[1] = seg_local(s.dtpos + len);
            w.state = ST_INTRON;
        }
        s.cur = 0u;
        s.dtpos += len;
    } else {
        for (int v = 0; v < 2; ++v) {
            if (is_di && in_anchor) s.ci[v].val += len;
            if (was_right) {
                seg_swap(s.cm[v], s.om[v]);
                seg_swap(s.ci[v], s.oi[v]);
            }
        }
        if (was_right) s.n_cand++;
        s.cur = is_di ? ncur : s.cur;
        w.state = ST_LIMBO;
        s.dtpos += (op == OP_D) ? len : 0u;
    }
}

IPX_HD uint32_t seg_position(const SegForm &f, const LeanState &e) {
    switch (f.src) {
        case SRC_A0: return e.a0;
        case SRC_I0: return e.i0;
        case SRC_I1: return e.i1;
        case SRC_A1: return e.a1;
        default: return e.tpos + f.val;
    }
}
IPX_HD uint32_t seg_counter(const SegForm &f, const LeanState &e) {
    switch (f.src) {
        case SRC_CM: return e.cm + f.val;
        case SRC_CI: return e.ci + f.val;
        case SRC_OM: return e.om + f.val;
        case SRC_OI: return e.oi + f.val;
        default: return f.val;
    }
}

/* the registers a segment is left with, from the registers it is entered with; false: walk the record in one piece.
 * (Not for empty segments: those assume nothing about their entry and change nothing.) */
IPX_HD bool seg_apply(const SegSummary &s, const LeanState &e, LeanState &x) {
    if (s.bad || (uint64_t)e.tpos + s.dtpos >= LEAN_TPOS_LIMIT) return false;
    const int v = e.state == ST_RIGHT ? 1 : 0;
    x.tpos = e.tpos + s.dtpos;
    x.a0 = seg_position(s.a0[v], e);
    x.i0 = seg_position(s.i0[v], e);
    x.i1 = seg_position(s.i1[v], e);
    x.a1 = seg_position(s.a1, e);
    x.cm = seg_counter(s.cm[v], e);
    x.ci = seg_counter(s.ci[v], e);
    x.om = seg_counter(s.om[v], e);
    x.oi = seg_counter(s.oi[v], e);
    x.cur = s.cur;
    x.n_emitted = e.n_emitted + s.n_cand + ((v == 1 && s.e0) ? 1u : 0u);
    x.state = s.synced ? (int)s.state : e.state;
    return true;
}

/* first operation of segment j (j >= 1) of a record with operations [0, n): the first index at or after j * stride
 * whose predecessor is an aligned operation, n when there is none */
template <class Word>
IPX_HD uint32_t seg_start(const Word &word_at, uint32_t n, uint32_t j, uint32_t stride) {
    uint64_t k = (uint64_t)j * stride;
    if (j == 0u) return 0u;
    for (; k < n; ++k)
        if ((0x181u >> (word_at((uint32_t)k - 1u) & 15u)) & 1u) return (uint32_t)k;
    return n;
}

}  // namespace ipx
#endif
