/* walk_fuzz.cc -- host-side fuzz of the device automaton in intronprospector_b200/csrc/ipx_walk.cuh: the branch-free
 * walk_op_flat, the two-path 32-bit lean_op (with its hand-over to a WalkState when relative positions grow large)
 * and the segmented walk of long records (seg_op summaries, seg_apply stitch, lean_op per segment)
 * must produce exactly the candidates (and error codes) of the branching walk_op, which the GPU parity tests pin
 * against the oracle.  Test infrastructure; built and run by tests/test_walk_native.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../intronprospector_b200/csrc/ipx_walk.cuh"

using namespace ipx;

struct Collect {
    std::vector<Candidate> out;
    void operator()(const Candidate &c) { out.push_back(c); }
};

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

int main(int argc, char **argv) {
    const long n_cases = argc > 1 ? atol(argv[1]) : 200000;
    long emitted = 0;
    for (long t = 0; t < n_cases; ++t) {
        const uint32_t n_ops = 1 + rnd() % (t % 7 == 0 ? 60 : 12);
        const uint32_t max_indel = (t & 1) ? 0u : (rnd() % 3 == 0 ? 36u : rnd() % 8);
        std::vector<uint32_t> cig(n_ops);
        for (uint32_t k = 0; k < n_ops; ++k) {
            static const uint32_t ops[] = {0, 0, 0, 0, 3, 3, 1, 2, 4, 5, 7, 8, 8, 0, 3, 6, 9, 12};
            uint32_t op = ops[rnd() % (t % 50 == 0 ? 18 : 15)];
            uint32_t len = (rnd() % 4 == 0) ? 1 + rnd() % 3 : 1 + rnd() % 200;
            if (op == 3 && rnd() % 2) len = 70 + rnd() % 5000;
            if (rnd() % 97 == 0) len = 0; /* zero-length operations must behave identically too */
            if (t % 11 == 0 && rnd() % 3 == 0) len = rnd() & 0x0FFFFFFFu; /* up to the 28 bits a BAM operation holds */
            cig[k] = (len << 4) | op;
        }
        const int64_t pos = (int64_t)(rnd() % 1000000) - (t % 1000 == 0 ? 10 : 0);
        Collect a, b;
        WalkState wa, wb;
        wa.init(pos);
        wb.init(pos);
        uint32_t bad_a = 0, bad_b = 0;
        for (uint32_t k = 0; k < n_ops && !bad_a; ++k) bad_a = walk_op(wa, cig[k], max_indel, a);
        if (!bad_a) walk_finish(wa, a);
        for (uint32_t k = 0; k < n_ops && !bad_b; ++k) bad_b = walk_op_flat(wb, cig[k], max_indel, b);
        if (!bad_b) walk_finish(wb, b);
        /* the long-read walker as the kernel drives it: lean_op, the position limit tested every eighth operation
         * (counted from an arbitrary phase, like a record that starts anywhere in a 32-byte group), then the rest of
         * the record on a widened state; every few cases with a tiny limit so the hand-over happens in every state */
        Collect c;
        uint32_t bad_c = 0;
        {
            const uint32_t limit = (t % 5 == 0) ? rnd() % 3000 : LEAN_TPOS_LIMIT;
            const uint32_t phase = rnd() & 7u;
            LeanState lw;
            lw.init();
            uint32_t k = 0;
            bool wide = false;
            for (; k < n_ops && !bad_c; ++k) {
                if (((k + phase) & 7u) == 0u && lw.tpos >= limit) { wide = true; break; }
                lean_op(lw, pos, cig[k], max_indel, c, bad_c);
            }
            if (wide) {
                WalkState wc;
                lw.widen(pos, wc);
                for (; k < n_ops && !bad_c; ++k) bad_c = walk_op_flat(wc, cig[k], max_indel, c);
                if (!bad_c) walk_finish(wc, c);
            } else if (!bad_c) {
                lean_finish(lw, pos, c);
            }
        }
        /* the segmented walk (seg_op / seg_apply / lean_op per segment), segment length drawn per case */
        Collect d;
        bool seg_fallback = false, seg_consistent = true;
        {
            const uint32_t stride = (t % 3 == 0) ? 1 + rnd() % 9 : (t % 3 == 1 ? 1 + rnd() % 3 : 16);
            const uint32_t m = (n_ops + stride - 1) / stride;
            auto word_at = [&](uint32_t k) { return cig[k]; };
            std::vector<uint32_t> start(m + 1);
            for (uint32_t j = 0; j < m; ++j) start[j] = seg_start(word_at, n_ops, j, stride);
            start[m] = n_ops;
            std::vector<SegSummary> sum(m);
            for (uint32_t j = 0; j < m; ++j) {
                SegWalk sw;
                sw.init(j == 0);
                for (uint32_t k = start[j]; k < start[j + 1]; ++k) seg_op(sw, cig[k], max_indel);
                sw.finish();
                sum[j] = sw.s;
            }
            std::vector<LeanState> entry(m + 1);
            entry[0].init();
            for (uint32_t j = 0; j < m && !seg_fallback; ++j) {
                if (start[j] == start[j + 1]) entry[j + 1] = entry[j]; /* empty: nothing is walked, nothing assumed */
                else seg_fallback = !seg_apply(sum[j], entry[j], entry[j + 1]);
            }
            if (!seg_fallback) {
                LeanState w;
                w.init();
                uint32_t bad_d = 0;
                for (uint32_t j = 0; j < m; ++j) {
                    w = entry[j];
                    for (uint32_t k = start[j]; k < start[j + 1]; ++k) lean_op(w, pos, cig[k], max_indel, d, bad_d);
                    const LeanState &x = entry[j + 1];
                    /* what the stitch predicted is what the walk arrives at (registers the automaton may read again) */
                    seg_consistent = seg_consistent && w.tpos == x.tpos && w.state == x.state && w.cur == x.cur &&
                                     w.n_emitted == x.n_emitted && w.cm == x.cm && w.ci == x.ci && w.om == x.om && w.oi == x.oi &&
                                     w.a1 == x.a1 && (w.state == ST_LIMBO || w.a0 == x.a0) &&
                                     ((w.state != ST_INTRON && w.state != ST_RIGHT) || (w.i0 == x.i0 && w.i1 == x.i1));
                }
                lean_finish(w, pos, d);
                seg_consistent = seg_consistent && bad_d == 0;
            }
        }
        bool same = bad_a == bad_b && a.out.size() == b.out.size() && bad_a == bad_c && a.out.size() == c.out.size();
        if (same && !seg_fallback) {
            same = seg_consistent && bad_a == 0 && a.out.size() == d.out.size();
            for (size_t i = 0; same && i < a.out.size(); ++i) {
                const Candidate &x = a.out[i], &z = d.out[i];
                same = x.anchor_start == z.anchor_start && x.intron_start == z.intron_start && x.intron_end == z.intron_end &&
                       x.anchor_end == z.anchor_end && x.left_bad == z.left_bad && x.right_bad == z.right_bad && x.ordinal == z.ordinal;
            }
            if (!same) printf("(segmented walk differs; consistent %d, %zu candidates)\n", (int)seg_consistent, d.out.size());
        }
        if (same && seg_fallback && bad_a == 0) {
            /* only spans near 2^32 may send a well-formed record to the one-piece walk */
            uint64_t span = 0;
            for (uint32_t k = 0; k < n_ops; ++k) span += cig[k] >> 4;
            same = span >= LEAN_TPOS_LIMIT;
            if (!same) printf("(segmented walk fell back without need)\n");
        }
        for (size_t i = 0; same && i < a.out.size(); ++i) {
            const Candidate &x = a.out[i], &y = b.out[i], &z = c.out[i];
            same = x.anchor_start == y.anchor_start && x.intron_start == y.intron_start && x.intron_end == y.intron_end &&
                   x.anchor_end == y.anchor_end && x.left_bad == y.left_bad && x.right_bad == y.right_bad && x.ordinal == y.ordinal &&
                   x.anchor_start == z.anchor_start && x.intron_start == z.intron_start && x.intron_end == z.intron_end &&
                   x.anchor_end == z.anchor_end && x.left_bad == z.left_bad && x.right_bad == z.right_bad && x.ordinal == z.ordinal;
        }
        if (!same) {
            printf("MISMATCH case %ld (max_indel %u, pos %lld): bad %u/%u/%u, %zu/%zu/%zu candidates\n  cigar:", t, max_indel,
                   (long long)pos, bad_a, bad_b, bad_c, a.out.size(), b.out.size(), c.out.size());
            for (uint32_t k = 0; k < n_ops; ++k) printf(" %u%c", cig[k] >> 4, "MIDNSHP=XB??????"[cig[k] & 15]);
            printf("\n");
            return 1;
        }
        emitted += (long)a.out.size();
    }
    printf("ok: %ld cases, %ld candidates\n", n_cases, emitted);
    return 0;
}
