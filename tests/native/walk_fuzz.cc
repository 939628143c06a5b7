/* walk_fuzz.cc -- host-side fuzz of the device automaton in intronprospector_b200/csrc/ipx_walk.cuh: the branch-free
 * walk_op_flat and the two-path 32-bit lean_op (with its hand-over to a WalkState when relative positions grow large)
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
        bool same = bad_a == bad_b && a.out.size() == b.out.size() && bad_a == bad_c && a.out.size() == c.out.size();
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
