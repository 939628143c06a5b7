"""Synthetic per-sample intron-call tables for the merge path (BASELINE config 5): every sample is a random subset of
one junction universe with its own counts, overhangs and confidence, a few '.'-strand rows and deliberate strand
conflicts.  Test data only."""
import numpy as np

HEADER = ("chrom\tintron_start\tintron_end\tstrand\tuniq_mapped_count\tmulti_mapped_count\tunsure_mapped_count\t"
          "max_left_overhang\tmax_right_overhang\tconfidence\tsplice_sites\tname\n")
MOTIFS = ["GT/AG", "GC/AG", "AT/AC", "ct/ac", "gg/tt"]


def write_tables(dirname, n_samples, n_universe, n_contigs=12, keep=0.6, seed=5, conflicts=5):
    rng = np.random.default_rng(seed)
    chrom = rng.integers(0, n_contigs, n_universe)
    start = rng.integers(1000, 200_000_000, n_universe)
    end = start + rng.integers(70, 100_000, n_universe)
    strand = rng.choice(np.array(list("+-")), n_universe)
    motif = rng.choice(len(MOTIFS), n_universe, p=[0.9, 0.04, 0.02, 0.02, 0.02])
    names = ["chr%d" % (c + 1) for c in range(n_contigs)]
    conflict_keys = set(int(x) for x in rng.choice(n_universe, conflicts, replace=False))
    paths = []
    for s in range(n_samples):
        pick = np.flatnonzero(rng.random(n_universe) < keep)
        order = np.lexsort((end[pick], start[pick], chrom[pick]))
        pick = pick[order]
        uniq = rng.integers(0, 500, pick.size)
        multi = rng.integers(0, 50, pick.size)
        unsure = rng.integers(0, 5, pick.size)
        lo = rng.integers(8, 140, pick.size)
        ro = rng.integers(8, 140, pick.size)
        conf = np.round(rng.random(pick.size) * 5, 3)
        dot = rng.random(pick.size) < 0.001
        path = "%s/sample%02d.introns.tsv" % (dirname, s)
        with open(path, "w") as f:
            f.write(HEADER)
            for k, j in enumerate(pick):
                st = strand[j]
                if dot[k]:
                    st = "."
                elif int(j) in conflict_keys and s % 2 == 1:
                    st = "+" if st == "-" else "-"
                f.write("%s\t%d\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%.3g\t%s\tsj%d\n" % (
                    names[chrom[j]], start[j], end[j], st, uniq[k], multi[k], unsure[k], lo[k], ro[k], conf[k],
                    MOTIFS[motif[j]], k))
        paths.append(path)
    return paths, names
