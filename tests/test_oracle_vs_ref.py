"""The plain-C oracle against the unmodified reference binary on seeded synthetic inputs (sizes the CPU finishes
in seconds): short unstranded, stranded with a planted-motif genome under both filters, long reads with and
without anchor indels.  Whole output files must be byte-identical."""
import os

import pytest

import helpers
from intronprospector_b200.synth import LONG, SHORT_STRANDED, SHORT_UNSTRANDED, Synth

CASES = [
    ("short", SHORT_UNSTRANDED, 300_000, [], False),
    ("short_conf0_multi", SHORT_UNSTRANDED, 100_000, ["-C", "0.0", "--exclude=multi", "-a", "12"], False),
    ("stranded_canon", SHORT_STRANDED, 200_000, ["--strandness=RF", "--min-anchor-length=12"], True),
    ("stranded_all", SHORT_STRANDED, 200_000, ["--strandness=RF", "--sj-filter=all", "-C", "0.0"], True),
    ("unstranded_genome", SHORT_UNSTRANDED, 200_000, ["-C", "0.0"], True),
    ("long_indels", LONG, 6_000, ["--allow-anchor-indels", "-C", "0.0"], False),
    ("long_default", LONG, 6_000, [], False),
    ("long_unsorted", LONG, 4_000, ["--unsorted", "--allow-anchor-indels", "-C", "0.0"], False),
]


@pytest.mark.parametrize("name,kind,n,args,genome", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_binary(name, kind, n, args, genome, built, tmp_path):
    syn = Synth(n, kind=kind, genome_len=50_000_000, n_genes=500, n_contigs=5, seed=7 + len(name))
    bam = str(tmp_path / "in.bam")
    syn.write_raw_bam(bam)
    extra = list(args)
    if genome:
        fa = str(tmp_path / "g.fa")
        syn.write_genome(fa)
        extra.append("--genome-fasta=" + fa)
    rc_r, err_r, ref = helpers.run_extract(helpers.REF_EXTRACT, str(tmp_path), "ref", [bam], extra)
    rc_o, err_o, ora = helpers.run_extract(helpers.ORACLE_EXTRACT, str(tmp_path), "ora", [bam], extra)
    assert rc_r == 0, err_r
    assert rc_o == 0, err_o
    assert ref["introns.tsv"].count(b"\n") > 50, "synthetic input produced too few junctions to be a test"
    helpers.assert_same_outputs(ref, ora, name)
