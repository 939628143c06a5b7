"""The reference's golden-test matrix (reference tests/Makefile.in:67-254) as data.

Every case is run from a scratch directory laid out like the reference's tests/
directory (input/, expected/, output/), so relative paths inside fixtures such as
input/basic134.files.txt keep working.  `prep` steps stand in for the Makefile's
samtools rules (tests/Makefile.in:264-274) and use oracle/_ref/hts_tool.

Each case: name, tool ("extract" | "merge"), args, prep, outputs (produced -> expected,
"gz:" prefix = gunzip before comparing), stderr (expected stderr file or None),
same_as (pairs that must be byte-identical, e.g. pass-through vs. input),
sam_of_bam (BAM outputs to convert to SAM text before a same_as comparison),
allow_fail (exit status may be non-zero).
"""

CONF0 = "--min-confidence-score=0.0"


def _std3(name, exp=None, gz=False):
    exp = exp or name
    sfx = ".gz" if gz else ""
    pre = "gz:" if gz else ""
    return {
        pre + "output/%s.introns.tsv%s" % (name, sfx): "expected/%s.introns.tsv" % exp,
        pre + "output/%s.juncs.bed%s" % (name, sfx): "expected/%s.juncs.bed" % exp,
        pre + "output/%s.introns.bed%s" % (name, sfx): "expected/%s.introns.bed" % exp,
    }


def _std3_args(name, gz=False):
    sfx = ".gz" if gz else ""
    return [
        "--junction-bed=output/%s.juncs.bed%s" % (name, sfx),
        "--intron-bed=output/%s.introns.bed%s" % (name, sfx),
        "--intron-calls=output/%s.introns.tsv%s" % (name, sfx),
    ]


BASIC_SAM = "input/test_basic.sam"
BASIC_BAM = "output/test_basic.bam"
FAKE10_FA = "output/fake10_genome.fa"
SCRAMBLE_FA = "output/fake10_scramble_genome.fa"
STAR_SAM = "input/fake10.star.sam"
HISAT_SAM = "input/fake10.hisat2.sam"
ONT_SAM = "input/ont.sam"

PREP_BAM = [("bam", BASIC_SAM, BASIC_BAM)]
PREP_FA = [("fa", "input/fake10_genome.fa", FAKE10_FA)]
PREP_SCR = [("fa", "input/fake10_scramble_genome.fa", SCRAMBLE_FA)]


def case(name, tool, args, outputs=None, prep=(), stderr=None, same_as=(), sam_of_bam=(), allow_fail=False,
         needs=()):
    return dict(name=name, tool=tool, args=list(args), outputs=dict(outputs or {}), prep=list(prep),
                stderr=stderr, same_as=list(same_as), sam_of_bam=list(sam_of_bam), allow_fail=allow_fail,
                needs=list(needs))


EXTRACT_CASES = [
    case("testBasic1", "extract",
         [CONF0] + _std3_args("testBasic1") + ["--intron-bed6=output/testBasic1.introns.bed6",
                                               "--pass-through=output/testBasic1.sam", BASIC_SAM],
         dict(_std3("testBasic1"), **{"output/testBasic1.introns.bed6": "expected/testBasic1.introns.bed6"}),
         same_as=[("output/testBasic1.sam", BASIC_SAM)], needs=["pass-through"]),
    case("testBasic1gz", "extract",
         [CONF0] + _std3_args("testBasic1gz", gz=True) + ["--intron-bed6=output/testBasic1gz.introns.bed6.gz", BASIC_SAM],
         dict(_std3("testBasic1gz", "testBasic1", gz=True),
              **{"gz:output/testBasic1gz.introns.bed6.gz": "expected/testBasic1.introns.bed6"})),
    case("testBasic2", "extract", ["--strandness=RF", CONF0] + _std3_args("testBasic2") + [BASIC_SAM],
         _std3("testBasic2")),
    case("testBasic3", "extract",
         [CONF0] + _std3_args("testBasic3") + ["--pass-through=output/testBasic3.bam", BASIC_BAM],
         _std3("testBasic3"), prep=PREP_BAM, sam_of_bam=[("output/testBasic3.bam", "output/testBasic3.sam")],
         same_as=[("output/testBasic3.sam", BASIC_SAM)], needs=["pass-through"]),
    case("testBasic4", "extract", ["--min-confidence-score=1.0"] + _std3_args("testBasic4") + [BASIC_BAM],
         _std3("testBasic4"), prep=PREP_BAM),
    case("testBasic5", "extract", ["--exclude=multi", CONF0] + _std3_args("testBasic5") + [BASIC_BAM],
         _std3("testBasic5"), prep=PREP_BAM),
    case("testMulti1", "extract", [CONF0] + _std3_args("testMulti1") + ["input/test_multi.sam"], _std3("testMulti1")),
    case("testFake10a", "extract",
         [CONF0, "--genome-fasta=" + FAKE10_FA] + _std3_args("testFake10a") +
         ["--pass-through=output/testFake10a.sam", "--debug-trace=output/testFake10a.debug.tsv", STAR_SAM],
         dict(_std3("testFake10a"), **{"output/testFake10a.debug.tsv": "expected/testFake10a.debug.tsv"}),
         prep=PREP_FA, same_as=[("output/testFake10a.sam", STAR_SAM)], needs=["pass-through", "debug-trace"]),
    case("testFake10agz", "extract",
         [CONF0, "--genome-fasta=" + FAKE10_FA] + _std3_args("testFake10agz", gz=True) +
         ["--debug-trace=output/testFake10agz.debug.tsv.gz", STAR_SAM],
         dict(_std3("testFake10agz", "testFake10a", gz=True),
              **{"gz:output/testFake10agz.debug.tsv.gz": "expected/testFake10a.debug.tsv"}),
         prep=PREP_FA, needs=["debug-trace"]),
    case("testFake10aConf", "extract",
         ["--genome-fasta=" + FAKE10_FA] + _std3_args("testFake10aConf") +
         ["--pass-through=output/testFake10aConf.sam", "--debug-trace=output/testFake10aConf.debug.tsv", STAR_SAM],
         dict(_std3("testFake10aConf"), **{"output/testFake10aConf.debug.tsv": "expected/testFake10aConf.debug.tsv"}),
         prep=PREP_FA, same_as=[("output/testFake10aConf.sam", STAR_SAM)], needs=["pass-through", "debug-trace"]),
    case("testFake10b", "extract",
         [CONF0, "--genome-fasta=" + FAKE10_FA] + _std3_args("testFake10b") +
         ["--pass-through=output/testFake10b.sam", HISAT_SAM],
         _std3("testFake10b"), prep=PREP_FA, same_as=[("output/testFake10b.sam", HISAT_SAM)], needs=["pass-through"]),
    case("testFake10c", "extract",
         [CONF0, "--genome-fasta=" + FAKE10_FA, "--skip-missing-targets"] + _std3_args("testFake10c") +
         ["input/fake10.missing.sam"],
         _std3("testFake10c"), prep=PREP_FA, stderr="expected/testFake10c.err"),
    case("testSjFilterDefault", "extract",
         [CONF0, "--genome-fasta=" + SCRAMBLE_FA] + _std3_args("testSjFilterDefault") + [STAR_SAM],
         _std3("testSjFilterDefault"), prep=PREP_SCR),
    case("testSjFilterCanon", "extract",
         [CONF0, "--sj-filter=canon", "--genome-fasta=" + SCRAMBLE_FA] + _std3_args("testSjFilterCanon") + [STAR_SAM],
         _std3("testSjFilterCanon"), prep=PREP_SCR),
    case("testSjFilterAll", "extract",
         [CONF0, "--sj-filter=all", "--genome-fasta=" + SCRAMBLE_FA] + _std3_args("testSjFilterAll") + [STAR_SAM],
         _std3("testSjFilterAll"), prep=PREP_SCR),
    # no golden in the reference (bedToBigBed only checks sortedness); the regenerated
    # expected files come from the compiled reference itself, see make_golden.py
    case("testSortBug", "extract",
         ["--junction-bed=output/testSortBug.juncs.bed", "--intron-bed=output/testSortBug.introns.bed",
          "input/sort-bug.sam"],
         {"output/testSortBug.juncs.bed": "expected/testSortBug.juncs.bed",
          "output/testSortBug.introns.bed": "expected/testSortBug.introns.bed"}),
    case("testOnt1", "extract",
         [CONF0] + _std3_args("testOnt1") + ["--pass-through=output/testOnt1.sam", ONT_SAM],
         _std3("testOnt1"), same_as=[("output/testOnt1.sam", ONT_SAM)], needs=["pass-through"]),
    case("testOntIndel", "extract", [CONF0, "--allow-anchor-indels"] + _std3_args("testOntIndel") + [ONT_SAM],
         _std3("testOntIndel")),
    case("testOntSkipTargets", "extract",
         [CONF0, "--allow-anchor-indels"] + _std3_args("testOntSkipTargets") + ["input/ont-skip-targets.sam"],
         _std3("testOntSkipTargets")),
    case("testOntOutOfOrderTargets", "extract", ["--allow-anchor-indels", "input/ont-out-of-order-targets.sam"],
         {}, stderr="expected/testOntOutOfOrderTargets.err", allow_fail=True),
    case("testMultiSam", "extract",
         [CONF0] + _std3_args("testMultiSam") + ["input/test_multi.one.sam", "input/test_multi.two.sam"],
         _std3("testMultiSam")),
    case("testOntUnsorted", "extract",
         ["--unsorted", "--allow-anchor-indels"] + _std3_args("testOntUnsorted") + ["input/ont-out-of-order-targets.sam"],
         _std3("testOntUnsorted")),
    # orphan goldens (no Makefile rule); commands recovered in SURVEY.md section 4
    case("testStrandTagAdd", "extract",
         ["--genome-fasta=" + FAKE10_FA, "--set-XS-strand-tag", "--set-TS-strand-tag",
          "--pass-through=output/testStrandTagAdd.sam", STAR_SAM],
         {"output/testStrandTagAdd.sam": "expected/testStrandTagAdd.sam"}, prep=PREP_FA,
         needs=["pass-through", "strand-tags"]),
    case("testStrandTagUpdate", "extract",
         ["--genome-fasta=" + FAKE10_FA, "--set-XS-strand-tag",
          "--pass-through=output/testStrandTagUpdate.sam", HISAT_SAM],
         {"output/testStrandTagUpdate.sam": "expected/testStrandTagUpdate.sam"}, prep=PREP_FA,
         needs=["pass-through", "strand-tags"]),
]


def _m3(name):
    return ["--junction-bed=output/%s.juncs.bed" % name, "--intron-bed=output/%s.introns.bed" % name,
            "--intron-calls=output/%s.introns.tsv" % name]


B134 = ["expected/testBasic1.introns.tsv", "expected/testBasic3.introns.tsv", "expected/testBasic4.introns.tsv"]
SC = "input/strand-confict/"

MERGE_CASES = [
    case("testBasic134Merge", "merge", _m3("testBasic134Merge") + B134, _std3("testBasic134Merge")),
    case("testBasic134MergeGz", "merge",
         ["--junction-bed=output/testBasic134MergeGz.juncs.bed.gz", "--intron-bed=output/testBasic134MergeGz.introns.bed.gz",
          "--intron-calls=output/testBasic134MergeGz.introns.tsv.gz",
          "output/testBasic1_gz.introns.tsv.gz", "output/testBasic3_gz.introns.tsv.gz", "output/testBasic4_gz.introns.tsv.gz"],
         _std3("testBasic134MergeGz", "testBasic134Merge", gz=True),
         prep=[("gzip", "expected/testBasic1.introns.tsv", "output/testBasic1_gz.introns.tsv.gz"),
               ("gzip", "expected/testBasic3.introns.tsv", "output/testBasic3_gz.introns.tsv.gz"),
               ("gzip", "expected/testBasic4.introns.tsv", "output/testBasic4_gz.introns.tsv.gz")]),
    case("testBasic134LstMerge", "merge", _m3("testBasic134LstMerge") + ["-i", "input/basic134.files.txt"],
         _std3("testBasic134LstMerge")),
    case("testFake10abMerge", "merge",
         ["--intron-calls=output/testFake10abMerge.introns.tsv", "expected/testFake10a.introns.tsv",
          "expected/testFake10b.introns.tsv"],
         {"output/testFake10abMerge.introns.tsv": "expected/testFake10abMerge.introns.tsv"}),
    case("testStrandConflictMerge", "merge",
         ["--intron-calls=output/testStrandConflictMerge.introns.tsv", SC + "SRR6225360.calls.tsv",
          SC + "SRR6225362.calls.tsv", SC + "SRR6225361.calls.tsv", SC + "SRR6225363.calls.tsv"],
         {"output/testStrandConflictMerge.introns.tsv": "expected/testStrandConflictMerge.introns.tsv"},
         stderr="expected/testStrandConflictMerge.err"),
    case("testStrand12Merge", "merge",
         ["--intron-calls=output/testStrand12Merge.introns.tsv", "input/merge1.calls.tsv", "input/merge2.calls.tsv"],
         {"output/testStrand12Merge.introns.tsv": "expected/testStrand12Merge.introns.tsv"}),
    case("testStrand21Merge", "merge",
         ["--intron-calls=output/testStrand21Merge.introns.tsv", "input/merge2.calls.tsv", "input/merge1.calls.tsv"],
         {"output/testStrand21Merge.introns.tsv": "expected/testStrand21Merge.introns.tsv"}),
    case("testSjFilterCanonMerge", "merge",
         ["--sj-filter=canon"] + _m3("testSjFilterCanonMerge") + ["expected/testSjFilterAll.introns.tsv"],
         _std3("testSjFilterCanonMerge")),
    case("testSjFilterAllMerge", "merge",
         ["--sj-filter=all"] + _m3("testSjFilterAllMerge") + ["expected/testSjFilterAll.introns.tsv"],
         _std3("testSjFilterAllMerge")),
    case("testSjFilterDefaultMerge", "merge", _m3("testSjFilterDefaultMerge") + ["expected/testSjFilterAll.introns.tsv"],
         _std3("testSjFilterDefaultMerge")),
]

ALL_CASES = EXTRACT_CASES + MERGE_CASES
