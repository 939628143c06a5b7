"""Config 1 of BASELINE.json on the GPU: the reference's whole golden matrix through the CUDA-backed CLIs, plus
synthetic BAMs where the CUDA CLI's files must be byte-identical to the unmodified reference binary's."""
import os

import pytest

import cases
import helpers
import runner
from intronprospector_b200.synth import LONG, SHORT_STRANDED, SHORT_UNSTRANDED, Synth

pytestmark = pytest.mark.gpu
NOT_YET = ()


def gpu_binaries():
    b = runner.ref_binaries()
    b["extract"] = helpers.GPU_EXTRACT
    b["merge"] = helpers.GPU_MERGE
    return b


@pytest.mark.parametrize("case", cases.ALL_CASES, ids=[c["name"] for c in cases.ALL_CASES])
def test_cuda_cli_reproduces_golden(case, built, fixture_dir):
    problems = runner.run_case(case, fixture_dir, gpu_binaries(), skip_needs=NOT_YET)
    assert not problems, "\n".join(problems)


SYN = [
    ("short", SHORT_UNSTRANDED, 1_000_000, [], False),
    ("short_20m", SHORT_UNSTRANDED, 20_000_000, [], False),  # bench-scale widths: several pinned batches, millions of events

    ("short_conf0_multi", SHORT_UNSTRANDED, 300_000, ["-C", "0.0", "--exclude=multi", "-a", "12"], False),
    ("stranded_canon", SHORT_STRANDED, 500_000, ["--strandness=RF", "--min-anchor-length=12"], True),
    ("stranded_all", SHORT_STRANDED, 300_000, ["--strandness=RF", "--sj-filter=all", "-C", "0.0"], True),
    ("unstranded_genome", SHORT_UNSTRANDED, 300_000, ["-C", "0.0"], True),
    ("long_indels", LONG, 20_000, ["--allow-anchor-indels", "-C", "0.0"], False),
    ("long_default", LONG, 20_000, [], False),
    ("long_unsorted", LONG, 8_000, ["--unsorted", "--allow-anchor-indels", "-C", "0.0"], False),
]


@pytest.mark.parametrize("name,kind,n,args,genome", SYN, ids=[c[0] for c in SYN])
def test_cuda_cli_matches_reference_binary(name, kind, n, args, genome, built, tmp_path):
    syn = Synth(n, kind=kind, genome_len=100_000_000, n_genes=1000, n_contigs=5, seed=11 + len(name))
    bam = str(tmp_path / "in.bam")
    syn.write_raw_bam(bam)
    extra = list(args)
    if genome:
        fa = str(tmp_path / "g.fa")
        syn.write_genome(fa)
        extra.append("--genome-fasta=" + fa)
    rc_r, err_r, ref = helpers.run_extract(helpers.REF_EXTRACT, str(tmp_path), "ref", [bam], extra)
    rc_g, err_g, gpu = helpers.run_extract(helpers.GPU_EXTRACT, str(tmp_path), "gpu", [bam], extra)
    assert rc_r == 0, err_r
    assert rc_g == 0, err_g
    helpers.assert_same_outputs(ref, gpu, name)


def test_cuda_cli_on_spec_built_bam(built, tmp_path):
    """The BAM assembled from the specification in tests/test_bam_spec_cpu.py (BGZF members of odd sizes, B arrays, a
    CG:B:I long CIGAR) through the CUDA tool -- pipeline and record reader -- against the reference on the SAM text."""
    import test_bam_spec_cpu as spec
    paths = spec.write_inputs(tmp_path)
    extra = ["--allow-anchor-indels", "-C", "0.0"]
    rc_r, err_r, ref = helpers.run_extract(helpers.REF_EXTRACT, str(tmp_path), "ref", [paths["hand.sam"]], extra)
    assert rc_r == 0, err_r
    for tag in ("many.bam", "one.bam", "raw.bam", "hand.sam"):
        rc_g, err_g, gpu = helpers.run_extract(helpers.GPU_EXTRACT, str(tmp_path), "gpu_" + tag.replace(".", "_"), [paths[tag]], extra)
        assert rc_g == 0, err_g
        helpers.assert_same_outputs(ref, gpu, tag)
