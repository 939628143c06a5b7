"""oracle/_ref (the unmodified reference compiled against hts_compat) must reproduce every golden
of the reference's own test-suite (reference tests/Makefile.in:67-254).  This pins the oracle binary
and, through the pass-through identity comparisons, the SAM/BAM/FASTA layer itself."""
import os

import pytest

import cases
import runner


@pytest.mark.parametrize("case", cases.ALL_CASES, ids=[c["name"] for c in cases.ALL_CASES])
def test_reference_binary_reproduces_golden(case, built, fixture_dir):
    bins = runner.ref_binaries()
    for k, p in bins.items():
        assert os.path.exists(p), "missing %s (run make -C oracle)" % p
    problems = runner.run_case(case, fixture_dir, bins)
    assert not problems, "\n".join(problems)
