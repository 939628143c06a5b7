"""The plain-C oracle (oracle/ipx_oracle.c) must reproduce the reference's goldens.

The product's CLI sources are linked against the oracle instead of the CUDA library (oracle/ipx_oracle_abi.c,
test-only) and run over the reference's whole golden matrix.  This pins the oracle AND checks the CLIs' host
logic (options, run modes, record selection, text formats) on a machine without a GPU.  The same matrix runs
against the CUDA-backed CLIs in tests/test_cli_gpu.py."""
import os

import pytest

import cases
import runner

# every side output of the reference is implemented (debug trace, pass-through, strand tags: SURVEY.md 8f rows 2-3)
NOT_YET = ()


def oracle_binaries():
    b = runner.ref_binaries()
    b["extract"] += ".oracle"
    b["merge"] += ".oracle"
    return b


@pytest.mark.parametrize("case", cases.ALL_CASES, ids=[c["name"] for c in cases.ALL_CASES])
def test_oracle_cli_reproduces_golden(case, built, fixture_dir):
    bins = oracle_binaries()
    for p in bins.values():
        assert os.path.exists(p), "missing %s (run make -C oracle)" % p
    problems = runner.run_case(case, fixture_dir, bins, skip_needs=NOT_YET)
    assert not problems, "\n".join(problems)
