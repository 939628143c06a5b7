"""Merge path on CPU: the product's intronProspectorMerge host code (TSV parsing, CLI, formatters) linked against the
CPU oracle reproduces the unmodified reference binary on synthetic per-sample tables."""
import os
import subprocess

import pytest

import helpers
import merge_synth


@pytest.mark.parametrize("extra", [[], ["--sj-filter=canon"]], ids=["all", "canon"])
def test_oracle_merge_cli_matches_reference(built, tmp_path, extra):
    paths, _ = merge_synth.write_tables(str(tmp_path), n_samples=10, n_universe=8_000)
    outs = {}
    for tag, binary in (("ref", helpers.REF_MERGE), ("ora", helpers.ORACLE_MERGE)):
        o = {s: str(tmp_path / ("%s.%s" % (tag, s))) for s in ("introns.tsv", "juncs.bed", "introns.bed")}
        proc = subprocess.run([binary, "--intron-calls=" + o["introns.tsv"], "--junction-bed=" + o["juncs.bed"],
                               "--intron-bed=" + o["introns.bed"]] + extra + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert proc.returncode == 0, proc.stderr
        outs[tag] = (proc.stderr, {s: open(p, "rb").read() for s, p in o.items()})
    assert outs["ref"][0] == outs["ora"][0] and b"different strand" in outs["ref"][0]
    for s in outs["ref"][1]:
        assert outs["ref"][1][s] == outs["ora"][1][s], s
