"""The branch-free automaton the kernels run (walk_op_flat, csrc/ipx_walk.cuh) against the branching restatement of
the reference's ReadJunctionExtractor (walk_op, same header; pinned to the oracle by the GPU parity tests): compiled
for the host and fuzzed on random CIGARs, including unknown and zero-length operations."""
import os
import subprocess

import helpers


def test_flat_automaton_equals_branching_automaton(tmp_path):
    exe = str(tmp_path / "walk_fuzz")
    src = os.path.join(helpers.REPO, "tests", "native", "walk_fuzz.cc")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, src], check=True)
    out = subprocess.run([exe, "1000000"], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.startswith("ok:")
