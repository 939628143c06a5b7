"""Shared helpers for the test-suite."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from intronprospector_b200 import _native  # noqa: E402

REF_DIR = os.path.join(REPO, "oracle", "_ref")
REF_EXTRACT = os.path.join(REF_DIR, "intronProspector")
REF_MERGE = os.path.join(REF_DIR, "intronProspectorMerge")
ORACLE_EXTRACT = REF_EXTRACT + ".oracle"
ORACLE_MERGE = REF_MERGE + ".oracle"
HTS_TOOL = os.path.join(REF_DIR, "hts_tool")
GPU_EXTRACT = _native.CLI_EXTRACT
GPU_MERGE = _native.CLI_MERGE

STD_OUTPUTS = ("introns.tsv", "juncs.bed", "introns.bed", "introns.bed6")


def run_extract(binary, workdir, tag, inputs, extra=()):
    """Run an intronProspector binary with all four outputs; returns (returncode, stderr, {suffix: bytes})."""
    out = {s: os.path.join(workdir, "%s.%s" % (tag, s)) for s in STD_OUTPUTS}
    cmd = [binary, "--intron-calls=" + out["introns.tsv"], "--junction-bed=" + out["juncs.bed"],
           "--intron-bed=" + out["introns.bed"], "--intron-bed6=" + out["introns.bed6"]] + list(extra) + list(inputs)
    proc = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    data = {}
    for s, p in out.items():
        data[s] = open(p, "rb").read() if os.path.exists(p) else None
    return proc.returncode, proc.stderr.decode(errors="replace"), data


def assert_same_outputs(a, b, what):
    for s in STD_OUTPUTS:
        assert a[s] is not None and b[s] is not None, "%s: missing output %s" % (what, s)
        if a[s] != b[s]:
            la, lb = a[s].split(b"\n"), b[s].split(b"\n")
            for i, (x, y) in enumerate(zip(la, lb)):
                if x != y:
                    raise AssertionError("%s: %s differs at line %d:\n  %r\n  %r" % (what, s, i + 1, x, y))
            raise AssertionError("%s: %s differs in length (%d vs %d lines)" % (what, s, len(la), len(lb)))
