"""Runs one case of golden/cases.py with a given pair of CLI binaries and reports differences.

Used three ways: make_golden.py (reference binary vs. the reference's own expected files, in the
build container), tests/test_ref_goldens.py (oracle/_ref binaries vs. the committed fixtures) and
tests/test_cli_gpu.py (this repo's CUDA-backed CLIs vs. the same fixtures, on a B200).
"""
import gzip
import os
import shutil
import subprocess
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
FIXTURE_TAR = os.path.join(HERE, "reference_fixtures.tar.gz")
REF_DIR = os.path.join(REPO, "oracle", "_ref")


def ref_binaries():
    return {"extract": os.path.join(REF_DIR, "intronProspector"),
            "merge": os.path.join(REF_DIR, "intronProspectorMerge"),
            "hts_tool": os.path.join(REF_DIR, "hts_tool")}


def unpack_fixtures(dest):
    """Extract the committed fixture archive (input/ + expected/) under dest and return dest."""
    with tarfile.open(FIXTURE_TAR, "r:gz") as tf:
        tf.extractall(dest, filter="data")
    os.makedirs(os.path.join(dest, "output"), exist_ok=True)
    return dest


def _prep(step, workdir, bins):
    kind, src, dst = step
    src_p, dst_p = os.path.join(workdir, src), os.path.join(workdir, dst)
    os.makedirs(os.path.dirname(dst_p), exist_ok=True)
    if kind == "bam":
        subprocess.run([bins["hts_tool"], "view-bam", src_p, dst_p], check=True)
    elif kind == "fa":
        shutil.copyfile(src_p, dst_p)
        subprocess.run([bins["hts_tool"], "faidx", dst_p], check=True)
    elif kind == "gzip":
        with open(src_p, "rb") as fi, gzip.open(dst_p, "wb") as fo:
            fo.write(fi.read())
    else:
        raise ValueError(kind)


def _read(workdir, spec):
    if spec.startswith("gz:"):
        with gzip.open(os.path.join(workdir, spec[3:]), "rb") as fh:
            return fh.read()
    with open(os.path.join(workdir, spec), "rb") as fh:
        return fh.read()


def run_case(case, workdir, bins, skip_needs=()):
    """Run one case in workdir (already holding input/ and expected/). Returns a list of problem strings."""
    problems = []
    for step in case["prep"]:
        _prep(step, workdir, bins)
    args = list(case["args"])
    outputs = dict(case["outputs"])
    same_as = list(case["same_as"])
    sam_of_bam = list(case["sam_of_bam"])
    if any(n in skip_needs for n in case["needs"]):
        # drop the optional side outputs this binary does not implement yet
        def keep(a):
            if "pass-through" in skip_needs and a.startswith("--pass-through="):
                return False
            if "debug-trace" in skip_needs and a.startswith("--debug-trace="):
                return False
            return True
        args = [a for a in args if keep(a)]
        if "pass-through" in skip_needs:
            same_as, sam_of_bam = [], []
        if "debug-trace" in skip_needs:
            outputs = {k: v for k, v in outputs.items() if "debug.tsv" not in k}
    proc = subprocess.run([bins[case["tool"]]] + args, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if proc.returncode != 0 and not case["allow_fail"]:
        problems.append("%s: exit %d: %s" % (case["name"], proc.returncode, proc.stderr.decode(errors="replace")[-2000:]))
        return problems
    if case["stderr"] is not None:
        want = _read(workdir, case["stderr"])
        got = proc.stdout + proc.stderr
        if got != want:
            problems.append("%s: stderr differs:\n--- want\n%s\n--- got\n%s" % (case["name"], want.decode(), got.decode()))
    for bam, sam in sam_of_bam:
        subprocess.run([bins["hts_tool"], "view-sam", os.path.join(workdir, bam), os.path.join(workdir, sam)], check=True)
    for got_spec, want_spec in outputs.items():
        try:
            got = _read(workdir, got_spec)
        except OSError as ex:
            problems.append("%s: missing output %s (%s)" % (case["name"], got_spec, ex))
            continue
        want = _read(workdir, want_spec)
        if got != want:
            problems.append("%s: %s differs from %s\n--- want\n%s\n--- got\n%s" % (
                case["name"], got_spec, want_spec, want.decode(errors="replace")[:3000], got.decode(errors="replace")[:3000]))
    for got_spec, want_spec in same_as:
        if _read(workdir, got_spec) != _read(workdir, want_spec):
            problems.append("%s: %s is not identical to %s" % (case["name"], got_spec, want_spec))
    return problems
