#!/usr/bin/env python3
"""Builds tests/golden/reference_fixtures.tar.gz -- run in the build container only.

The archive holds TEST DATA only: the reference's test inputs (tests/input) and golden outputs
(tests/expected), plus expected files for the one reference test that has no golden (testSortBug),
generated here by the unmodified reference compiled under oracle/_ref.  Before writing the archive
every case in cases.py is run with that reference binary and compared with the reference's own
expected files, which is what pins oracle/_ref (and, through it, hts_compat) to the reference.

    make -C oracle && python tests/golden/make_golden.py
"""
import hashlib
import io
import os
import shutil
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import runner  # noqa: E402

REF_TESTS = "/root/reference/tests"
GENERATED = {"testSortBug": ["expected/testSortBug.juncs.bed", "expected/testSortBug.introns.bed"]}


def main():
    bins = runner.ref_binaries()
    work = tempfile.mkdtemp(prefix="ipx_golden_")
    for sub in ("input", "expected"):
        shutil.copytree(os.path.join(REF_TESTS, sub), os.path.join(work, sub),
                        ignore=lambda d, names: [n for n in names if "..." in n or n == "*...(1)"])
    os.makedirs(os.path.join(work, "output"))
    # cases without a reference golden: generate expected from the reference binary first
    for c in cases.ALL_CASES:
        if c["name"] in GENERATED:
            tmp = dict(c, outputs={})
            probs = runner.run_case(tmp, work, bins)
            assert not probs, probs
            for got, want in c["outputs"].items():
                shutil.copyfile(os.path.join(work, got), os.path.join(work, want))
    failed = 0
    for c in cases.ALL_CASES:
        probs = runner.run_case(c, work, bins)
        print("%-28s %s" % (c["name"], "ok" if not probs else "FAIL"))
        for p in probs:
            failed += 1
            print(p)
    if failed:
        sys.exit("reference binary does not reproduce the reference goldens; archive not written")
    manifest = []
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz", format=tarfile.PAX_FORMAT) as tf:
        for sub in ("input", "expected"):
            for root, _dirs, files in sorted(os.walk(os.path.join(work, sub))):
                for fn in sorted(files):
                    full = os.path.join(root, fn)
                    rel = os.path.relpath(full, work)
                    with open(full, "rb") as fh:
                        data = fh.read()
                    manifest.append("%s  %s" % (hashlib.sha256(data).hexdigest(), rel))
                    ti = tarfile.TarInfo(rel)
                    ti.size = len(data)
                    ti.mtime = 0
                    ti.mode = 0o644
                    tf.addfile(ti, io.BytesIO(data))
    with open(runner.FIXTURE_TAR, "wb") as fh:
        fh.write(buf.getvalue())
    with open(os.path.join(HERE, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    print("wrote %s (%d files, %d bytes)" % (runner.FIXTURE_TAR, len(manifest), len(buf.getvalue())))
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
