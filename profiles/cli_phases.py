"""Phases of the extraction tool (IPX_TRACE) on a BGZF BAM of aligner-style records.  Usage: python profiles/cli_phases.py [records]"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
d = tempfile.mkdtemp(dir="/dev/shm")
syn = Synth(n, kind=SHORT_UNSTRANDED, seed=20261021, model_seed=20261018)
bam = os.path.join(d, "a.bam")
syn.write_full_bam(bam, threads=0, level=1)
syn.close()
tool = os.path.join("intronprospector_b200", "csrc", "build", "intronProspector")
for i in range(3):
    t0 = time.perf_counter()
    p = subprocess.run([tool, "--intron-calls=" + d + "/o.tsv", "--junction-bed=" + d + "/o.bed", "--intron-bed=" + d + "/o9.bed", bam],
                       env=dict(os.environ, IPX_TRACE="1"), stderr=subprocess.PIPE, text=True)
    print("run %d: %.2f s" % (i, time.perf_counter() - t0))
    print("".join(l + "\n" for l in p.stderr.splitlines() if l.startswith("[ipx cli]") or "flush:" in l))
import shutil; shutil.rmtree(d)
