"""Host I/O of the drop-in CLIs (include/ipx_io.h) on CPU: the genome loader reads what faidx_fetch_seq returns for
whole sequences, whatever the number of loader threads."""
import ctypes as C

import pytest

import helpers  # noqa: F401
from intronprospector_b200 import _native
from intronprospector_b200.synth import SHORT_UNSTRANDED, Synth


def _load(fa, threads, monkeypatch):
    monkeypatch.setenv("IPX_GENOME_THREADS", str(threads))
    io = _native.load_io_library()
    g = C.c_void_p()
    assert io.ipxio_genome_load(fa.encode(), C.byref(g)) == 0
    n = io.ipxio_genome_n_seqs(g)
    out = [(io.ipxio_genome_names(g)[i].decode(), C.string_at(io.ipxio_genome_seqs(g)[i], int(io.ipxio_genome_lengths(g)[i])))
           for i in range(n)]
    io.ipxio_genome_free(g)
    return out


def test_genome_loader_threads_and_content(built, tmp_path, monkeypatch):
    syn = Synth(1000, kind=SHORT_UNSTRANDED, genome_len=30_000_000, n_genes=50, n_contigs=11)
    fa = str(tmp_path / "g.fa")
    syn.write_genome(fa)
    syn.close()
    # what the file says, parsed the slow way
    want, name, parts = [], None, []
    for line in open(fa, "rb"):
        if line.startswith(b">"):
            if name is not None:
                want.append((name, b"".join(parts)))
            name, parts = line[1:].split()[0].decode(), []
        else:
            parts.append(bytes(c for c in line if 33 <= c <= 126))
    want.append((name, b"".join(parts)))
    one = _load(fa, 1, monkeypatch)
    assert [n for n, _ in one] == [n for n, _ in want]
    assert all(a[1] == b[1] for a, b in zip(one, want))
    assert sum(len(s) for _, s in one) > 25_000_000
    for threads in (3, 8):
        assert _load(fa, threads, monkeypatch) == one


def _extract(binary, workdir, tag, source, stdin_file=None, env=None):
    """intron calls of `source` (a path, or None = standard input fed from stdin_file through a pipe)"""
    import os
    import subprocess
    out = os.path.join(workdir, tag + ".tsv")
    cmd = [binary, "--intron-calls=" + out, "-C", "0.0"] + ([source] if source else [])
    e = dict(os.environ)
    e.update(env or {})
    if stdin_file is None:
        proc = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    else:
        feeder = subprocess.Popen(["cat", stdin_file], stdout=subprocess.PIPE)  # a real pipe: not seekable
        proc = subprocess.run(cmd, cwd=workdir, stdin=feeder.stdout, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
        feeder.stdout.close()
        feeder.wait()
    assert proc.returncode == 0, proc.stderr
    return open(out, "rb").read()


@pytest.mark.parametrize("binary", ["ORACLE_EXTRACT", "REF_EXTRACT"])
def test_piped_input_of_every_kind(built, tmp_path, binary):
    """SAM, gzipped SAM, BGZF BAM and uncompressed BAM through a pipe on standard input give what the files give --
    with the BGZF decoder pool (the default) and with one decoding thread (the gzip stream reader)."""
    import gzip
    import subprocess
    binary = getattr(helpers, binary)
    tool = helpers.HTS_TOOL if hasattr(helpers, "HTS_TOOL") else helpers.REF_EXTRACT.replace("intronProspector", "hts_tool")
    syn = Synth(60_000, kind=SHORT_UNSTRANDED, genome_len=50_000_000, n_genes=400, n_contigs=4)
    raw = str(tmp_path / "raw.bam")
    syn.write_raw_bam(raw)
    syn.close()
    bam, sam, samgz = str(tmp_path / "in.bam"), str(tmp_path / "in.sam"), str(tmp_path / "in.sam.gz")
    subprocess.run([tool, "view-bam", raw, bam], check=True)
    subprocess.run([tool, "view-sam", raw, sam], check=True)
    with open(sam, "rb") as fi, gzip.open(samgz, "wb") as fo:
        fo.write(fi.read())
    want = _extract(binary, str(tmp_path), "file", bam)
    assert want.count(b"\n") > 200
    n = 0
    for path in (bam, raw, sam, samgz):
        for threads in ("4", "1"):
            n += 1
            got = _extract(binary, str(tmp_path), "pipe%d" % n, None, stdin_file=path, env={"IPX_DECODE_THREADS": threads})
            assert got == want, (path, threads)
            got = _extract(binary, str(tmp_path), "devstdin%d" % n, "/dev/stdin", stdin_file=path, env={"IPX_DECODE_THREADS": threads})
            assert got == want, (path, threads)


@pytest.mark.parametrize("kind", ["short", "long"])
def test_records_packed_in_place_equal_records_through_bam1_t(built, tmp_path, kind):
    """BAM records are packed where they lie in the reader's buffer (the default when nothing needs the record
    afterwards); IPX_NO_RAW_PACK=1 goes through sam_read1 and a bam1_t.  Same junction calls either way, sorted mode
    (with its per-target loop and its read-ahead) and --unsorted."""
    import subprocess
    from intronprospector_b200.synth import LONG
    syn = Synth(5_000 if kind == "long" else 80_000, kind=LONG if kind == "long" else SHORT_UNSTRANDED, genome_len=50_000_000,
                n_genes=400, n_contigs=4)
    raw, bam = str(tmp_path / "raw.bam"), str(tmp_path / "in.bam")
    syn.write_raw_bam(raw)
    syn.close()
    subprocess.run([helpers.HTS_TOOL, "view-bam", raw, bam], check=True)
    for extra in ([], ["--unsorted"]):
        outs = []
        for env in ({}, {"IPX_NO_RAW_PACK": "1"}):
            import os
            out = str(tmp_path / "o.tsv")
            e = dict(os.environ)
            e.update(env)
            proc = subprocess.run([helpers.ORACLE_EXTRACT, "--intron-calls=" + out, "-C", "0.0", "--allow-anchor-indels"] + extra + [bam],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
            assert proc.returncode == 0, proc.stderr
            outs.append(open(out, "rb").read())
        assert outs[0] == outs[1] and outs[0].count(b"\n") > 100


def test_aligner_style_tags_in_place_and_through_bam1_t(built, tmp_path):
    """Records that carry a dozen tags the way aligners write them (integers of several widths, strings, arrays, XS on
    spliced reads only, NH in the middle): the in-place packer, the bam1_t packer and the unmodified reference agree."""
    import os
    import subprocess
    import numpy as np
    syn = Synth(40_000, kind=SHORT_UNSTRANDED, genome_len=50_000_000, n_genes=300, n_contigs=3)
    raw, sam, sam2, bam = (str(tmp_path / n) for n in ("raw.bam", "plain.sam", "tagged.sam", "tagged.bam"))
    syn.write_raw_bam(raw)
    syn.close()
    subprocess.run([helpers.HTS_TOOL, "view-sam", raw, sam], check=True)
    rng = np.random.default_rng(3)
    with open(sam) as fi, open(sam2, "w") as fo:
        for line in fi:
            if line.startswith("@"):
                fo.write(line)
                continue
            f = line.rstrip("\n").split("\t")
            core, tags = f[:11], [t for t in f[11:] if not t.startswith(("NH:", "XS:"))]
            spliced = "N" in core[5]
            nh = int(rng.choice([1, 1, 1, 2, 5, 300, 70000]))
            new = ["HI:i:1", "AS:i:%d" % rng.integers(0, 300), "nM:i:%d" % rng.integers(0, 5), "NH:i:%d" % nh,
                   "MD:Z:%d" % rng.integers(1, 150), "jM:B:c,%d" % (-1 if not spliced else 1), "jI:B:i,%d,%d" % (10, 20)]
            if spliced and rng.random() < 0.8:
                new.append("XS:A:%s" % rng.choice(["+", "-"]))
            new.append("RG:Z:sample1")
            fo.write("\t".join(core + tags + new) + "\n")
    subprocess.run([helpers.HTS_TOOL, "view-bam", sam2, bam], check=True)
    outs = {}
    for tag, binary, env in (("ref", helpers.REF_EXTRACT, {}), ("raw", helpers.ORACLE_EXTRACT, {}),
                             ("bam1_t", helpers.ORACLE_EXTRACT, {"IPX_NO_RAW_PACK": "1"})):
        e = dict(os.environ)
        e.update(env)
        out = str(tmp_path / (tag + ".tsv"))
        proc = subprocess.run([binary, "--intron-calls=" + out, "-C", "0.0", "--exclude=multi", bam], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, env=e)
        assert proc.returncode == 0, proc.stderr
        outs[tag] = open(out, "rb").read()
    assert outs["ref"].count(b"\n") > 100
    assert outs["raw"] == outs["ref"]
    assert outs["bam1_t"] == outs["ref"]


def test_reference_inputs_as_bam_through_both_packers(built, tmp_path):
    """Every SAM input of the reference's own test suite, converted to BAM: sorted mode (with its per-target loop,
    out-of-order and skipped targets, the sortedness error) and --unsorted, records packed in place and through bam1_t
    -- return code, messages and all outputs as the unmodified reference gives them."""
    import os
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(helpers.REPO, "tests", "golden"))
    import runner
    work = runner.unpack_fixtures(str(tmp_path / "fx"))
    sams = sorted(f for f in os.listdir(os.path.join(work, "input")) if f.endswith(".sam"))
    assert len(sams) >= 8
    checked = 0
    for name in sams:
        bam = str(tmp_path / (name[:-4] + ".bam"))
        if subprocess.run([helpers.HTS_TOOL, "view-bam", os.path.join(work, "input", name), bam]).returncode != 0:
            continue  # an input the writer cannot represent
        for extra in ([], ["--unsorted"]):
            runs = {}
            for tag, binary, env in (("ref", helpers.REF_EXTRACT, {}), ("in_place", helpers.ORACLE_EXTRACT, {}),
                                     ("bam1_t", helpers.ORACLE_EXTRACT, {"IPX_NO_RAW_PACK": "1"})):
                e = dict(os.environ)
                e.update(env)
                d = tmp_path / ("%s.%s.%d" % (name, tag, len(extra)))
                d.mkdir()
                out = {s: str(d / s) for s in ("introns.tsv", "juncs.bed", "introns.bed")}
                proc = subprocess.run([binary, "--intron-calls=" + out["introns.tsv"], "--junction-bed=" + out["juncs.bed"],
                                       "--intron-bed=" + out["introns.bed"], "-C", "0.0"] + extra + [bam],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
                runs[tag] = (proc.returncode, proc.stderr,
                             {s: (open(p, "rb").read() if os.path.exists(p) else None) for s, p in out.items()})
            for tag in ("in_place", "bam1_t"):
                assert runs[tag] == runs["ref"], (name, extra, tag)
            checked += 1
    assert checked >= 16


def _read_columns(path, pipeline, threads=3, max_records=2_000_000, max_cigar=8_000_000):
    """every record of a BAM file as columns: through the BAM pipeline or record by record"""
    import ctypes as C
    import numpy as np
    io = _native.load_io_library()
    cols = {"tid": np.empty(max_records, np.int32), "pos": np.empty(max_records, np.int32), "flag": np.empty(max_records, np.uint16),
            "nh_class": np.empty(max_records, np.uint8), "xs": np.empty(max_records, np.uint8),
            "cigar_off": np.empty(max_records + 1, np.uint32), "cigar": np.empty(max_cigar, np.uint32)}
    ptrs = [cols[k].ctypes.data for k in ("tid", "pos", "flag", "nh_class", "xs", "cigar_off", "cigar")]
    nc = C.c_uint64()
    if pipeline:
        n = io.ipxio_pipeline_read(path.encode(), threads, max_records, max_cigar, *ptrs, C.byref(nc))
    else:
        r = C.c_void_p()
        assert io.ipxio_open(path.encode(), C.byref(r)) == 0
        n = io.ipxio_read_batch(r, max_records, max_cigar, *ptrs, C.byref(nc))
        io.ipxio_close(r)
    if n < 0:
        return n, None
    out = {k: cols[k][:n].copy() for k in ("tid", "pos", "flag", "nh_class", "xs")}
    out["cigar_off"] = cols["cigar_off"][:n + 1].copy()
    out["cigar"] = cols["cigar"][:nc.value].copy()
    return n, out


def _same_columns(a, b):
    import numpy as np
    for k in ("tid", "pos", "flag", "cigar_off", "cigar"):
        assert np.array_equal(a[k], b[k]), k
    # the pipeline looks at the tags of reads with an N operation only (nothing else on the path reads them)
    off = a["cigar_off"].astype(np.int64)
    has_n = np.zeros(len(a["tid"]), bool)
    is_n = (a["cigar"] & 15) == 3
    idx = np.flatnonzero(is_n)
    has_n[np.searchsorted(off, idx, side="right") - 1] = True
    for k in ("nh_class", "xs"):
        assert np.array_equal(a[k][has_n], b[k][has_n]), k


@pytest.mark.parametrize("kind", ["short", "long"])
def test_bam_pipeline_packs_what_the_record_reader_reads(built, tmp_path, kind):
    """The multi-threaded BAM pipeline (chunks of BGZF blocks inflated and packed in parallel, records that straddle
    chunks gathered by the framer) against sam_read1 + the bam1_t packer: aligner-style records in BGZF blocks, and the
    same stream uncompressed; 1, 3 and 8 threads."""
    from intronprospector_b200.synth import LONG
    syn = Synth(6_000 if kind == "long" else 150_000, kind=LONG if kind == "long" else SHORT_UNSTRANDED, genome_len=50_000_000,
                n_genes=400, n_contigs=4)
    full, raw = str(tmp_path / "full.bam"), str(tmp_path / "raw.bam")
    syn.write_full_bam(full, threads=2, level=1)
    syn.write_raw_bam(raw)
    n_want = syn.n
    syn.close()
    for path in (full, raw):
        n0, want = _read_columns(path, pipeline=False)
        assert n0 == n_want
        for threads in (1, 3, 8):
            n1, got = _read_columns(path, pipeline=True, threads=threads)
            assert n1 == n_want
            _same_columns(want, got)


def test_bam_pipeline_reports_damage_after_the_records_before_it(built, tmp_path):
    """a file cut in the middle of a BGZF block, in the middle of a record, and one with a corrupted block: the
    pipeline gives up where the record reader gives up"""
    syn = Synth(120_000, kind=SHORT_UNSTRANDED, genome_len=50_000_000, n_genes=400, n_contigs=4)
    full, raw = str(tmp_path / "full.bam"), str(tmp_path / "raw.bam")
    syn.write_full_bam(full, threads=2, level=1)
    syn.write_raw_bam(raw)
    syn.close()
    data, rawdata = open(full, "rb").read(), open(raw, "rb").read()
    cut1, cut2, bad = str(tmp_path / "cut1.bam"), str(tmp_path / "cut2.bam"), str(tmp_path / "bad.bam")
    open(cut1, "wb").write(data[:len(data) * 2 // 3])
    open(cut2, "wb").write(rawdata[:len(rawdata) * 2 // 3 + 7])
    spoiled = bytearray(data)
    for i in range(len(data) // 2, len(data) // 2 + 64):
        spoiled[i] ^= 0x5A
    open(bad, "wb").write(bytes(spoiled))
    for path in (cut1, cut2, bad):
        n0, _ = _read_columns(path, pipeline=False)
        n1, _ = _read_columns(path, pipeline=True)
        assert n0 == -1 and n1 == -1, path


def test_command_line_tool_through_the_pipeline_and_record_by_record(built, tmp_path):
    """the tool's two ways of reading a BAM file give the same calls: sorted (per-target selection on runs of equal
    target index), --unsorted, two files, a file whose targets come out of order (same error text, same partial output)"""
    import os
    import subprocess
    syn = Synth(90_000, kind=SHORT_UNSTRANDED, genome_len=50_000_000, n_genes=400, n_contigs=5)
    bam = str(tmp_path / "in.bam")
    syn.write_full_bam(bam, threads=2, level=1)
    cols = syn.columns()
    # the same records with the third and fourth target swapped in the stream: "not coordinate sorted"
    import numpy as np
    tid = cols["tid"]
    order = np.concatenate([np.flatnonzero(tid == t) for t in (0, 1, 3, 2, 4)])
    syn.close()
    shuffled = str(tmp_path / "shuffled.bam")
    _write_raw_bam_from_columns(shuffled, ["chr%d" % (i + 1) for i in range(5)], cols, order)

    def run(args, env):
        out = {k: str(tmp_path / ("o." + k)) for k in ("tsv", "bed")}
        e = dict(os.environ)
        e.update(env)
        proc = subprocess.run([helpers.ORACLE_EXTRACT, "--intron-calls=" + out["tsv"], "--junction-bed=" + out["bed"]] + args,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
        return proc.returncode, proc.stderr, open(out["tsv"], "rb").read(), open(out["bed"], "rb").read()

    for args in ([bam], ["--unsorted", bam], [bam, bam], [shuffled], ["-u", shuffled]):
        fast = run(args, {})
        slow = run(args, {"IPX_NO_PIPELINE": "1"})
        ref = subprocess.run([helpers.REF_EXTRACT, "--intron-calls=" + str(tmp_path / "r.tsv"), "--junction-bed=" + str(tmp_path / "r.bed")] + args,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert fast == slow, args
        assert fast[0] == ref.returncode and fast[1] == ref.stderr, args
        assert fast[2] == open(str(tmp_path / "r.tsv"), "rb").read() and fast[3] == open(str(tmp_path / "r.bed"), "rb").read(), args
    assert run([shuffled], {})[0] != 0


def _write_raw_bam_from_columns(path, names, cols, order):
    """uncompressed BAM with the records of `cols` in the given order (test fixture)"""
    import struct
    out = bytearray(b"BAM\x01")
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:100000000\n" % n for n in names)
    out += struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(names))
    for n in names:
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", 100000000)
    off = cols["cigar_off"]
    for r in order:
        cig = cols["cigar"][int(off[r]):int(off[r + 1])]
        aux = b""
        if cols["nh_class"][r]:
            aux += b"NHC" + bytes([1 if cols["nh_class"][r] == 1 else 3])
        if cols["xs"][r]:
            aux += b"XSA" + bytes([int(cols["xs"][r])])
        body = struct.pack("<iiIIiiii", int(cols["tid"][r]), int(cols["pos"][r]), (4680 << 16) | (255 << 8) | 2,
                           (int(cols["flag"][r]) << 16) | len(cig), 0, -1, -1, 0) + b"r\0" + cig.astype("<u4").tobytes() + aux
        out += struct.pack("<i", len(body)) + body
    open(path, "wb").write(bytes(out))


def _bgzip(src, dst, block=65280):
    """bgzip + its .gzi index, from the BGZF description in SAMv1 4.1 (members of at most 64 KB, BC extra field, EOF marker)."""
    import struct
    import test_bam_spec_cpu as spec
    data = open(src, "rb").read()
    out, index, coff = b"", [], 0
    for uoff in range(0, len(data), block):
        if uoff:
            index.append((coff, uoff))
        member = spec.bgzf_block(data[uoff:uoff + block])
        out += member
        coff += len(member)
    open(dst, "wb").write(out + spec.bgzf_block(b""))
    open(dst + ".gzi", "wb").write(struct.pack("<Q", len(index)) + b"".join(struct.pack("<QQ", c, u) for c, u in index))


def test_bgzip_compressed_genome(built, tmp_path, monkeypatch):
    """docs/intronProspector.md:151 of the reference recommends a bgzip'ed genome (fa.gz + .fai + .gzi): the loader reads
    the same sequences from it, the tools make the same calls, and what htslib refuses (plain gzip, no .gzi) is refused."""
    import gzip
    import os
    import shutil
    import subprocess
    syn = Synth(200_000, kind=SHORT_UNSTRANDED, genome_len=20_000_000, n_genes=300, n_contigs=7, seed=4)
    fa = str(tmp_path / "g.fa")
    syn.write_genome(fa)
    bam = str(tmp_path / "in.bam")
    syn.write_raw_bam(bam)
    syn.close()
    gz = fa + ".gz"
    _bgzip(fa, gz)
    shutil.copy(fa + ".fai", gz + ".fai")
    plain = _load(fa, 4, monkeypatch)
    for threads in (1, 5):
        assert _load(gz, threads, monkeypatch) == plain
    for binary in (helpers.REF_EXTRACT, helpers.ORACLE_EXTRACT):
        rc_a, err_a, a = helpers.run_extract(binary, str(tmp_path), "plain", [bam], ["-C", "0.0", "--genome-fasta=" + fa])
        rc_b, err_b, b = helpers.run_extract(binary, str(tmp_path), "bgz", [bam], ["-C", "0.0", "--genome-fasta=" + gz])
        assert rc_a == 0 and rc_b == 0, (err_a, err_b)
        helpers.assert_same_outputs(a, b, os.path.basename(binary))
        assert a["introns.tsv"].count(b"GT/AG") > 100
    # no .gzi beside it; a gzip file that is not BGZF
    os.rename(gz + ".gzi", gz + ".gzi.away")
    rc, err, _ = helpers.run_extract(helpers.ORACLE_EXTRACT, str(tmp_path), "nogzi", [bam], ["--genome-fasta=" + gz])
    assert rc != 0 and "can't open genome FASTA file or associated index" in err
    os.rename(gz + ".gzi.away", gz + ".gzi")
    with open(fa, "rb") as f, gzip.open(str(tmp_path / "p.fa.gz"), "wb") as g:
        g.write(f.read())
    shutil.copy(fa + ".fai", str(tmp_path / "p.fa.gz.fai"))
    open(str(tmp_path / "p.fa.gz.gzi"), "wb").write(b"\0" * 8)
    rc, err, _ = helpers.run_extract(helpers.ORACLE_EXTRACT, str(tmp_path), "gzip", [bam], ["--genome-fasta=" + str(tmp_path / "p.fa.gz")])
    assert rc != 0 and "please use bgzip" in err
