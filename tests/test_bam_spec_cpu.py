"""Known answers for the BAM reader that both the product and the reference binary are built on (csrc/hts_compat): BAM
bytes assembled here from the SAM/BAM specification (SAMv1 section 4: the record layout, the auxiliary types, BGZF
blocks), with nothing of the repository involved in writing them -- several BGZF members of different sizes with records
straddling them, empty members in the middle, every integer width for NH, `B` arrays of every element type ahead of the
tags the packer looks for, `Z` / `H` / `f` values, a long CIGAR parked in `CG:B:I` behind the `<l_seq>S<len>N`
placeholder, an unmapped record without CIGAR and bases.  The records are also written down as SAM text, by hand: the
reader must turn the bytes into exactly that text, and the tools must produce the same calls from either form."""
import os
import struct
import subprocess
import zlib

import pytest

import helpers

SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
CIGAR_OPS = "MIDNSHP=XB"
REFS = [("chr1", 100000), ("chr2", 50000)]
HEADER_TEXT = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in REFS)


def reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def cigar_words(text):
    out, num = [], ""
    for ch in text:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | CIGAR_OPS.index(ch))
            num = ""
    return out


def ref_len(words):
    return sum(w >> 4 for w in words if (w & 15) in (0, 2, 3, 7, 8))


def aux_bytes(tag, typ, value):
    b = tag.encode()
    if typ in "cCsSiI":
        return b + typ.encode() + struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I"}[typ], value)
    if typ == "A":
        return b + b"A" + value.encode()
    if typ == "f":
        return b + b"f" + struct.pack("<f", value)
    if typ == "Z":
        return b + b"Z" + value.encode() + b"\0"
    if typ == "H":
        return b + b"H" + value.encode() + b"\0"
    if typ.startswith("B"):
        sub = typ[1]
        fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
        return b + b"B" + sub.encode() + struct.pack("<i", len(value)) + b"".join(struct.pack("<" + fmt, v) for v in value)
    raise ValueError(typ)


def record(name, flag, ref, pos0, mapq, cigar, seq, qual, aux, stored_cigar=None):
    """One alignment record, SAMv1 4.2.  stored_cigar: the words actually stored (the CG placeholder); `cigar` the real one."""
    real = cigar_words(cigar) if cigar != "*" else []
    stored = real if stored_cigar is None else cigar_words(stored_cigar)
    end = pos0 + (ref_len(real) if real and not flag & 4 else 1)
    l_seq = 0 if seq == "*" else len(seq)
    body = struct.pack("<iiBBHHHiiii", ref, pos0, len(name) + 1, mapq, reg2bin(pos0, end), len(stored), flag, l_seq, -1, -1, 0)
    body += name.encode() + b"\0"
    body += b"".join(struct.pack("<I", w) for w in stored)
    packed = bytearray((l_seq + 1) // 2)
    for i in range(l_seq):
        packed[i >> 1] |= SEQ_CODE[seq[i]] << (4 if i % 2 == 0 else 0)
    body += bytes(packed)
    body += bytes([0xFF] * l_seq) if qual == "*" else bytes(ord(c) - 33 for c in qual)
    body += b"".join(aux_bytes(*a) for a in aux)
    return struct.pack("<i", len(body)) + body


def bam_stream(records):
    text = HEADER_TEXT.encode()
    out = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(REFS))
    for name, length in REFS:
        out += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", length)
    return out + b"".join(records)


def bgzf_block(data):
    """One BGZF member, SAMv1 4.1: gzip with the BC extra field holding the member's size - 1."""
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    deflated = comp.compress(data) + comp.flush()
    bsize = len(deflated) + 25
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + deflated +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def bgzf(stream, cuts):
    """The stream cut at `cuts` (positions) into members; an empty member after every third one; the EOF marker last."""
    out, last = b"", 0
    for k, c in enumerate(list(cuts) + [len(stream)]):
        out += bgzf_block(stream[last:c])
        if k % 3 == 2:
            out += bgzf_block(b"")
        last = c
    return out + bgzf_block(b"")


BASES = "ACGTTGCAAGCTTCGA"
SEQ = lambda n: (BASES * (n // 16 + 1))[:n]  # noqa: E731
QUAL = lambda n: "".join(chr(33 + (7 * i) % 40) for i in range(n))  # noqa: E731

# (name, flag, ref, pos0, mapq, cigar, seq, qual, aux, stored_cigar, aux as SAM text)
RECORDS = [
    ("r1", 0, 0, 100, 60, "50M1000N50M", SEQ(100), QUAL(100), [("NH", "C", 1), ("XS", "A", "+")], None, "NH:i:1\tXS:A:+"),
    ("r2", 16, 0, 200, 3, "30M5000N40M2I30M", SEQ(102), QUAL(102),
     [("ZB", "Bc", [-1, 2, -3]), ("ZS", "BS", [1, 65535]), ("NH", "s", 2), ("XS", "A", "-"), ("ZF", "f", 1.5), ("ZZ", "Z", "hello world"),
      ("ZH", "H", "1AE301")], None, "ZB:B:c,-1,2,-3\tZS:B:S,1,65535\tNH:i:2\tXS:A:-\tZF:f:1.5\tZZ:Z:hello world\tZH:H:1AE301"),
    ("r3", 256, 0, 300, 0, "75M", SEQ(75), "*", [("NH", "I", 3)], None, "NH:i:3"),
    ("r4", 0, 0, 400, 255, "40M2000N30M4000N30M", SEQ(100), QUAL(100),
     [("ZI", "Bi", [-70000, 5]), ("CG", "BI", cigar_words("40M2000N30M4000N30M")), ("NH", "i", 1)], "100S6100N", "ZI:B:i,-70000,5\tNH:i:1"),
    ("r5", 0, 0, 420, 20, "60M700N40M", SEQ(100), QUAL(100), [("Zf", "Bf", [0.5, -2.0]), ("ZC", "BC", [0, 255]), ("ZU", "BI", [4000000000]), ("XS", "A", "?")],
     None, "Zf:B:f,0.5,-2\tZC:B:C,0,255\tZU:B:I,4000000000\tXS:A:?"),
    ("r6", 4, 1, 9, 0, "*", "*", "*", [("NH", "c", 0)], None, "NH:i:0"),
    ("r7", 1024, 1, 1000, 30, "20S40M300N40M", SEQ(100), QUAL(100), [("NH", "S", 1), ("XS", "A", "+"), ("ZL", "Z", "")], None, "NH:i:1\tXS:A:+\tZL:Z:"),
    ("r8", 16, 1, 1000, 30, "40M300N35M1D5M", SEQ(80), QUAL(80), [("XS", "A", "+"), ("NH", "C", 200)], None, "XS:A:+\tNH:i:200"),
]


def sam_text():
    lines = [HEADER_TEXT]
    for name, flag, ref, pos0, mapq, cigar, seq, qual, _aux, _stored, aux_text in RECORDS:
        lines.append("\t".join([name, str(flag), REFS[ref][0], str(pos0 + 1), str(mapq), cigar, "*", "0", "0", seq, qual, aux_text]) + "\n")
    return "".join(lines)


def raw_records():
    return [record(*r[:9], stored_cigar=r[9]) for r in RECORDS]


def write_inputs(tmp_path):
    stream = bam_stream(raw_records())
    # members of very different sizes; cuts inside the header, inside a record's fixed part, inside an aux array
    cuts = [3, 40, 41, 300, 301, 302, 700, 1100, 1101, 1500]
    cuts = [c for c in cuts if c < len(stream)]
    paths = {}
    for tag, data in (("many.bam", bgzf(stream, cuts)), ("one.bam", bgzf(stream, [])), ("raw.bam", stream)):
        paths[tag] = str(tmp_path / tag)
        open(paths[tag], "wb").write(data)
    paths["hand.sam"] = str(tmp_path / "hand.sam")
    open(paths["hand.sam"], "w").write(sam_text())
    return paths


def test_spec_built_bam_reads_back_as_the_sam_text_written_by_hand(built, tmp_path):
    paths = write_inputs(tmp_path)
    for tag in ("many.bam", "one.bam", "raw.bam"):
        out = str(tmp_path / (tag + ".sam"))
        subprocess.run([helpers.HTS_TOOL, "view-sam", paths[tag], out], check=True)
        assert open(out).read() == sam_text(), tag


@pytest.mark.parametrize("extra", [["-C", "0.0"], ["--allow-anchor-indels", "-C", "0.0"], ["--exclude=multi", "-C", "0.0"]], ids=["all", "indels", "no_multi"])
def test_spec_built_bam_and_hand_written_sam_give_the_same_calls(built, tmp_path, extra):
    """Through the reference binary and through the product's command-line tool (host side, linked against the CPU
    oracle here; tests/test_cli_gpu.py runs the same tool on the device): BAM in every framing == SAM."""
    paths = write_inputs(tmp_path)
    for binary in (helpers.REF_EXTRACT, helpers.ORACLE_EXTRACT):
        want = None
        for tag in ("hand.sam", "many.bam", "one.bam", "raw.bam"):
            rc, err, out = helpers.run_extract(binary, str(tmp_path), tag.replace(".", "_"), [paths[tag]], extra)
            assert rc == 0, err
            if want is None:
                want = out
                calls = out["introns.tsv"] if "introns.tsv" in out else list(out.values())[0]
                assert calls.count(b"\n") >= 4  # header + the introns of r1, r2, r4, r5 / r7 / r8
            else:
                assert out == want, (os.path.basename(binary), tag)
