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
