"""CPU-side checks of the boundary: the CUDA library loads, exports every symbol include/ipx.h declares, and
refuses to compute without a GPU (no fallback).  The host I/O library decodes the reference's fixtures."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers
from intronprospector_b200 import _native


def declared(header):
    text = open(os.path.join(helpers.REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipx(?:io)?_[a-z0-9_]+)\s*\(", text)))


def test_libipx_exports_every_declared_symbol(built):
    lib = C.CDLL(_native.LIBIPX)
    names = declared("ipx.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libipx.so does not export %s" % n
    assert set(_native.ABI_SYMBOLS) <= set(names)


def test_libipx_io_exports_every_declared_symbol(built):
    lib = C.CDLL(_native.LIBIPX_IO)
    for n in declared("ipx_io.h"):
        assert hasattr(lib, n), "libipx_io.so does not export %s" % n


def test_no_cpu_fallback(built):
    """Without a device ipx_create must fail; with one this test is meaningless and is skipped."""
    lib = _native.load_library()
    if lib.ipx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_native.IpxError):
        _native.Engine()


def test_cli_fails_loudly_without_gpu(built, fixture_dir):
    lib = _native.load_library()
    if lib.ipx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    rc, err, _ = helpers.run_extract(helpers.GPU_EXTRACT, fixture_dir, "x", ["input/test_basic.sam"])
    assert rc == 1
    assert "no usable CUDA device" in err


def test_reader_decodes_fixture(built, fixture_dir):
    io = _native.load_io_library()
    r = C.c_void_p()
    assert io.ipxio_open(os.path.join(fixture_dir, "input", "test_basic.sam").encode(), C.byref(r)) == 0
    assert io.ipxio_n_targets(r) == 1
    assert io.ipxio_target_name(r, 0) == b"chr10"
    n, c = 1024, 65536
    tid, pos = np.empty(n, np.int32), np.empty(n, np.int32)
    flag = np.empty(n, np.uint16)
    nh, xs = np.empty(n, np.uint8), np.empty(n, np.uint8)
    off, cig = np.empty(n + 1, np.uint32), np.empty(c, np.uint32)
    ncig = C.c_uint64()
    got = io.ipxio_read_batch(r, n, c, *[a.ctypes.data for a in (tid, pos, flag, nh, xs, off, cig)], C.byref(ncig))
    io.ipxio_close(r)
    lines = [l for l in open(os.path.join(fixture_dir, "input", "test_basic.sam")) if not l.startswith("@")]
    assert got == len(lines)
    f = lines[0].split("\t")
    assert flag[0] == int(f[1]) and pos[0] == int(f[3]) - 1 and tid[0] == 0
    # first record: 72M131352N29M, NH:i:2, no XS
    assert list(cig[off[0]:off[1]]) == [(72 << 4) | 0, (131352 << 4) | 3, (29 << 4) | 0]
    assert nh[0] == 2 and xs[0] == 0
    assert off[got] == ncig.value
