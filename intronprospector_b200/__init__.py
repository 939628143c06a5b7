"""intronprospector_b200 -- B200-native junction extraction behind the intronProspector CLI.

The product is native: `csrc/build/libipx.so` (sm_100a kernels + C ABI, include/ipx.h), `csrc/build/libipx_io.so`
(SAM/BAM/FASTA/TSV host I/O, include/ipx_io.h) and the two drop-in command line tools next to them.  This Python
package is only the thin ctypes binding used by tests and bench.py; it never computes anything itself and it
raises when the CUDA library is missing or finds no GPU.
"""
from ._native import (  # noqa: F401
    CLI_EXTRACT,
    CLI_MERGE,
    IpxError,
    Engine,
    Table,
    load_io_library,
    load_library,
)
