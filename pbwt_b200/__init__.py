"""pbwt_b200 -- B200 (sm_100a) PBWT construction and match engine behind a C ABI.

The product is ``pbwt_b200/libpbwtgpu.so`` (CUDA kernels + include/pbwtgpu.h); this package is
the thin ctypes binding used by tests and bench.py.  No CPU fallback exists.
"""
from .api import (Engine, Panel, Shard, QueryShard, PbwtGpuError, Result, KERNEL_CLASSES, MATCH_DTYPE, NOMATCH_DTYPE, WANT_D,  # noqa: F401
                  WANT_MAXWITHIN, WANT_YZ, WANT_LONGWITHIN, WANT_STATS, LIB_PATH, load_library, words_per_col, expand_groups)
