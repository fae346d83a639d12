"""ctypes binding of include/pbwtgpu.h (the C ABI of libpbwtgpu.so).

Python here is plumbing for tests and bench.py only: every call goes straight into the CUDA
library.  There is no CPU fallback -- if the library is missing or no sm_100 device is present
the constructors raise.

Host-side mirror of the reference interface for this path (names as in pbwt.h):
  * ``Engine.build``          <- readers' loop + pbwtCursorWriteForwards  (pbwtIO.c:572-574)
  * ``Engine.max_within``     <- matchMaximalWithin(p, report)            (pbwtMatch.c:115-142)
  * ``Engine.build_reverse``  <- pbwtBuildReverse(p)                      (pbwtCore.c:151-191)
  * ``Engine.long_within``    <- pbwtLongMatches(p, L), L > 0             (pbwtMatch.c:85-113,148-183)
  * ``Engine.match_sweep``    <- matchSequencesSweep(p, q, report)        (pbwtMatch.c:363-443)
  * ``Engine.pack3/unpack3``  <- pack3 / unpack3                          (pbwtCore.c:254-305)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PBWTGPU_LIB") or os.path.join(_HERE, "libpbwtgpu.so")  # (the override is a development aid: A/B builds)

WANT_YZ, WANT_D, WANT_MAXWITHIN, WANT_LONGWITHIN, WANT_STATS = 1, 2, 4, 8, 16
KERNEL_CLASSES = ("sweep", "pack3", "unpack3", "query", "synth", "aux", "prepass", "combine", "copy")

MATCH_DTYPE = np.dtype([("ai", "<i4"), ("bi", "<i4"), ("start", "<i4"), ("end", "<i4")])
NOMATCH_DTYPE = np.dtype([("query", "<i4"), ("allele", "<i4"), ("site", "<i4")])

_vp = C.c_void_p
_i64 = C.c_int64


class PbwtGpuError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load libpbwtgpu.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PbwtGpuError(f"{LIB_PATH} not built: run __graft_entry__.build() (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.pbwtgpu_strerror.restype = C.c_char_p
        lib.pbwtgpu_last_error.restype = C.c_char_p
        lib.pbwtgpu_last_error.argtypes = [_vp]
        lib.pbwtgpu_get_stream.restype = _vp
        lib.pbwtgpu_get_stream.argtypes = [_vp]
        lib.pbwtgpu_set_stream.argtypes = [_vp, _vp]
        lib.pbwtgpu_set_option.argtypes = [_vp, C.c_char_p, C.c_longlong]
        lib.pbwtgpu_init.argtypes = [C.c_int, C.POINTER(_vp)]
        lib.pbwtgpu_destroy.argtypes = [_vp]
        lib.pbwtgpu_destroy.restype = None
        lib.pbwtgpu_synchronize.argtypes = [_vp]
        lib.pbwtgpu_profile.argtypes = [_vp, C.c_int]
        lib.pbwtgpu_last_timings.argtypes = [_vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        lib.pbwtgpu_panel_from_host.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)]
        lib.pbwtgpu_panel_synthetic.argtypes = [_vp, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(_vp)]
        lib.pbwtgpu_panel_synthetic_range.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(_vp)]
        lib.pbwtgpu_shard_aggregate_ints.restype = C.c_size_t
        lib.pbwtgpu_shard_aggregate_ints.argtypes = [C.c_int]
        lib.pbwtgpu_shard_prepare.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_uint, C.POINTER(_vp)]
        lib.pbwtgpu_shard_export.argtypes = [_vp, _vp, _vp]
        lib.pbwtgpu_shard_finish.argtypes = [_vp, _vp, _vp, C.c_int, C.c_uint, C.POINTER(_vp)]
        lib.pbwtgpu_shard_free.argtypes = [_vp, _vp]
        lib.pbwtgpu_shard_free.restype = None
        lib.pbwtgpu_panel_to_host.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp]
        lib.pbwtgpu_panel_free.argtypes = [_vp, _vp]
        lib.pbwtgpu_panel_free.restype = None
        lib.pbwtgpu_build_device.argtypes = [_vp, _vp, _vp, C.c_uint, C.POINTER(_vp)]
        lib.pbwtgpu_sweep_yz_device.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, C.c_uint, C.POINTER(_vp)]
        lib.pbwtgpu_result_sizes.argtypes = [_vp, C.POINTER(_i64), C.POINTER(_i64)]
        lib.pbwtgpu_result_fetch.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
        lib.pbwtgpu_result_free.argtypes = [_vp, _vp]
        lib.pbwtgpu_result_free.restype = None
        lib.pbwtgpu_build.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_uint, C.POINTER(_vp), C.POINTER(_i64), _vp, _vp,
                                      C.POINTER(_vp), C.POINTER(_i64)]
        lib.pbwtgpu_max_within.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_i64)]
        lib.pbwtgpu_result_group.argtypes = [_vp, _vp, C.POINTER(_i64), C.POINTER(_i64)]
        lib.pbwtgpu_result_fetch_groups.argtypes = [_vp, _vp, _vp]
        lib.pbwtgpu_build_grouped.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_uint, C.POINTER(_vp), C.POINTER(_i64), _vp, _vp,
                                              C.POINTER(_vp), C.POINTER(_i64)]
        lib.pbwtgpu_max_within_grouped.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_i64)]
        lib.pbwtgpu_max_within_stats.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp]
        lib.pbwtgpu_result_fetch_hist.argtypes = [_vp, _vp, _vp]
        lib.pbwtgpu_cursor_at.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp]
        lib.pbwtgpu_build_reverse.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, C.POINTER(_vp), C.POINTER(_i64), _vp]
        lib.pbwtgpu_long_within.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, C.POINTER(_vp), C.POINTER(_i64)]
        lib.pbwtgpu_match_sweep.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _i64, C.c_int, _vp, C.POINTER(_vp),
                                            C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_i64)]
        lib.pbwtgpu_qshard_open.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _i64, C.c_int, _vp, C.c_int, C.c_int, C.POINTER(_vp)]
        lib.pbwtgpu_qshard_export_queries.argtypes = [_vp, _vp, _vp]
        lib.pbwtgpu_qshard_prepare.argtypes = [_vp, _vp, _vp, C.c_int, _vp]
        lib.pbwtgpu_qshard_chain.argtypes = [_vp, _vp, _vp, C.c_int, _vp, _vp]
        lib.pbwtgpu_qshard_finish.argtypes = [_vp, _vp, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_i64)]
        lib.pbwtgpu_qshard_free.argtypes = [_vp, _vp]
        lib.pbwtgpu_qshard_free.restype = None
        lib.pbwtgpu_pack3.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_i64)]
        lib.pbwtgpu_unpack3.argtypes = [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp]
        lib.pbwtgpu_free.argtypes = [_vp]
        lib.pbwtgpu_free.restype = None
        lib.pbwtgpu_host_alloc.argtypes = [C.c_size_t]
        lib.pbwtgpu_host_alloc.restype = _vp
        lib.pbwtgpu_host_trim.argtypes = []
        lib.pbwtgpu_host_trim.restype = None
        _lib = lib
    return _lib


def words_per_col(M: int) -> int:
    return 4 * ((M + 127) // 128)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _take(lib, p: _vp, nbytes: int, dtype) -> np.ndarray:
    """Copy a malloc()'d buffer returned by the library into numpy and free it."""
    out = np.frombuffer(C.string_at(p, nbytes), dtype=dtype).copy() if nbytes else np.zeros(0, dtype)
    lib.pbwtgpu_free(p)
    return out


class Panel:
    def __init__(self, eng: "Engine", handle: _vp, M: int, N: int):
        self.eng, self.handle, self.M, self.N = eng, handle, M, N

    def to_host(self, first: int = 0, nsites: int | None = None) -> np.ndarray:
        n = self.N - first if nsites is None else nsites
        out = np.zeros((n, words_per_col(self.M)), np.uint32)
        self.eng._check(self.eng.lib.pbwtgpu_panel_to_host(self.eng.ctx, self.handle, first, n, _ptr(out)))
        return out

    def free(self):
        if self.handle:
            self.eng.lib.pbwtgpu_panel_free(self.eng.ctx, self.handle)
            self.handle = None


class Result:
    def __init__(self, eng: "Engine", handle: _vp, M: int, N: int, flags: int):
        self.eng, self.handle, self.M, self.N, self.flags = eng, handle, M, N, flags
        nyz, nm = _i64(0), _i64(0)
        eng.lib.pbwtgpu_result_sizes(handle, C.byref(nyz), C.byref(nm))
        self.nyz, self.nmatch = nyz.value, nm.value

    def fetch(self, yz=True, a_end=True, d_end=True, matches=True, out=None):
        """D2H. ``out`` may hold preallocated (e.g. pinned) numpy buffers keyed yz/a_end/d_end/matches."""
        out = out or {}
        has_d = bool(self.flags & (WANT_D | WANT_MAXWITHIN | WANT_LONGWITHIN))
        def buf(key, want, make):
            if not want:
                return None
            return out[key] if key in out else make()

        bufs = {
            "yz": buf("yz", yz and (self.flags & WANT_YZ), lambda: np.zeros(max(self.nyz, 1), np.uint8)),
            "a_end": buf("a_end", a_end, lambda: np.zeros(self.M, np.int32)),
            "d_end": buf("d_end", d_end and has_d, lambda: np.zeros(self.M + 1, np.int32)),
            "matches": buf("matches", matches and (self.flags & (WANT_MAXWITHIN | WANT_LONGWITHIN)), lambda: np.zeros(max(self.nmatch, 1), MATCH_DTYPE)),
        }
        self.eng._check(self.eng.lib.pbwtgpu_result_fetch(self.eng.ctx, self.handle, _ptr(bufs["yz"]), _ptr(bufs["a_end"]),
                                                          _ptr(bufs["d_end"]), _ptr(bufs["matches"])))
        if bufs["yz"] is not None:
            bufs["yz"] = bufs["yz"][: self.nyz]
        if bufs["matches"] is not None:
            bufs["matches"] = bufs["matches"][: self.nmatch]
        return bufs

    def fetch_groups(self) -> np.ndarray:
        """The match records as a grouped int32 stream ({ai, start, end, n} + n x bi per run), built on the device."""
        nw, ng = _i64(0), _i64(0)
        self.eng._check(self.eng.lib.pbwtgpu_result_group(self.eng.ctx, self.handle, C.byref(nw), C.byref(ng)))
        out = np.zeros(max(nw.value, 1), np.int32)
        self.eng._check(self.eng.lib.pbwtgpu_result_fetch_groups(self.eng.ctx, self.handle, _ptr(out)))
        return out[: nw.value]

    def fetch_hist(self) -> np.ndarray:
        """WANT_STATS: the -stats length histogram accumulated on the device, hist[0 .. N]."""
        hist = np.zeros(self.N + 1, np.int64)
        self.eng._check(self.eng.lib.pbwtgpu_result_fetch_hist(self.eng.ctx, self.handle, _ptr(hist)))
        return hist

    def free(self):
        if self.handle:
            self.eng.lib.pbwtgpu_result_free(self.eng.ctx, self.handle)
            self.handle = None


class Shard:
    """One GPU's site range of a longer panel (see include/pbwtgpu.h, multi-GPU section)."""

    def __init__(self, eng: "Engine", handle: _vp, M: int, N: int):
        self.eng, self.handle, self.M, self.N = eng, handle, M, N

    def export(self, dev_ptr: int):
        """Copy this shard's aggregate node into device memory at dev_ptr (e.g. a torch tensor's data_ptr())."""
        self.eng._check(self.eng.lib.pbwtgpu_shard_export(self.eng.ctx, self.handle, dev_ptr))

    def finish(self, dev_aggs_ptr, rank: int, flags: int) -> Result:
        h = _vp()
        self.eng._check(self.eng.lib.pbwtgpu_shard_finish(self.eng.ctx, self.handle, dev_aggs_ptr, rank, flags, C.byref(h)))
        return Result(self.eng, h, self.M, self.N, flags)

    def free(self):
        if self.handle:
            self.eng.lib.pbwtgpu_shard_free(self.eng.ctx, self.handle)
            self.handle = None


class QueryShard:
    """One GPU's site range of a multi-GPU matchSequencesSweep (include/pbwtgpu.h, pbwtgpu_qshard_*)."""

    def __init__(self, eng: "Engine", handle: _vp):
        self.eng, self.handle = eng, handle

    def export_queries(self, dev_ptr: int):
        self.eng._check(self.eng.lib.pbwtgpu_qshard_export_queries(self.eng.ctx, self.handle, dev_ptr))

    def prepare(self, dev_query_aggs_ptr, rank: int, dev_panel_agg_ptr: int):
        self.eng._check(self.eng.lib.pbwtgpu_qshard_prepare(self.eng.ctx, self.handle, dev_query_aggs_ptr, rank, dev_panel_agg_ptr))

    def chain(self, dev_panel_aggs_ptr, rank: int, dev_state_in_ptr, dev_state_out_ptr: int):
        self.eng._check(self.eng.lib.pbwtgpu_qshard_chain(self.eng.ctx, self.handle, dev_panel_aggs_ptr, rank, dev_state_in_ptr,
                                                          dev_state_out_ptr))

    def finish(self):
        mt, nm, nom, nn = _vp(), _i64(0), _vp(), _i64(0)
        self.eng._check(self.eng.lib.pbwtgpu_qshard_finish(self.eng.ctx, self.handle, C.byref(mt), C.byref(nm), C.byref(nom), C.byref(nn)))
        return _take(self.eng.lib, mt, nm.value * 16, MATCH_DTYPE), _take(self.eng.lib, nom, nn.value * 12, NOMATCH_DTYPE)

    def free(self):
        if self.handle:
            self.eng.lib.pbwtgpu_qshard_free(self.eng.ctx, self.handle)
            self.handle = None


class Engine:
    """One per GPU.  Raises PbwtGpuError if the CUDA library or an sm_100 device is unavailable."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        ctx = _vp()
        rc = self.lib.pbwtgpu_init(device, C.byref(ctx))
        if rc != 0:
            raise PbwtGpuError(f"pbwtgpu_init({device}) failed: {self.lib.pbwtgpu_strerror(rc).decode()}")
        self.ctx = ctx
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pbwtgpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise PbwtGpuError(f"{self.lib.pbwtgpu_strerror(rc).decode()}: {self.lib.pbwtgpu_last_error(self.ctx).decode()}")

    # ---- plumbing ------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self.lib.pbwtgpu_set_stream(self.ctx, cuda_stream_ptr))

    def get_stream(self) -> int:
        return self.lib.pbwtgpu_get_stream(self.ctx)

    def synchronize(self):
        self._check(self.lib.pbwtgpu_synchronize(self.ctx))

    def set_option(self, name: str, value: int):
        self._check(self.lib.pbwtgpu_set_option(self.ctx, name.encode(), value))

    def profile(self, enable: bool):
        self._check(self.lib.pbwtgpu_profile(self.ctx, int(enable)))

    def last_timings(self):
        n = len(KERNEL_CLASSES)
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        self._check(self.lib.pbwtgpu_last_timings(self.ctx, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(KERNEL_CLASSES)}

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Page-locked host buffer from the library's pool (pbwtgpu_host_alloc) as a uint8 array; give it back with
        host_free(array).  What the host shim assembles its batches of bit columns in."""
        p = self.lib.pbwtgpu_host_alloc(nbytes)
        if not p:
            raise PbwtGpuError(f"pbwtgpu_host_alloc({nbytes}) failed")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def host_free(self, arr: np.ndarray):
        self.lib.pbwtgpu_free(arr.ctypes.data_as(_vp))

    # ---- panels --------------------------------------------------------------------------
    def panel_from_host(self, xbits: np.ndarray, M: int) -> Panel:
        xbits = np.ascontiguousarray(xbits, np.uint32)
        N = xbits.shape[0] if xbits.ndim == 2 else xbits.size // words_per_col(M)
        h = _vp()
        self._check(self.lib.pbwtgpu_panel_from_host(self.ctx, _ptr(xbits), M, N, C.byref(h)))
        self.synchronize()
        return Panel(self, h, M, N)

    def panel_synthetic(self, M: int, N: int, seed: int, L: int = 2000, mu: float = 2.0, first_site: int = 0) -> Panel:
        thresh = min(int(mu / M * 2**32), 2**32 - 1)
        h = _vp()
        self._check(self.lib.pbwtgpu_panel_synthetic_range(self.ctx, M, N, first_site, C.c_uint64(seed), L, thresh, C.byref(h)))
        return Panel(self, h, M, N)

    # ---- multi-GPU site shards -----------------------------------------------------------
    def shard_aggregate_ints(self, M: int) -> int:
        return int(self.lib.pbwtgpu_shard_aggregate_ints(M))

    def shard_prepare(self, panel: Panel, site_offset: int, total_sites: int, flags: int, a_start=None) -> "Shard":
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        h = _vp()
        self._check(self.lib.pbwtgpu_shard_prepare(self.ctx, panel.handle, _ptr(a0), site_offset, total_sites, flags, C.byref(h)))
        return Shard(self, h, panel.M, panel.N)

    # ---- device-resident sweeps ----------------------------------------------------------
    def build_device(self, panel: Panel, flags: int, a_start=None) -> Result:
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        h = _vp()
        self._check(self.lib.pbwtgpu_build_device(self.ctx, panel.handle, _ptr(a0), flags, C.byref(h)))
        return Result(self, h, panel.M, panel.N, flags)

    def sweep_yz_device(self, yz: np.ndarray, M: int, N: int, flags: int, a_start=None) -> Result:
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        h = _vp()
        self._check(self.lib.pbwtgpu_sweep_yz_device(self.ctx, _ptr(yz), yz.size, M, N, _ptr(a0), flags, C.byref(h)))
        return Result(self, h, M, N, flags | WANT_D)

    # ---- one-shot host API (the drop-in boundary) ----------------------------------------
    def build(self, xbits: np.ndarray, M: int, N: int, a_start=None, with_d=True, with_matches=False, with_yz=True, long_L=0):
        """-> dict(yz, a_end, d_end, matches) exactly as the reference would leave them.
        long_L > 0: matches are those of -longWithin long_L instead of -maxWithin."""
        xbits = np.ascontiguousarray(xbits, np.uint32)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        flags = (WANT_YZ if with_yz else 0) | (WANT_D if with_d else 0) | (WANT_MAXWITHIN if with_matches and not long_L else 0)
        if long_L:
            flags |= WANT_LONGWITHIN
            with_matches = True
            self.set_option("long_within", long_L)
        yz, nyz, mt, nm = _vp(), _i64(0), _vp(), _i64(0)
        a_end = np.zeros(M, np.int32)
        d_end = np.zeros(M + 1, np.int32) if (with_d or with_matches) else None
        self._check(self.lib.pbwtgpu_build(self.ctx, _ptr(xbits), M, N, _ptr(a0), flags, C.byref(yz), C.byref(nyz),
                                           _ptr(a_end), _ptr(d_end), C.byref(mt), C.byref(nm)))
        return {
            "yz": _take(self.lib, yz, nyz.value, np.uint8) if with_yz else None,
            "a_end": a_end,
            "d_end": d_end,
            "matches": _take(self.lib, mt, nm.value * 16, MATCH_DTYPE) if with_matches else None,
        }

    def build_grouped_raw(self, xbits: np.ndarray, M: int, N: int, flags: int, a_end: np.ndarray, d_end: np.ndarray | None):
        """pbwtgpu_build_grouped: as build_raw, the matches as a grouped stream -> (yz_ptr, nyz, groups_ptr, nwords)."""
        yz, nyz, gp, nw = _vp(), _i64(0), _vp(), _i64(0)
        self._check(self.lib.pbwtgpu_build_grouped(self.ctx, _ptr(xbits), M, N, None, flags, C.byref(yz), C.byref(nyz), _ptr(a_end),
                                                   _ptr(d_end), C.byref(gp), C.byref(nw)))
        return yz, nyz.value, gp, nw.value

    def max_within_grouped_raw(self, yz_ptr, nyz: int, M: int, N: int):
        gp, nw = _vp(), _i64(0)
        self._check(self.lib.pbwtgpu_max_within_grouped(self.ctx, yz_ptr, nyz, M, N, None, None, None, C.byref(gp), C.byref(nw)))
        return gp, nw.value

    def max_within_grouped(self, yz: np.ndarray, M: int, N: int, a_start=None) -> np.ndarray:
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        gp, nw = _vp(), _i64(0)
        self._check(self.lib.pbwtgpu_max_within_grouped(self.ctx, _ptr(yz), yz.size, M, N, _ptr(a0), None, None, C.byref(gp), C.byref(nw)))
        return _take(self.lib, gp, nw.value * 4, np.int32)

    def build_raw(self, xbits: np.ndarray, M: int, N: int, flags: int, a_end: np.ndarray, d_end: np.ndarray | None):
        """pbwtgpu_build exactly as the C host shim calls it: host columns in, library-owned host buffers out.
        -> (yz_ptr, nyz, matches_ptr, nmatch); the caller reads them in place and returns them with free_raw()."""
        yz, nyz, mt, nm = _vp(), _i64(0), _vp(), _i64(0)
        self._check(self.lib.pbwtgpu_build(self.ctx, _ptr(xbits), M, N, None, flags, C.byref(yz), C.byref(nyz), _ptr(a_end),
                                           _ptr(d_end), C.byref(mt), C.byref(nm)))
        return yz, nyz.value, mt, nm.value

    def max_within_raw(self, yz_ptr, nyz: int, M: int, N: int):
        """pbwtgpu_max_within on a host stream given by pointer -> (matches_ptr, nmatch); free_raw() afterwards."""
        mt, nm = _vp(), _i64(0)
        self._check(self.lib.pbwtgpu_max_within(self.ctx, yz_ptr, nyz, M, N, None, None, None, C.byref(mt), C.byref(nm)))
        return mt, nm.value

    def free_raw(self, *ptrs):
        for p in ptrs:
            if p:
                self.lib.pbwtgpu_free(p)

    def max_within(self, yz: np.ndarray, M: int, N: int, a_start=None, with_matches=True):
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        mt, nm = _vp(), _i64(0)
        a_end, d_end = np.zeros(M, np.int32), np.zeros(M + 1, np.int32)
        self._check(self.lib.pbwtgpu_max_within(self.ctx, _ptr(yz), yz.size, M, N, _ptr(a0), _ptr(a_end), _ptr(d_end),
                                                C.byref(mt) if with_matches else None, C.byref(nm)))
        return {"a_end": a_end, "d_end": d_end,
                "matches": _take(self.lib, mt, nm.value * 16, MATCH_DTYPE) if with_matches else None}

    def max_within_stats(self, yz: np.ndarray, M: int, N: int, a_start=None) -> np.ndarray:
        """pbwtLongMatches(p, 0) under -stats: the reporters' length histogram hist[0 .. N], matches consumed on the device."""
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        hist = np.zeros(N + 1, np.int64)
        self._check(self.lib.pbwtgpu_max_within_stats(self.ctx, _ptr(yz), yz.size, M, N, _ptr(a0), _ptr(hist)))
        return hist

    def cursor_at(self, yz: np.ndarray, M: int, N: int, k: int, a_start=None):
        """State of a PbwtCursor entering site k of a stored stream -> (y or None, a, d, c or None)."""
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        y, a, d, c = np.zeros(M, np.uint8), np.zeros(M, np.int32), np.zeros(M + 1, np.int32), np.zeros(1, np.int32)
        self._check(self.lib.pbwtgpu_cursor_at(self.ctx, _ptr(yz), yz.size, M, N, _ptr(a0), k, _ptr(y), _ptr(a), _ptr(d), _ptr(c)))
        return (y, a, d, int(c[0])) if k < N else (None, a, d, None)

    def build_reverse(self, yz: np.ndarray, M: int, N: int, a_start=None):
        """pbwtBuildReverse (pbwtCore.c:151-191) -> (aFend, zz, aRend)."""
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        zz, nzz = _vp(), _i64(0)
        a_fend, a_rend = np.zeros(M, np.int32), np.zeros(M, np.int32)
        self._check(self.lib.pbwtgpu_build_reverse(self.ctx, _ptr(yz), yz.size, M, N, _ptr(a0), _ptr(a_fend), C.byref(zz),
                                                   C.byref(nzz), _ptr(a_rend)))
        return a_fend, _take(self.lib, zz, nzz.value, np.uint8), a_rend

    def long_within(self, yz: np.ndarray, M: int, N: int, L: int, a_start=None):
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        mt, nm = _vp(), _i64(0)
        a_end, d_end = np.zeros(M, np.int32), np.zeros(M + 1, np.int32)
        self._check(self.lib.pbwtgpu_long_within(self.ctx, _ptr(yz), yz.size, M, N, _ptr(a0), L, _ptr(a_end), _ptr(d_end),
                                                 C.byref(mt), C.byref(nm)))
        return {"a_end": a_end, "d_end": d_end, "matches": _take(self.lib, mt, nm.value * 16, MATCH_DTYPE)}

    def match_sweep(self, pyz, M, N, qyz, Q, pa_start=None, qa_start=None):
        pyz = np.ascontiguousarray(pyz, np.uint8)
        qyz = np.ascontiguousarray(qyz, np.uint8)
        pa = None if pa_start is None else np.ascontiguousarray(pa_start, np.int32)
        qa = None if qa_start is None else np.ascontiguousarray(qa_start, np.int32)
        mt, nm, nom, nn = _vp(), _i64(0), _vp(), _i64(0)
        self._check(self.lib.pbwtgpu_match_sweep(self.ctx, _ptr(pyz), pyz.size, M, N, _ptr(pa), _ptr(qyz), qyz.size, Q,
                                                 _ptr(qa), C.byref(mt), C.byref(nm), C.byref(nom), C.byref(nn)))
        return _take(self.lib, mt, nm.value * 16, MATCH_DTYPE), _take(self.lib, nom, nn.value * 12, NOMATCH_DTYPE)

    def query_shard(self, pyz, M, N, qyz, Q, first_site: int, nsites: int, pa_start=None, qa_start=None) -> QueryShard:
        pyz = np.ascontiguousarray(pyz, np.uint8)
        qyz = np.ascontiguousarray(qyz, np.uint8)
        pa = None if pa_start is None else np.ascontiguousarray(pa_start, np.int32)
        qa = None if qa_start is None else np.ascontiguousarray(qa_start, np.int32)
        h = _vp()
        self._check(self.lib.pbwtgpu_qshard_open(self.ctx, _ptr(pyz), pyz.size, M, N, _ptr(pa), _ptr(qyz), qyz.size, Q, _ptr(qa),
                                                 first_site, nsites, C.byref(h)))
        return QueryShard(self, h)

    def pack3(self, y: np.ndarray) -> np.ndarray:
        """y: uint8 [N, M] sorted columns -> concatenated pack3 stream."""
        y = np.ascontiguousarray(y, np.uint8)
        N, M = y.shape
        yz, nyz = _vp(), _i64(0)
        self._check(self.lib.pbwtgpu_pack3(self.ctx, _ptr(y), M, N, C.byref(yz), C.byref(nyz)))
        return _take(self.lib, yz, nyz.value, np.uint8)

    def unpack3(self, yz: np.ndarray, M: int, N: int):
        yz = np.ascontiguousarray(yz, np.uint8)
        y = np.zeros((max(N, 1), M), np.uint8)
        zc = np.zeros(max(N, 1), np.int32)
        self._check(self.lib.pbwtgpu_unpack3(self.ctx, _ptr(yz), yz.size, M, N, _ptr(y), _ptr(zc)))
        return y[:N], zc[:N]


def expand_groups(g: np.ndarray) -> np.ndarray:
    """grouped stream -> MATCH_DTYPE records, in order (what the host shim does while it calls report())"""
    g = np.asarray(g, np.int32)
    heads, i = [], 0
    while i < g.size:  # (group headers are found by walking: a Python loop over groups, vectorised inside)
        heads.append(i)
        i += 4 + int(g[i + 3])
    heads = np.asarray(heads, np.int64)
    n = g[heads + 3].astype(np.int64) if heads.size else np.zeros(0, np.int64)
    out = np.zeros(int(n.sum()), MATCH_DTYPE)
    if heads.size:
        rep = np.repeat(np.arange(heads.size), n)
        first = np.cumsum(n) - n
        within = np.arange(out.size) - first[rep]
        out["ai"], out["start"], out["end"] = g[heads][rep], g[heads + 1][rep], g[heads + 2][rep]
        out["bi"] = g[heads[rep] + 4 + within]
    return out
