"""ctypes loader for the CPU oracles (TEST INFRASTRUCTURE ONLY).

Two libraries export the same flat API (oracle/oracle_api.h):
  * ``ref``  -- oracle/_ref/libpbwtref.so: the unmodified reference compiled in place.
  * ``orc``  -- oracle/libpbwt_oracle.so : the from-scratch restatement (kind "port").
Nothing under pbwt_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_u8p = C.POINTER(C.c_uint8)
c_i64p = C.POINTER(C.c_int64)

MATCH_DTYPE = np.dtype([("ai", "<i4"), ("bi", "<i4"), ("start", "<i4"), ("end", "<i4")])
NOMATCH_DTYPE = np.dtype([("query", "<i4"), ("allele", "<i4"), ("site", "<i4")])


def words_per_col(M: int) -> int:
    return 4 * ((M + 127) // 128)


def pack_panel(hap: np.ndarray) -> np.ndarray:
    """hap: uint8 [M, N] (haplotype-major 0/1) -> site-major bit columns uint32 [N, wpc]."""
    M, N = hap.shape
    wpc = words_per_col(M)
    bits = np.zeros((N, wpc * 32), dtype=np.uint8)
    bits[:, :M] = hap.T
    packed = np.packbits(bits, axis=1, bitorder="little")
    return np.ascontiguousarray(packed).view("<u4").reshape(N, wpc)


def unpack_panel(xbits: np.ndarray, M: int) -> np.ndarray:
    N = xbits.shape[0]
    b = np.unpackbits(np.ascontiguousarray(xbits).view(np.uint8).reshape(N, -1), axis=1, bitorder="little")
    return np.ascontiguousarray(b[:, :M].T)


def _p(arr, typ):
    return None if arr is None else arr.ctypes.data_as(typ)


class Oracle:
    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.path = path

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- build ---------------------------------------------------------------------------
    def build(self, xbits, M, N, a_start=None, with_d=True):
        wpc = words_per_col(M)
        xbits = np.ascontiguousarray(xbits, dtype=np.uint32).reshape(N, wpc) if N else np.zeros((0, wpc), np.uint32)
        cap = max(16, N * (M + 8))
        yz = np.zeros(cap, np.uint8)
        nyz = C.c_int64(0)
        a_end = np.zeros(M, np.int32)
        d_end = np.zeros(M + 1, np.int32)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        rc = self._f("build")(_p(xbits, c_u32p), wpc, M, N, _p(a0, c_i32p), int(with_d), _p(yz, c_u8p),
                              C.c_int64(cap), C.byref(nyz), _p(a_end, c_i32p), _p(d_end, c_i32p))
        assert rc == 0
        return yz[: nyz.value].copy(), a_end, (d_end if with_d else None)

    def trace(self, xbits, M, N, a_start=None):
        wpc = words_per_col(M)
        xbits = np.ascontiguousarray(xbits, dtype=np.uint32)
        a_tr = np.zeros((N + 1, M), np.int32)
        d_tr = np.zeros((N + 1, M + 1), np.int32)
        y_tr = np.zeros((max(N, 1), M), np.uint8)
        c_tr = np.zeros(max(N, 1), np.int32)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        rc = self._f("trace")(_p(xbits, c_u32p), wpc, M, N, _p(a0, c_i32p), _p(a_tr, c_i32p), _p(d_tr, c_i32p),
                              _p(y_tr, c_u8p), _p(c_tr, c_i32p))
        assert rc == 0
        return a_tr, d_tr, y_tr[:N], c_tr[:N]

    # ---- sweeps over a stored stream -----------------------------------------------------
    def max_within(self, yz, M, N, a_start=None, cap=None):
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        n = C.c_int64(0)
        if cap is None:  # count first
            self._f("max_within")(_p(yz, c_u8p), C.c_int64(yz.size), M, N, _p(a0, c_i32p), None, C.c_int64(0), C.byref(n))
            cap = n.value
        rec = np.zeros(max(cap, 1), MATCH_DTYPE)
        self._f("max_within")(_p(yz, c_u8p), C.c_int64(yz.size), M, N, _p(a0, c_i32p),
                              rec.ctypes.data_as(C.c_void_p), C.c_int64(cap), C.byref(n))
        return rec[: min(cap, n.value)], n.value

    def long_within(self, yz, M, N, T, a_start=None):
        """pbwtLongMatches(p, T), T > 0 (matchLongWithin2, pbwtMatch.c:85-113)."""
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        n = C.c_int64(0)
        self._f("long_within")(_p(yz, c_u8p), C.c_int64(yz.size), M, N, _p(a0, c_i32p), T, None, C.c_int64(0), C.byref(n))
        cap = n.value
        rec = np.zeros(max(cap, 1), MATCH_DTYPE)
        self._f("long_within")(_p(yz, c_u8p), C.c_int64(yz.size), M, N, _p(a0, c_i32p), T,
                               rec.ctypes.data_as(C.c_void_p), C.c_int64(cap), C.byref(n))
        return rec[:cap], n.value

    def build_reverse(self, yz, M, N, a_start=None):
        """pbwtBuildReverse (pbwtCore.c:151-191) -> (aFend, zz, aRend)."""
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        cap = M * max(N, 1) + 16
        zz = np.zeros(cap, np.uint8)
        a_fend, a_rend = np.zeros(M, np.int32), np.zeros(M, np.int32)
        n = C.c_int64(0)
        rc = self._f("build_reverse")(_p(yz, c_u8p), C.c_int64(yz.size), M, N, _p(a0, c_i32p), _p(a_fend, c_i32p),
                                      _p(zz, c_u8p), C.c_int64(cap), C.byref(n), _p(a_rend, c_i32p))
        assert rc == 0
        return a_fend, zz[: n.value].copy(), a_rend

    def sweep_ad(self, yz, M, N, a_start=None):
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        a_end = np.zeros(M, np.int32)
        d_end = np.zeros(M + 1, np.int32)
        self._f("sweep_ad")(_p(yz, c_u8p), C.c_int64(yz.size), M, N, _p(a0, c_i32p), _p(a_end, c_i32p), _p(d_end, c_i32p))
        return a_end, d_end

    def sweep_a(self, yz, M, N, a_start=None):
        yz = np.ascontiguousarray(yz, np.uint8)
        a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
        a_end = np.zeros(M, np.int32)
        self._f("sweep_a")(_p(yz, c_u8p), C.c_int64(yz.size), M, N, _p(a0, c_i32p), _p(a_end, c_i32p))
        return a_end

    def match_sweep(self, pyz, M, N, qyz, Q, pa_start=None, qa_start=None):
        pyz = np.ascontiguousarray(pyz, np.uint8)
        qyz = np.ascontiguousarray(qyz, np.uint8)
        pa = None if pa_start is None else np.ascontiguousarray(pa_start, np.int32)
        qa = None if qa_start is None else np.ascontiguousarray(qa_start, np.int32)
        n = C.c_int64(0)
        nn = C.c_int64(0)
        args = (_p(pyz, c_u8p), C.c_int64(pyz.size), M, N, _p(pa, c_i32p), _p(qyz, c_u8p), C.c_int64(qyz.size), Q, _p(qa, c_i32p))
        self._f("match_sweep")(*args, None, C.c_int64(0), C.byref(n), None, C.c_int64(0), C.byref(nn))
        rec = np.zeros(max(n.value, 1), MATCH_DTYPE)
        nom = np.zeros(max(nn.value, 1), NOMATCH_DTYPE)
        cap, nmcap = n.value, nn.value
        self._f("match_sweep")(*args, rec.ctypes.data_as(C.c_void_p), C.c_int64(cap), C.byref(n),
                               nom.ctypes.data_as(C.c_void_p), C.c_int64(nmcap), C.byref(nn))
        return rec[:cap], nom[:nmcap]

    # ---- codec ---------------------------------------------------------------------------
    def pack3(self, y):
        y = np.ascontiguousarray(y, np.uint8)
        M = y.size
        buf = np.zeros(M + 1, np.uint8)
        buf[:M] = y
        out = np.zeros(M + 8, np.uint8)
        n = self._f("pack3")(_p(buf, c_u8p), M, _p(out, c_u8p))
        return out[:n].copy()

    def unpack3(self, yz, M):
        yz = np.ascontiguousarray(yz, np.uint8)
        y = np.zeros(M + 64, np.uint8)
        n0 = C.c_int32(0)
        used = self._f("unpack3")(_p(yz, c_u8p), M, _p(y, c_u8p), C.byref(n0))
        return y[:M].copy(), used, n0.value

    def time_build_maxwithin(self, xbits, M, N):
        wpc = words_per_col(M)
        xbits = np.ascontiguousarray(xbits, np.uint32)
        tb, tm = C.c_double(0), C.c_double(0)
        nyz, nm = C.c_int64(0), C.c_int64(0)
        self._f("time_build_maxwithin")(_p(xbits, c_u32p), wpc, M, N, C.byref(tb), C.byref(tm), C.byref(nyz), C.byref(nm))
        return tb.value, tm.value, nyz.value, nm.value


    def time_build_maxwithin_ck(self, xbits, M, N):
        """-> (t_build, t_match, nyz, nmatch, ck[4]); ck = checksums of the pack3 stream and of the records (oracle_api.h)."""
        wpc = words_per_col(M)
        xbits = np.ascontiguousarray(xbits, np.uint32)
        tb, tm = C.c_double(0), C.c_double(0)
        nyz, nm = C.c_int64(0), C.c_int64(0)
        ck = np.zeros(4, np.uint64)
        self._f("time_build_maxwithin_ck")(_p(xbits, c_u32p), wpc, M, N, C.byref(tb), C.byref(tm), C.byref(nyz), C.byref(nm),
                                           ck.ctypes.data_as(C.c_void_p))
        return tb.value, tm.value, nyz.value, nm.value, [int(v) for v in ck]

    def time_build(self, xbits, M, N, with_d):
        """-> (t_build, nyz, ck[2]): the build alone (pbwtCursorWriteForwards / ...AD over all sites)."""
        wpc = words_per_col(M)
        xbits = np.ascontiguousarray(xbits, np.uint32)
        tb, nyz = C.c_double(0), C.c_int64(0)
        ck = np.zeros(2, np.uint64)
        self._f("time_build")(_p(xbits, c_u32p), wpc, M, N, int(with_d), C.byref(tb), C.byref(nyz), ck.ctypes.data_as(C.c_void_p))
        return tb.value, nyz.value, [int(v) for v in ck]


_K1, _K2 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F)


def checksum_parts(mix: np.ndarray):
    """(count, S, T) of a run of items: S = sum m_i, T = sum (i+1) m_i (mod 2^64)."""
    n = mix.size
    with np.errstate(over="ignore"):
        S = int(mix.sum(dtype=np.uint64))
        T = int((mix * np.arange(1, n + 1, dtype=np.uint64)).sum(dtype=np.uint64))
    return n, S, T


def checksum_join(parts):
    """checksums of a concatenation from the (count, S, T) of its pieces, in order -> [S, T]"""
    mask = (1 << 64) - 1
    off = S = T = 0
    for n, s, t in parts:
        S = (S + s) & mask
        T = (T + t + off * s) & mask
        off += n
    return [S, T]


def mix_bytes(b: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return b.astype(np.uint64) * _K1 + np.uint64(1)


def mix_records(rec: np.ndarray) -> np.ndarray:
    """oracle_api.h: oracle_mix_record over a MATCH_DTYPE array"""
    w = np.ascontiguousarray(rec).view("<u8").reshape(-1, 2)
    with np.errstate(over="ignore"):
        m = (w[:, 0] * _K1) ^ (w[:, 1] * _K2)
        m ^= m >> np.uint64(29)
    return m | np.uint64(1)


def sweep_from(xbits, M, N, k0, ntotal, a_start, d_start):
    """oracle/pbwt_oracle.c:orc_sweep_from -- build + maxWithin of one site shard from a given state."""
    o = port()
    wpc = words_per_col(M)
    xbits = np.ascontiguousarray(xbits, np.uint32)
    a0 = None if a_start is None else np.ascontiguousarray(a_start, np.int32)
    d0 = None if d_start is None else np.ascontiguousarray(d_start, np.int32)
    cap = max(16, N * (M + 8))
    yz = np.zeros(cap, np.uint8)
    nyz, nrec = C.c_int64(0), C.c_int64(0)
    a_end, d_end = np.zeros(M, np.int32), np.zeros(M + 1, np.int32)
    f = o.lib.orc_sweep_from
    args = lambda rec, rcap: (_p(xbits, c_u32p), wpc, M, N, k0, ntotal, _p(a0, c_i32p), _p(d0, c_i32p), _p(yz, c_u8p), C.c_int64(cap),
                              C.byref(nyz), _p(a_end, c_i32p), _p(d_end, c_i32p), rec, C.c_int64(rcap), C.byref(nrec))
    f(*args(None, 0))
    rec = np.zeros(max(nrec.value, 1), MATCH_DTYPE)
    n = nrec.value
    f(*args(rec.ctypes.data_as(C.c_void_p), n))
    return yz[: nyz.value].copy(), a_end, d_end, rec[:n]


def build_oracles():
    """(Re)compile the oracle libraries. The reference arm is only rebuilt where /root/reference exists."""
    subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, capture_output=True)
    subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)


_cache = {}


def port() -> Oracle:
    if "orc" not in _cache:
        path = os.path.join(ORACLE_DIR, "libpbwt_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, capture_output=True)
        o = Oracle(path, "orc_")
        o.lib.orc_generate.argtypes = [c_u32p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
        _cache["orc"] = o
    return _cache["orc"]


def reference() -> Oracle | None:
    """The compiled reference, or None when neither /root/reference nor a prebuilt _ref/ exists."""
    if "ref" not in _cache:
        path = os.path.join(ORACLE_DIR, "_ref", "libpbwtref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/pbwtCore.c"):
            subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)
        _cache["ref"] = Oracle(path, "ref_") if os.path.exists(path) else None
    return _cache["ref"]


def generate_panel(M: int, N: int, seed: int, L: int = 2000, mu: float = 2.0) -> np.ndarray:
    """Counter-based synthetic mosaic panel (oracle/pbwt_oracle.c: orc_generate) -> xbits [N, wpc]."""
    o = port()
    wpc = words_per_col(M)
    x = np.zeros((N, wpc), np.uint32)
    thresh = min(int(mu / M * 2**32), 2**32 - 1)
    o.lib.orc_generate(_p(x, c_u32p), wpc, M, N, C.c_uint64(seed), C.c_uint32(L), C.c_uint32(thresh))
    return x
