"""ctypes wrappers around the CHECKERS (test infrastructure, never shipped).

* ``Oracle``    -- oracle/liboracle.so, our CPU restatement (seeq_oracle.c)
* ``Reference`` -- oracle/_ref/libseeq_ref.so, the unmodified reference
                   compiled in place from /root/reference (oracle/Makefile)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libseeq_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "seeq_ref")
REFERENCE_DIR = os.environ.get("SEEQ_REFERENCE_DIR", "/root/reference")

SQ_FIRST, SQ_BEST, SQ_ALL = 0, 1, 2
SQ_FAIL, SQ_CONVERT, SQ_IGNORE = 0, 4, 8
SQ_STREAM = 0x10
SQ_ANY, SQ_MATCH, SQ_NOMATCH, SQ_COUNTLINES, SQ_COUNTMATCH = 0, 1, 2, 3, 4


def build(force: bool = False) -> None:
    """Compile liboracle.so and, when the reference tree is mounted, _ref/."""
    args = ["make", "-C", HERE, "REF=" + REFERENCE_DIR]
    if force:
        args.append("-B")
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


class _Rec(C.Structure):
    _fields_ = [("line", C.c_uint64), ("start", C.c_uint64),
                ("end", C.c_uint64), ("dist", C.c_uint64)]


def _as_bytes(x) -> bytes:
    return x.encode("latin-1") if isinstance(x, str) else bytes(x)


class Oracle:
    """CPU restatement of the matching path (seeq_oracle.c)."""

    def __init__(self) -> None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.orc_parse.restype = C.c_int
        L.orc_parse.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
        L.orc_code.restype = C.c_int
        L.orc_code.argtypes = [C.c_ubyte, C.c_int]
        L.orc_distances.restype = C.c_int
        L.orc_distances.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
        pp = C.POINTER(C.POINTER(_Rec))
        L.orc_string_match.restype = C.c_long
        L.orc_string_match.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                       C.c_int, pp, C.POINTER(C.c_size_t)]
        L.orc_string_match_segmented.restype = C.c_long
        L.orc_string_match_segmented.argtypes = [
            C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            pp, C.POINTER(C.c_size_t)]
        L.orc_buffer_scan.restype = C.c_long
        L.orc_buffer_scan.argtypes = [
            C.c_char_p, C.c_size_t, C.c_char_p, C.c_int, C.c_int, C.c_int, pp,
            C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        self.L = L
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]

    # -- pattern -----------------------------------------------------------
    def parse(self, pattern):
        """-> (keys bytes, 0) or (None, seeqerr)."""
        p = _as_bytes(pattern)
        keys = C.create_string_buffer(max(1, len(p)))
        err = C.c_int(0)
        m = self.L.orc_parse(p, keys, C.byref(err))
        if m < 0:
            return None, err.value
        return keys.raw[:m], 0

    def code(self, byte: int, convert: bool = False) -> int:
        return self.L.orc_code(byte, int(convert))

    def distances(self, text, pattern, tau):
        keys, err = self.parse(pattern)
        assert keys is not None, err
        t = _as_bytes(text)
        d = (C.c_int * (len(t) + 1))()
        mm = (C.c_int * (len(t) + 1))()
        n = self.L.orc_distances(t, keys, len(keys), tau, d, mm)
        return list(d[:n]), list(mm[:n])

    # -- matching ----------------------------------------------------------
    def _take(self, recs, n):
        out = np.empty((max(n, 0), 4), dtype=np.uint64)
        if n > 0:
            src = np.ctypeslib.as_array(C.cast(recs, C.POINTER(C.c_uint64)),
                                        shape=(n, 4))
            out[:] = src
        if recs:
            self.libc.free(C.cast(recs, C.c_void_p))
        return out

    def string_match(self, text, keys: bytes, tau: int, options: int):
        """-> uint64 array (n,4): line(=0), start, end, dist; left-to-right."""
        recs = C.POINTER(_Rec)()
        cap = C.c_size_t(0)
        n = self.L.orc_string_match(_as_bytes(text), keys, len(keys), tau,
                                    options, C.byref(recs), C.byref(cap))
        assert n >= 0
        return self._take(recs, n)

    def string_match_segmented(self, text, keys, tau, options, seg, warm):
        recs = C.POINTER(_Rec)()
        cap = C.c_size_t(0)
        n = self.L.orc_string_match_segmented(
            _as_bytes(text), keys, len(keys), tau, options, seg, warm,
            C.byref(recs), C.byref(cap))
        assert n >= 0
        return self._take(recs, n)

    def buffer_scan(self, buf, keys: bytes, tau: int, options: int):
        """-> (records (n,4) uint64, counted lines, lines with >= 1 match)."""
        b = _as_bytes(buf) if not isinstance(buf, np.ndarray) else buf
        if isinstance(b, np.ndarray):
            ptr = b.ctypes.data_as(C.c_char_p)
            n_in = b.size
        else:
            ptr = b
            n_in = len(b)
        recs = C.POINTER(_Rec)()
        cap = C.c_size_t(0)
        nl = C.c_uint64(0)
        nm = C.c_uint64(0)
        n = self.L.orc_buffer_scan(ptr, n_in, keys, len(keys), tau, options,
                                   C.byref(recs), C.byref(cap), C.byref(nl),
                                   C.byref(nm))
        assert n >= 0
        return self._take(recs, n), nl.value, nm.value


class Reference:
    """The unmodified reference behind oracle/ref_driver.c."""

    def __init__(self) -> None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = C.CDLL(REF_SO)
        u64p = C.POINTER(C.c_uint64)
        L.ref_string_match.restype = C.c_long
        L.ref_string_match.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                       u64p, C.c_long]
        L.ref_buffer_scan.restype = C.c_long
        L.ref_buffer_scan.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p,
                                      C.c_int, C.c_int, u64p, C.c_long, u64p, u64p]
        L.ref_buffer_count.restype = C.c_long
        L.ref_buffer_count.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p,
                                       C.c_int, C.c_int, C.c_int]
        L.ref_bench.restype = C.c_double
        L.ref_bench.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.ref_bench_ck.restype = C.c_double
        L.ref_bench_ck.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long), u64p]
        self.L = L

    def string_match(self, pattern, tau, text, options, cap=4096):
        """-> (hits or negative error, (k,3) uint64 start,end,dist)."""
        out = np.zeros((cap, 3), dtype=np.uint64)
        n = self.L.ref_string_match(_as_bytes(pattern), tau, _as_bytes(text),
                                    options, out.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
        return n, out[:max(0, min(n, cap))]

    @staticmethod
    def _ptr(buf):
        if isinstance(buf, np.ndarray):
            return buf.ctypes.data_as(C.c_char_p), buf.size
        b = _as_bytes(buf)
        return b, len(b)

    def buffer_scan(self, buf, pattern, tau, options, cap=None):
        ptr, n_in = self._ptr(buf)
        if cap is None:
            cap = max(1024, n_in)
        out = np.zeros((cap, 4), dtype=np.uint64)
        nl = C.c_uint64(0)
        nm = C.c_uint64(0)
        n = self.L.ref_buffer_scan(ptr, n_in, _as_bytes(pattern), tau, options,
                                   out.ctypes.data_as(C.POINTER(C.c_uint64)),
                                   cap, C.byref(nl), C.byref(nm))
        assert 0 <= n <= cap, n
        return out[:n], nl.value, nm.value

    def buffer_count(self, buf, pattern, tau, options, file_opt):
        ptr, n_in = self._ptr(buf)
        return self.L.ref_buffer_count(ptr, n_in, _as_bytes(pattern), tau,
                                       options, file_opt)

    def bench(self, buf, pattern, tau, options, mode, nproc):
        """-> (wall seconds, summed result)."""
        ptr, n_in = self._ptr(buf)
        total = C.c_long(0)
        s = self.L.ref_bench(ptr, n_in, _as_bytes(pattern), tau, options, mode,
                             nproc, C.byref(total))
        return s, total.value

    def bench_ck(self, buf, pattern, tau, options, mode, nproc):
        """-> (wall seconds, summed result, checksum of the records: see records_checksum)."""
        ptr, n_in = self._ptr(buf)
        total = C.c_long(0)
        ck = C.c_uint64(0)
        s = self.L.ref_bench_ck(ptr, n_in, _as_bytes(pattern), tau, options, mode,
                                nproc, C.byref(total), C.byref(ck))
        return s, total.value, ck.value


CK_P = (0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x27D4EB2F165667C5, 0x85EBCA77C2B2AE63)


def records_checksum(line1, start, end, dist) -> int:
    """The checksum of oracle/ref_driver.c (ref_bench_ck) over arrays of records; line1 = 1-based
    buffer-global line numbers.  Arithmetic mod 2^64 (numpy uint64 wraps)."""
    with np.errstate(over="ignore"):
        p = [np.uint64(x) for x in CK_P]
        a = (np.asarray(line1, dtype=np.uint64) * p[0] + np.asarray(start, dtype=np.uint64) * p[1] +
             np.asarray(end, dtype=np.uint64) * p[2] + np.asarray(dist, dtype=np.uint64) * p[3] + p[4])
        return int(a.sum(dtype=np.uint64))
