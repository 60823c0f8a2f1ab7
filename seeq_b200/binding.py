"""ctypes binding of libseeq_b200.so (the C-ABI of include/*.h).

This module is plumbing for the tests, bench.py and the multi-GPU driver; the
product is the shared library.  Every call goes through the same C entry points
a C program (or the re-linked reference CLI / CPython module) uses.  If the
library is missing it is built; if no CUDA device is present the matching calls
fail loudly -- there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

# libseeq.h / seeq.h / seeq_b200.h constants
SQ_FIRST, SQ_BEST, SQ_ALL, SQ_COUNT = 0, 1, 2, 3
SQ_FAIL, SQ_CONVERT, SQ_IGNORE = 0, 4, 8
SQ_LINES, SQ_STREAM = 0, 0x10
SQ_ANY, SQ_MATCH, SQ_NOMATCH, SQ_COUNTLINES, SQ_COUNTMATCH = 0, 1, 2, 3, 4
SQB_COUNT_ONLY, SQB_FASTA, SQB_SINGLE_LINE, SQB_TIMING, SQB_KEEP_LINES = 0x100, 0x200, 0x400, 0x800, 0x1000
SQB_DEVICE_RESULTS = 0x2000
SQB_FASTQ = 0x4000

REC_DTYPE = np.dtype([("line", "<u4"), ("start", "<u4"), ("end", "<u4"), ("dist", "<u4")])


class MatchT(C.Structure):
    _fields_ = [("start", C.c_size_t), ("end", C.c_size_t), ("dist", C.c_size_t)]


class SeeqT(C.Structure):
    _fields_ = [("hits", C.c_size_t), ("stacksize", C.c_size_t), ("match", C.POINTER(MatchT)),
                ("bufsz", C.c_size_t), ("string", C.c_char_p), ("tau", C.c_int), ("wlen", C.c_int),
                ("keys", C.POINTER(C.c_char)), ("rkeys", C.POINTER(C.c_char)),
                ("dfa", C.c_void_p), ("rdfa", C.c_void_p)]


class SeeqFileT(C.Structure):
    _fields_ = [("flags", C.c_int), ("line", C.c_size_t), ("info", C.c_char_p), ("fdi", C.c_void_p)]


class SeeqArgT(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("showdist", "showpos", "showline", "printline", "matchonly", "count", "compact",
                 "dist", "verbose", "endline", "prefix", "split", "invert", "best", "non_dna", "all")] + \
               [("memory", C.c_size_t)]


class StatsT(C.Structure):
    _fields_ = [("nbytes", C.c_uint64), ("nlines", C.c_uint64), ("nmatched", C.c_uint64),
                ("nrecs", C.c_uint64), ("device_ms", C.c_double), ("kernel_ms", C.c_double * 8),
                ("launches", C.c_uint32), ("reruns", C.c_uint32), ("path", C.c_uint32), ("devices", C.c_uint32)]


SQB_PATH_BITSLICE, SQB_PATH_FUSED, SQB_PATH_CUTS, SQB_PATH_FILTER = 1, 2, 4, 8


class BgzfMemberT(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("in_len", C.c_uint32), ("isize", C.c_uint32), ("out_off", C.c_uint64)]


class GenT(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("line_len", C.c_uint32), ("plant_per_1024", C.c_uint32),
                ("max_edits", C.c_uint32), ("n_per_1024", C.c_uint32), ("junk_per_1024", C.c_uint32),
                ("fastq", C.c_uint32), ("plant_len", C.c_uint32), ("plant", C.c_char * 256)]


# every symbol include/*.h declares: name -> (restype, argtypes)
_u64p = C.POINTER(C.c_uint64)
_SEEQ = C.POINTER(SeeqT)
_FILE = C.POINTER(SeeqFileT)
SYMBOLS = {
    # libseeq.h
    "seeqNew": (_SEEQ, [C.c_char_p, C.c_int, C.c_size_t]),
    "seeqFree": (None, [_SEEQ]),
    "seeqMatchIter": (C.POINTER(MatchT), [_SEEQ]),
    "seeqGetString": (C.c_char_p, [_SEEQ]),
    "seeqStringMatch": (C.c_long, [C.c_char_p, _SEEQ, C.c_int]),
    "seeqPrintError": (C.c_char_p, []),
    "seeqAddMatch": (C.c_int, [_SEEQ, MatchT]),
    "stackNew": (C.c_void_p, [C.c_size_t]),
    "stackAddMatch": (C.c_int, [C.POINTER(C.c_void_p), MatchT]),
    "recursive_merge": (C.c_int, [C.c_size_t, C.c_size_t, C.c_int, _SEEQ, C.c_void_p]),
    # seeq.h
    "seeq": (C.c_int, [C.c_char_p, C.c_char_p, SeeqArgT]),
    "seeqFileMatch": (C.c_long, [_FILE, _SEEQ, C.c_int, C.c_int]),
    "seeqOpen": (_FILE, [C.c_char_p]),
    "seeqClose": (C.c_int, [_FILE]),
    # seeq_b200.h
    "sqbEngineNew": (C.c_void_p, [C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "sqbEngineFree": (None, [C.c_void_p]),
    "sqbLastError": (C.c_char_p, []),
    "sqbDeviceCount": (C.c_int, []),
    "sqbMaxPatternLength": (C.c_int, []),
    "sqbScanDevice": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(StatsT)]),
    "sqbScanDeviceIssue": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "sqbScanDeviceWait": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(StatsT)]),
    "sqbScanDeviceLarge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(StatsT)]),
    "sqbDeviceRecordsAll": (C.c_void_p, [C.c_void_p, _u64p]),
    "sqbDeviceRecords": (C.c_void_p, [C.c_void_p]),
    "sqbDeviceLineStarts": (C.c_void_p, [C.c_void_p]),
    "sqbFetchRecords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "sqbFetchLineStarts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "sqbScanHost": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(StatsT)]),
    "sqbHostRecords": (C.c_void_p, [C.c_void_p, _u64p]),
    "sqbScanGeneration": (C.c_ulonglong, [C.c_void_p]),
    "sqbHostLineStarts": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), _u64p]),
    "sqbHostAlloc": (C.c_void_p, [C.c_size_t]),
    "sqbHostFree": (None, [C.c_void_p]),
    "sqbDeviceAlloc": (C.c_void_p, [C.c_size_t]),
    "sqbDeviceFree": (None, [C.c_void_p]),
    "sqbMemcpyH2D": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "sqbMemcpyD2H": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "sqbMultiNew": (C.c_void_p, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
    "sqbMultiFree": (None, [C.c_void_p]),
    "sqbMultiCount": (C.c_int, [C.c_void_p]),
    "sqbMultiEngine": (C.c_void_p, [C.c_void_p, C.c_int]),
    "sqbMultiScanHost": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(StatsT)]),
    "sqbMultiScanDevice": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(StatsT)]),
    "seeqBatchMatch": (C.c_long, [_SEEQ, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                  C.POINTER(C.c_void_p), C.POINTER(StatsT)]),
    "seeqEngine": (C.c_void_p, [_SEEQ]),
    "sqbBgzfIndex": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(BgzfMemberT), C.c_uint64, _u64p, _u64p]),
    "sqbBgzfInflateDevice": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(BgzfMemberT), C.c_uint64, C.c_void_p,
                                       C.c_void_p, C.POINTER(C.c_double)]),
    "sqbScanHostBgzf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(StatsT)]),
    "sqbBgzfDeviceText": (C.c_void_p, [C.c_void_p, _u64p]),
    "sqbEngineDevice": (C.c_int, [C.c_void_p]),
    "sqbShardRange": (None, [C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                             C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "sqbGenBytes": (C.c_size_t, [C.POINTER(GenT), C.c_uint64, C.c_uint64]),
    "sqbGenHost": (C.c_int, [C.POINTER(GenT), C.c_uint64, C.c_uint64, C.c_void_p]),
    "sqbGenDevice": (C.c_int, [C.POINTER(GenT), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (building if necessary) libseeq_b200.so and type its entry points."""
    global _lib
    if _lib is None:
        path = os.environ.get("SEEQ_B200_LIB") or _build.LIB      # experiments: another build of the library
        if path == _build.LIB and not os.path.exists(path):
            _build.build_library()
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)       # AttributeError = symbol missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def seeqerr() -> int:
    return C.c_int.in_dll(lib(), "seeqerr").value


def last_error() -> str:
    return lib().sqbLastError().decode()


def _ptr(buf):
    """(address, nbytes, keepalive) of bytes / numpy uint8 array."""
    if isinstance(buf, np.ndarray):
        assert buf.dtype == np.uint8 and buf.flags.c_contiguous
        return buf.ctypes.data, buf.size, buf
    if isinstance(buf, str):
        buf = buf.encode("latin-1")
    b = bytes(buf)
    return C.cast(C.c_char_p(b), C.c_void_p).value, len(b), b


def _as_recs(ptr, n) -> np.ndarray:
    if n == 0:
        return np.zeros(0, dtype=REC_DTYPE)
    raw = (C.c_uint32 * (4 * n)).from_address(ptr)
    return np.frombuffer(raw, dtype=REC_DTYPE).copy()


class Seeq:
    """seeq_t through the libseeq API (seeqNew ... seeqFree)."""

    def __init__(self, pattern, tau: int, memory: int = 0):
        self.L = lib()
        p = pattern.encode() if isinstance(pattern, str) else pattern
        self.sq = self.L.seeqNew(p, tau, memory)
        if not self.sq:
            raise ValueError("seeqNew failed: seeqerr=%d (%s)" % (seeqerr(), self.L.seeqPrintError().decode()))

    def close(self):
        if self.sq:
            self.L.seeqFree(self.sq)
            self.sq = None

    __del__ = close

    @property
    def keys(self) -> bytes:
        return bytes(self.sq.contents.keys[i][0] for i in range(self.sq.contents.wlen))

    def string_match(self, text, options: int = 0):
        """seeqStringMatch + seeqMatchIter -> list of (start, end, dist), left to right."""
        t = text.encode("latin-1") if isinstance(text, str) else bytes(text)
        n = self.L.seeqStringMatch(t, self.sq, options)
        if n < 0:
            raise RuntimeError("seeqStringMatch failed: " + last_error())
        out = []
        while True:
            m = self.L.seeqMatchIter(self.sq)
            if not m:
                break
            out.append((m.contents.start, m.contents.end, m.contents.dist))
        assert len(out) == n
        return out

    def batch(self, buf, match_opt: int = 0, file_opt: int = SQ_ANY, stats: StatsT | None = None):
        """seeqBatchMatch -> records (SQ_ANY) or a count (SQ_COUNTLINES / SQ_COUNTMATCH)."""
        addr, n, keep = _ptr(buf)
        recs = C.c_void_p()
        st = stats if stats is not None else StatsT()
        r = self.L.seeqBatchMatch(self.sq, addr, n, match_opt, file_opt, C.byref(recs), C.byref(st))
        if r < 0:
            raise RuntimeError("seeqBatchMatch failed: " + last_error())
        if file_opt != SQ_ANY:
            return r
        return _as_recs(recs.value, r)

    def engine(self) -> int:
        e = self.L.seeqEngine(self.sq)
        if not e:
            raise RuntimeError("no engine: " + last_error())
        return e


class Engine:
    """sqb_engine_t directly (device-resident scans)."""

    def __init__(self, keys: bytes, tau: int, device: int = -1):
        self.L = lib()
        self.owned = True
        self.e = self.L.sqbEngineNew(keys, len(keys), tau, device)
        if not self.e:
            raise RuntimeError("sqbEngineNew failed: " + last_error())

    @classmethod
    def borrowed(cls, ptr: int) -> "Engine":
        """View of an engine owned by someone else (a seeq_t: seeqEngine())."""
        self = cls.__new__(cls)
        self.L, self.e, self.owned = lib(), ptr, False
        return self

    def close(self):
        if getattr(self, "e", None) and self.owned:
            self.L.sqbEngineFree(self.e)
        self.e = None

    __del__ = close

    def scan_device(self, d_ptr: int, nbytes: int, options: int, stream: int = 0) -> StatsT:
        st = StatsT()
        if self.L.sqbScanDevice(self.e, d_ptr, nbytes, options, stream, C.byref(st)):
            raise RuntimeError("sqbScanDevice failed: " + last_error())
        return st

    def scan_device_large(self, d_ptr: int, nbytes: int, options: int, stream: int = 0) -> StatsT:
        """Any size; records via host_records() (buffer-global line indices)."""
        st = StatsT()
        if self.L.sqbScanDeviceLarge(self.e, d_ptr, nbytes, options, stream, C.byref(st)):
            raise RuntimeError("sqbScanDeviceLarge failed: " + last_error())
        return st

    def scan_device_issue(self, slot: int, d_ptr: int, nbytes: int, options: int, stream: int = 0) -> None:
        if self.L.sqbScanDeviceIssue(self.e, slot, d_ptr, nbytes, options, stream):
            raise RuntimeError("sqbScanDeviceIssue failed: " + last_error())

    def scan_device_wait(self, slot: int) -> StatsT:
        st = StatsT()
        if self.L.sqbScanDeviceWait(self.e, slot, C.byref(st)):
            raise RuntimeError("sqbScanDeviceWait failed: " + last_error())
        return st

    def scan_host(self, buf, options: int) -> StatsT:
        addr, n, keep = _ptr(buf)
        st = StatsT()
        if self.L.sqbScanHost(self.e, addr, n, options, C.byref(st)):
            raise RuntimeError("sqbScanHost failed: " + last_error())
        return st

    def scan_host_ptr(self, addr: int, n: int, options: int) -> StatsT:
        st = StatsT()
        if self.L.sqbScanHost(self.e, addr, n, options, C.byref(st)):
            raise RuntimeError("sqbScanHost failed: " + last_error())
        return st

    def scan_host_bgzf(self, buf, options: int) -> StatsT:
        """sqbScanHostBgzf: `buf` is a BGZF (bgzip) buffer; results as after scan_host of the inflated text."""
        addr, n, keep = _ptr(buf)
        return self.scan_host_bgzf_ptr(addr, n, options)

    def scan_host_bgzf_ptr(self, addr: int, n: int, options: int) -> StatsT:
        st = StatsT()
        if self.L.sqbScanHostBgzf(self.e, addr, n, options, C.byref(st)):
            raise RuntimeError("sqbScanHostBgzf failed: " + last_error())
        return st

    def bgzf_text(self) -> np.ndarray:
        """the inflated text of the last scan_host_bgzf, copied back from the device"""
        n = C.c_uint64(0)
        p = self.L.sqbBgzfDeviceText(self.e, C.byref(n))
        out = np.empty(n.value, dtype=np.uint8)
        if n.value and self.L.sqbMemcpyD2H(out.ctypes.data, p, n.value):
            raise RuntimeError("sqbMemcpyD2H failed: " + last_error())
        return out

    def host_records(self) -> np.ndarray:
        n = C.c_uint64(0)
        p = self.L.sqbHostRecords(self.e, C.byref(n))
        return _as_recs(p, n.value)

    def device_records_all(self) -> np.ndarray:
        """Records of the last chunked scan run with SQB_DEVICE_RESULTS, copied back."""
        n = C.c_uint64(0)
        p = self.L.sqbDeviceRecordsAll(self.e, C.byref(n))
        out = np.zeros(n.value, dtype=REC_DTYPE)
        if n.value and self.L.sqbMemcpyD2H(out.ctypes.data, p, n.value * 16):
            raise RuntimeError("sqbMemcpyD2H failed: " + last_error())
        return out

    def host_line_starts(self) -> np.ndarray:
        n = C.c_uint64(0)
        p = C.c_void_p()
        self.L.sqbHostLineStarts(self.e, C.byref(p), C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=np.uint64)
        raw = (C.c_uint64 * n.value).from_address(p.value)
        return np.frombuffer(raw, dtype=np.uint64).copy()

    def fetch_records(self, count: int, first: int = 0) -> np.ndarray:
        out = np.zeros(count, dtype=REC_DTYPE)
        if count and self.L.sqbFetchRecords(self.e, out.ctypes.data, first, count):
            raise RuntimeError("sqbFetchRecords failed: " + last_error())
        return out

    def fetch_line_starts(self, count: int, first: int = 0) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint32)
        if count and self.L.sqbFetchLineStarts(self.e, out.ctypes.data, first, count):
            raise RuntimeError("sqbFetchLineStarts failed: " + last_error())
        return out


class Multi:
    """sqb_multi_t: a pattern set scanned over one pass of the text."""

    def __init__(self, keys_list, taus, device: int = -1):
        self.L = lib()
        n = len(keys_list)
        self.n = n
        self._keep = [bytes(k) for k in keys_list]
        arr = (C.c_char_p * n)(*self._keep)
        ms = (C.c_int * n)(*[len(k) for k in self._keep])
        ts = (C.c_int * n)(*taus)
        self.mp = self.L.sqbMultiNew(n, arr, ms, ts, device)
        if not self.mp:
            raise RuntimeError("sqbMultiNew failed: " + last_error())

    def close(self):
        if getattr(self, "mp", None):
            self.L.sqbMultiFree(self.mp)
        self.mp = None

    __del__ = close

    def scan_host(self, buf, options: int):
        addr, n, keep = _ptr(buf)
        st = (StatsT * self.n)()
        if self.L.sqbMultiScanHost(self.mp, addr, n, options, st):
            raise RuntimeError("sqbMultiScanHost failed: " + last_error())
        return list(st)

    def scan_host_ptr(self, addr: int, n: int, options: int):
        st = (StatsT * self.n)()
        if self.L.sqbMultiScanHost(self.mp, addr, n, options, st):
            raise RuntimeError("sqbMultiScanHost failed: " + last_error())
        return list(st)

    def scan_device(self, d_ptr: int, nbytes: int, options: int, stream: int = 0):
        st = (StatsT * self.n)()
        if self.L.sqbMultiScanDevice(self.mp, d_ptr, nbytes, options, stream, st):
            raise RuntimeError("sqbMultiScanDevice failed: " + last_error())
        return list(st)

    def records(self, pattern: int) -> np.ndarray:
        return Engine.borrowed(self.L.sqbMultiEngine(self.mp, pattern)).host_records()


def make_gen(seed: int, line_len: int, plant: str = "", plant_per_1024: int = 0, max_edits: int = 0,
             n_per_1024: int = 0, junk_per_1024: int = 0, fastq: bool = False) -> GenT:
    g = GenT()
    g.seed = seed
    g.line_len = line_len
    g.plant_per_1024 = plant_per_1024
    g.max_edits = max_edits
    g.n_per_1024 = n_per_1024
    g.junk_per_1024 = junk_per_1024
    g.fastq = 1 if fastq else 0
    g.plant_len = len(plant)
    g.plant = plant.encode()
    return g


def gen_host(g: GenT, nreads: int, first: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    L = lib()
    nbytes = L.sqbGenBytes(C.byref(g), first, nreads)
    if out is None:
        out = np.empty(nbytes, dtype=np.uint8)
    assert out.size >= nbytes
    L.sqbGenHost(C.byref(g), first, nreads, out.ctypes.data)
    return out[:nbytes]


def bgzf_index(buf):
    """sqbBgzfIndex -> (array of BgzfMemberT, bytes of text); host only"""
    addr, n, keep = _ptr(buf)
    L = lib()
    cnt, tb = C.c_uint64(0), C.c_uint64(0)
    if L.sqbBgzfIndex(addr, n, None, 0, C.byref(cnt), C.byref(tb)):
        raise ValueError("sqbBgzfIndex failed: " + last_error())
    members = (BgzfMemberT * max(1, cnt.value))()
    if L.sqbBgzfIndex(addr, n, members, cnt.value, C.byref(cnt), C.byref(tb)):
        raise ValueError("sqbBgzfIndex failed: " + last_error())
    return members, cnt.value, tb.value


def bgzf_inflate_device(buf, device: int = 0):
    """A BGZF buffer inflated on the device (H2D, sqbBgzfInflateDevice, D2H) -> (text as uint8 array, kernel ms)."""
    addr, n, keep = _ptr(buf)
    L = lib()
    members, cnt, tb = bgzf_index(buf)
    d_gz = L.sqbDeviceAlloc(n + 64)
    d_text = L.sqbDeviceAlloc(tb + 64)
    if not d_gz or not d_text:
        raise RuntimeError("sqbDeviceAlloc failed: " + last_error())
    try:
        if n and L.sqbMemcpyH2D(d_gz, addr, n):
            raise RuntimeError("sqbMemcpyH2D failed: " + last_error())
        ms = C.c_double(0)
        if L.sqbBgzfInflateDevice(device, d_gz, members, cnt, d_text, None, C.byref(ms)):
            raise RuntimeError("sqbBgzfInflateDevice failed: " + last_error())
        out = np.empty(tb, dtype=np.uint8)
        if tb and L.sqbMemcpyD2H(out.ctypes.data, d_text, tb):
            raise RuntimeError("sqbMemcpyD2H failed: " + last_error())
        return out, ms.value
    finally:
        L.sqbDeviceFree(d_gz)
        L.sqbDeviceFree(d_text)


def shard_range(buf: np.ndarray, rank: int, world: int):
    b = C.c_size_t(0)
    e = C.c_size_t(0)
    lib().sqbShardRange(buf.ctypes.data, buf.size, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value
