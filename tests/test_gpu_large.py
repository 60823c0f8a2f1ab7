"""GPU parity of sqbScanDeviceLarge (device-resident buffers of any size, scanned in
newline-aligned chunks where they lie) and parity at BASELINE.json's FULL sizes.

Small buffers with a chunk size of 1 MiB exercise what only this entry point has: chunks
that start at arbitrary (not 16-byte aligned) addresses (K1Args::skip), the device-side
search for the chunk boundaries, buffer-global line numbers across chunks.  The full-size
tests run the five BASELINE configurations at the sizes the benchmark uses and check them
through properties that do not need a full CPU pass:

  * the oracle on WINDOWS of the buffer (a few thousand lines each, first / last / random),
    record for record against the records of the full scan for those lines;
  * mode algebra on the whole buffer (count-only == records; FIRST, BEST and ALL agree on the
    matched lines; BEST is the earliest minimum of ALL; FIRST is the first record of ALL);
  * a checksum of checksums: the buffer scanned whole == scanned as newline-aligned shards.
"""
import ctypes as C
import random

import numpy as np
import pytest

from oracle.pyoracle import SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_FAIL, SQ_FIRST, SQ_IGNORE

from .test_gpu_parity import fastq_like, long_line_buffer, make_buffer, rand_pattern

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


class DevBuf:
    """A host buffer copied to the device (16-byte aligned, padded)."""

    def __init__(self, B, data: bytes):
        self.L = B.lib()
        self.n = len(data)
        self.ptr = self.L.sqbDeviceAlloc(self.n + 64)
        assert self.ptr, B.last_error()
        if self.n:
            arr = np.frombuffer(data, dtype=np.uint8)
            assert self.L.sqbMemcpyH2D(self.ptr, arr.ctypes.data, self.n) == 0

    def free(self):
        self.L.sqbDeviceFree(self.ptr)


def rec_rows(recs):
    return [(int(r["line"]) + 1, int(r["start"]), int(r["end"]), int(r["dist"])) for r in recs]


def check_large(B, oracle, pattern, tau, buf, opt, keep_lines=False):
    sq = B.Seeq(pattern, tau)
    keys = sq.keys
    eng = B.Engine.borrowed(sq.engine())
    exp, nl, nm = oracle.buffer_scan(buf, keys, tau, opt)
    exp = [tuple(int(x) for x in row) for row in exp]
    d = DevBuf(B, buf)
    flags = B.SQB_FASTA if buf[:1] == b">" else 0
    st = eng.scan_device_large(d.ptr, d.n, opt | flags | (B.SQB_KEEP_LINES if keep_lines else 0))
    got = rec_rows(eng.host_records())
    assert (st.nbytes, st.nlines, st.nmatched) == (len(buf), nl, nm), (pattern, tau, opt)
    assert got == exp, (pattern, tau, opt)
    if keep_lines:
        starts = eng.host_line_starts()
        # counted lines: every line of the buffer, minus FASTA headers
        pos, want = 0, []
        for ln in buf.split(b"\n")[:-1] if buf.endswith(b"\n") else buf.split(b"\n"):
            if not (flags and ln[:1] == b">"):
                want.append(pos)
            pos += len(ln) + 1
        assert starts.tolist() == want
    # the records kept in HBM are the same records
    st3 = eng.scan_device_large(d.ptr, d.n, opt | flags | B.SQB_DEVICE_RESULTS)
    assert (st3.nlines, st3.nmatched, st3.nrecs) == (nl, nm, len(exp))
    assert rec_rows(eng.device_records_all()) == exp and eng.host_records().size == 0
    # count-only agrees
    st2 = eng.scan_device_large(d.ptr, d.n, opt | flags | B.SQB_COUNT_ONLY)
    assert (st2.nlines, st2.nmatched) == (nl, nm)
    d.free()
    sq.close()


@pytest.mark.parametrize("mrange", [(3, 12), (20, 32), (40, 64), (90, 120)])
def test_large_small_chunks_ragged(B, oracle, mrange, matcher, monkeypatch):
    """Ragged lines, ~3-6 MB, 1 MiB chunks: every chunk but the first starts at an odd address."""
    monkeypatch.setenv("SEEQ_B200_DEVICE_CHUNK_MB", "1")
    rng = random.Random(mrange[0] * 977)
    for it in range(3):
        pattern = rand_pattern(rng, *mrange)
        keys, _ = oracle.parse(pattern)
        tau = rng.randint(0, min(len(keys) - 1, 2 + len(keys) // 12))
        alphabet = ["ACGT", "ACGTN", "ACGTNXacgu-"][it % 3]
        buf = make_buffer(rng, keys, tau, 30000 + 5000 * it, 200, alphabet, final_newline=it != 1)
        assert len(buf) > (2 << 20)
        for mo, nd in ((SQ_FIRST, SQ_FAIL), (SQ_BEST, SQ_CONVERT), (SQ_ALL, SQ_IGNORE), (SQ_ALL, SQ_FAIL)):
            check_large(B, oracle, pattern, tau, buf, mo | nd, keep_lines=(mo == SQ_FIRST))


def test_large_small_chunks_fasta_and_fastq(B, oracle, monkeypatch):
    monkeypatch.setenv("SEEQ_B200_DEVICE_CHUNK_MB", "1")
    rng = random.Random(4242)
    pattern = "GATCGGAAGAGC"
    keys, _ = oracle.parse(pattern)
    # FASTA: headers are skipped and not counted
    lines = []
    for r in range(40000):
        if r % 3 == 0:
            lines.append(">seq%d GATCGGAAGAGC" % r)
        lines.append("".join(rng.choice("ACGT") for _ in range(rng.randint(0, 120))) +
                     (pattern if rng.random() < 0.3 else ""))
    fasta = ("\n".join(lines) + "\n").encode()
    assert len(fasta) > (2 << 20)
    for mo in (SQ_FIRST, SQ_BEST, SQ_ALL):
        check_large(B, oracle, pattern, 2, fasta, mo, keep_lines=(mo == SQ_BEST))
    # FASTQ-like: the line filter drops three lines out of four
    for filt in ("1", "2"):
        monkeypatch.setenv("SEEQ_B200_FILTER", filt)
        fq = fastq_like(rng, keys, 2, 24000, 150)
        assert len(fq) > (2 << 20)
        for mo in (SQ_FIRST, SQ_ALL):
            check_large(B, oracle, pattern, 2, fq, mo)


def test_large_small_chunks_long_lines(B, oracle, monkeypatch):
    """10-kb lines with 1 MiB chunks: segment cuts inside chunks that start mid-vector."""
    monkeypatch.setenv("SEEQ_B200_DEVICE_CHUNK_MB", "1")
    rng = random.Random(31337)
    pattern = "".join(rng.choice("ACGT") for _ in range(40))
    keys, _ = oracle.parse(pattern)
    buf = long_line_buffer(rng, keys, 4, 1100, 10000, junk=5)
    assert len(buf) > (2 << 20)
    for mo in (SQ_FIRST, SQ_BEST, SQ_ALL):
        check_large(B, oracle, pattern, 4, buf, mo | SQ_CONVERT)


def test_large_one_line_longer_than_a_chunk(B, oracle, monkeypatch):
    """A window without any newline is extended to the next one."""
    monkeypatch.setenv("SEEQ_B200_DEVICE_CHUNK_MB", "1")
    rng = random.Random(5)
    pattern = "ACGTTGCAAC"
    keys, _ = oracle.parse(pattern)
    parts = [make_buffer(rng, keys, 1, 4000, 150, "ACGT"),
             ("".join(rng.choice("ACGT") for _ in range(2_600_000)) + pattern + "\n").encode(),
             make_buffer(rng, keys, 1, 9000, 150, "ACGT")]
    buf = b"".join(parts)
    for mo in (SQ_FIRST, SQ_BEST):
        check_large(B, oracle, pattern, 1, buf, mo)


# ---------------------------------------------------------------------------------------
# BASELINE.json's configurations at full size
# ---------------------------------------------------------------------------------------
def _fixed_pattern(seed, n):
    rng = np.random.default_rng(seed)
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


FULL = {
    "cfg1": dict(pattern="GATCGGAAGAGC", tau=2, nd=SQ_FAIL, reads=1_000_000,
                 gen=dict(seed=1, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=102, max_edits=2)),
    "cfg2": dict(pattern="A[CG]TNNGATC", tau=1, nd=SQ_FAIL, reads=10_000_000,
                 gen=dict(seed=2, line_len=150, n_per_1024=5)),
    "cfg3": dict(pattern=_fixed_pattern(3, 40), tau=4, nd=SQ_FAIL, reads=100_000,
                 gen=dict(seed=3, line_len=10_000, plant=_fixed_pattern(3, 40), plant_per_1024=1024, max_edits=4)),
    "cfg4": dict(pattern=_fixed_pattern(4, 100), tau=8, nd=SQ_CONVERT, reads=10_000_000,
                 gen=dict(seed=4, line_len=250, plant=_fixed_pattern(4, 100), plant_per_1024=102, max_edits=8,
                          junk_per_1024=1)),
    # one GPU's share of the 100 GB read set: 12.5 GB of FASTQ-like records
    "cfg5": dict(pattern="GATCGGAAGAGC", tau=2, nd=SQ_FAIL, reads=39_800_000,
                 gen=dict(seed=5, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=307, max_edits=2, fastq=True)),
}


def _checksum(recs, line_base=0):
    """Order-independent 64-bit checksum of a record array."""
    if recs.size == 0:
        return 0
    with np.errstate(over="ignore"):
        h = (recs["line"].astype(np.uint64) + np.uint64(line_base)) * np.uint64(0x9E3779B97F4A7C15)
        h ^= recs["start"].astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
        h ^= recs["end"].astype(np.uint64) * np.uint64(0x165667B19E3779F9)
        h ^= (recs["dist"].astype(np.uint64) + np.uint64(1)) * np.uint64(0x27D4EB2F165667C5)
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        return int(np.bitwise_xor.reduce(h)) ^ (int(recs.size) << 1)


@pytest.mark.parametrize("cfg", sorted(FULL))
def test_full_size_properties(B, oracle, cfg):
    w = FULL[cfg]
    L = B.lib()
    g = B.make_gen(**w["gen"])
    reads = w["reads"]
    rec_bytes = L.sqbGenBytes(C.byref(g), 0, 1)
    nbytes = rec_bytes * reads
    lines_per_read = 4 if w["gen"].get("fastq") else 1
    d_text = L.sqbDeviceAlloc(nbytes + 64)
    assert d_text, B.last_error()
    try:
        assert L.sqbGenDevice(C.byref(g), 0, reads, d_text, None) == 0, B.last_error()
        sq = B.Seeq(w["pattern"], w["tau"])
        keys = sq.keys
        eng = B.Engine.borrowed(sq.engine())
        res = {}
        for name, mo in (("first", SQ_FIRST), ("best", SQ_BEST), ("all", SQ_ALL)):
            st = eng.scan_device_large(d_text, nbytes, mo | w["nd"])
            recs = eng.host_records()
            assert st.nbytes == nbytes and st.nlines == reads * lines_per_read
            assert st.nrecs == recs.size
            stc = eng.scan_device_large(d_text, nbytes, mo | w["nd"] | B.SQB_COUNT_ONLY)
            assert (stc.nlines, stc.nmatched) == (st.nlines, st.nmatched)
            if mo == SQ_ALL:
                assert stc.nrecs == st.nrecs
            res[name] = (st, recs)
            # records are in file order: by line, then by end
            key = recs["line"].astype(np.uint64) << np.uint64(32) | recs["end"].astype(np.uint64)
            assert np.all(key[1:] > key[:-1]) if mo == SQ_ALL else np.all(np.diff(recs["line"].astype(np.int64)) > 0)
            assert np.all(recs["dist"] <= w["tau"]) and np.all(recs["start"] <= recs["end"])

        # ---- mode algebra ------------------------------------------------------------
        (sf, rf), (sb, rb), (sa, ra) = res["first"], res["best"], res["all"]
        assert sf.nmatched == sb.nmatched == sa.nmatched == rf.size == rb.size
        assert np.array_equal(rf["line"], rb["line"])
        lines_all, first_idx = np.unique(ra["line"], return_index=True)
        assert np.array_equal(lines_all, rf["line"])
        # FIRST is the first ALL record of its line
        for f in ("start", "end", "dist"):
            assert np.array_equal(ra[f][first_idx], rf[f])
        # BEST: the smallest distance among the ALL records of the line, the earliest on a tie
        mind = np.minimum.reduceat(ra["dist"], first_idx)
        assert np.array_equal(mind, rb["dist"])
        group = np.repeat(np.arange(lines_all.size), np.diff(np.append(first_idx, ra.size)))
        is_min = ra["dist"] == mind[group]
        first_min = np.full(lines_all.size, ra.size, dtype=np.int64)
        np.minimum.at(first_min, group[is_min], np.nonzero(is_min)[0])
        assert np.array_equal(ra["end"][first_min], rb["end"])
        assert np.array_equal(ra["start"][first_min], rb["start"])

        # ---- the oracle on windows of the buffer -------------------------------------
        rng = random.Random(1000 + sorted(FULL).index(cfg))
        win_reads = max(8, min(3000, (6 << 20) // rec_bytes))
        firsts = [0, reads - win_reads] + [rng.randrange(0, reads - win_reads) for _ in range(3)]
        host = np.empty(win_reads * rec_bytes, dtype=np.uint8)
        for r0 in firsts:
            assert L.sqbMemcpyD2H(host.ctypes.data, d_text + r0 * rec_bytes, host.size) == 0
            l0, l1 = r0 * lines_per_read, (r0 + win_reads) * lines_per_read
            for name, mo in (("first", SQ_FIRST), ("best", SQ_BEST), ("all", SQ_ALL)):
                exp, nl, nm = oracle.buffer_scan(host, keys, w["tau"], mo | w["nd"])
                recs = res[name][1]
                lo, hi = np.searchsorted(recs["line"], [l0, l1])
                got = np.stack([recs["line"][lo:hi].astype(np.uint64) - np.uint64(l0) + np.uint64(1),
                                recs["start"][lo:hi], recs["end"][lo:hi], recs["dist"][lo:hi]],
                               axis=1).astype(np.uint64) if hi > lo else np.zeros((0, 4), np.uint64)
                assert nl == l1 - l0
                assert np.array_equal(got, np.asarray(exp, dtype=np.uint64).reshape(-1, 4)), (cfg, name, r0)

        # ---- checksum of checksums: whole == newline-aligned shards -------------------
        # the record multiset, line numbers rebased: xor of per-record hashes is additive over shards
        # (shards start wherever a record starts: not 16-byte aligned)
        def body(recs, lb=0):
            return _checksum(recs, lb) ^ (int(recs.size) << 1)
        acc, base, tot_matched = 0, 0, 0
        for k in range(4):
            a, b = reads * k // 4, reads * (k + 1) // 4
            st = eng.scan_device_large(d_text + a * rec_bytes, (b - a) * rec_bytes, SQ_BEST | w["nd"])
            acc ^= body(eng.host_records(), base)
            base += st.nlines
            tot_matched += st.nmatched
        assert (base, tot_matched) == (sb.nlines, sb.nmatched)
        assert acc == body(rb)
        sq.close()
    finally:
        L.sqbDeviceFree(d_text)


def test_repeated_scans_replay_a_graph(B, oracle):
    """A scan that repeats the one before it is captured into a CUDA graph and replayed: same
    results as the call-by-call scan, and a replay follows the CONTENT of the buffer (the graph
    holds addresses and sizes, not data)."""
    rng = random.Random(77)
    pattern, tau = "GATCGGAAGAGC", 2
    keys, _ = oracle.parse(pattern)
    sq = B.Seeq(pattern, tau)
    eng = B.Engine.borrowed(sq.engine())
    L = B.lib()
    bufs = [make_buffer(rng, keys, tau, 40000, 150, "ACGTN") for _ in range(2)]
    n = min(len(b) for b in bufs)
    bufs = [b[:n - 1] + b"\n" for b in bufs]
    d = DevBuf(B, bufs[0])
    for mo in (SQ_BEST, SQ_ALL):
        for rep in range(5):
            data = bufs[rep >= 3]                      # same address and size, other bytes
            if rep == 3:
                arr = np.frombuffer(data, dtype=np.uint8)
                assert L.sqbMemcpyH2D(d.ptr, arr.ctypes.data, n) == 0
            exp, nl, nm = oracle.buffer_scan(data, keys, tau, mo)
            st = eng.scan_device(d.ptr, n, mo)
            got = eng.fetch_records(st.nrecs)
            assert (st.nlines, st.nmatched) == (nl, nm)
            assert rec_rows(got) == [tuple(int(x) for x in row) for row in exp], (mo, rep)
        arr = np.frombuffer(bufs[0], dtype=np.uint8)
        assert L.sqbMemcpyH2D(d.ptr, arr.ctypes.data, n) == 0
    # two scans in flight, both slots replaying
    for rep in range(4):
        eng.scan_device_issue(0, d.ptr, n, SQ_BEST)
        eng.scan_device_issue(1, d.ptr, n, SQ_BEST)
        a, b = eng.scan_device_wait(0), eng.scan_device_wait(1)
        assert (a.nlines, a.nmatched, a.nrecs) == (b.nlines, b.nmatched, b.nrecs)
    d.free()
    sq.close()
