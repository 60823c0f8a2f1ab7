"""SQB_FASTQ: record-aware scans of 4-line FASTQ (the part in front of the hot path).
Only the sequence line of every record is matched; every line still counts.  The reference scans
every line, so the expectation is the oracle's result for the buffer with the records of the other
lines removed -- every mode, every -x mode, the bit-sliced and the word-parallel matchers, one batch
and chunked (host and device text; chunks cut at record boundaries), and a pattern set."""
import random

import numpy as np
import pytest

from oracle.pyoracle import SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_FAIL, SQ_FIRST, SQ_IGNORE

from .test_gpu_large import DevBuf
from .test_gpu_parity import plant, rand_pattern

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


def fastq(rng, keys, tau, nreads, maxlen, tricky=True):
    """Well-formed 4-line records.  Quality strings use the full Phred+33 range, so they start with
    '@' or '+' now and then and hold A, C, G, T, N; ids repeat the pattern; some reads are empty."""
    inst = "".join("ACGT"[[i for i in range(4) if k >> i & 1][0]] if k & 15 else "A" for k in keys)
    out = []
    for r in range(nreads):
        n = rng.choice([0, rng.randint(1, 12), rng.randint(30, maxlen), rng.randint(30, maxlen)])
        seq = [rng.choice("ACGT") for _ in range(n)]
        if n > 20 and rng.random() < 0.5:
            at = rng.randrange(n)
            seq[at:at] = list(plant(rng, keys, tau))
        if rng.random() < 0.05 and seq:
            seq[rng.randrange(len(seq))] = rng.choice("NXn-")
        qual = [chr(rng.randint(33, 74)) for _ in range(len(seq))]
        if tricky and qual and rng.random() < 0.3:
            qual[0] = rng.choice("@+")
        if tricky and len(qual) > len(inst) and rng.random() < 0.3:
            at = rng.randrange(len(qual) - len(inst))
            qual[at:at + len(inst)] = list(inst)                # a hit the reference would report
        out += ["@r%d %s" % (r, inst if rng.random() < 0.2 else ""), "".join(seq), "+", "".join(qual)]
    return ("\n".join(out) + "\n").encode()


def expect(oracle, buf, keys, tau, opt):
    exp, nl, _ = oracle.buffer_scan(buf, keys, tau, opt)
    exp = np.asarray(exp, dtype=np.uint64).reshape(-1, 4)
    keep = (exp[:, 0] - 1) % 4 == 1
    exp = exp[keep]
    return exp, nl, len(np.unique(exp[:, 0]))


def rows(r):
    return np.stack([r["line"].astype(np.uint64) + 1, r["start"], r["end"], r["dist"]], axis=1).astype(np.uint64) \
        if r.size else np.zeros((0, 4), np.uint64)


@pytest.mark.parametrize("mrange", [(4, 12), (20, 32), (40, 60), (130, 140)])
def test_fastq_records_every_path(B, oracle, mrange, matcher, monkeypatch):
    monkeypatch.setenv("SEEQ_B200_CHUNK_MB", "1")
    monkeypatch.setenv("SEEQ_B200_DEVICE_CHUNK_MB", "1")
    rng = random.Random(mrange[0] * 31 + 7)
    pattern = rand_pattern(rng, *mrange)
    keys, _ = oracle.parse(pattern)
    tau = rng.randint(0, min(2 + len(keys) // 12, (len(keys) - 1) // 3))
    buf = fastq(rng, keys, tau, 24000, 150)
    assert len(buf) > (2 << 20)
    sq = B.Seeq(pattern, tau)
    eng = B.Engine.borrowed(sq.engine())
    d = DevBuf(B, buf)
    dropped = 0
    for mo in (SQ_FIRST, SQ_BEST, SQ_ALL):
        for nd in (SQ_FAIL, SQ_CONVERT, SQ_IGNORE):
            opt = mo | nd
            exp, nl, nm = expect(oracle, buf, keys, tau, opt)
            full, _, _ = oracle.buffer_scan(buf, keys, tau, opt)
            dropped += len(full) - len(exp)
            # one batch through the libseeq-level entry, chunked host text, chunked device text, one device batch
            st = B.StatsT()
            got = sq.batch(buf, opt | B.SQB_FASTQ, B.SQ_ANY, st)
            assert (st.nlines, st.nmatched) == (nl, nm) and np.array_equal(rows(got), exp), (pattern, tau, opt, "batch")
            st = eng.scan_host(buf, opt | B.SQB_FASTQ)
            assert (st.nlines, st.nmatched) == (nl, nm) and np.array_equal(rows(eng.host_records()), exp), (pattern, tau, opt, "host")
            st = eng.scan_device_large(d.ptr, d.n, opt | B.SQB_FASTQ)
            assert (st.nlines, st.nmatched) == (nl, nm) and np.array_equal(rows(eng.host_records()), exp), (pattern, tau, opt, "large")
            st = eng.scan_device(d.ptr, d.n, opt | B.SQB_FASTQ)
            assert (st.nlines, st.nmatched, st.nrecs) == (nl, nm, len(exp)), (pattern, tau, opt, "device")
            assert np.array_equal(rows(eng.fetch_records(st.nrecs)), exp)
            st = eng.scan_device(d.ptr, d.n, opt | B.SQB_FASTQ | B.SQB_COUNT_ONLY)
            assert (st.nlines, st.nmatched) == (nl, nm)
    assert dropped > 0          # the quality strings and ids did hold hits the reference reports
    d.free()
    sq.close()


def test_fastq_small_buffers_and_pattern_set(B, oracle, monkeypatch):
    monkeypatch.setenv("SEEQ_B200_CHUNK_MB", "1")
    rng = random.Random(11)
    pats = []
    for mm in ((8, 12), (20, 30), (5, 6)):
        p = rand_pattern(rng, *mm)
        keys, _ = oracle.parse(p)
        pats.append((p, keys, rng.randint(0, (len(keys) - 1) // 3)))
    small = fastq(rng, pats[0][1], pats[0][2], 300, 120)            # below the bit-sliced threshold
    big = fastq(rng, pats[0][1], pats[0][2], 24000, 150)
    for buf in (small, big, b""):
        mp = B.Multi([k for _, k, _ in pats], [t for _, _, t in pats])
        for mo in (SQ_BEST, SQ_ALL):
            stats = mp.scan_host(buf, mo | B.SQB_FASTQ)
            for i, (p, keys, tau) in enumerate(pats):
                exp, nl, nm = expect(oracle, buf, keys, tau, mo)
                assert (stats[i].nlines, stats[i].nmatched) == (nl, nm), (i, p, mo, len(buf))
                assert np.array_equal(rows(mp.records(i)), exp), (i, p, mo, len(buf))
        mp.close()


def test_fastq_chunks_need_record_boundaries(B, oracle, monkeypatch):
    """Text that is not FASTQ cannot be cut into record-aligned chunks: the scan says so."""
    monkeypatch.setenv("SEEQ_B200_CHUNK_MB", "1")
    rng = random.Random(5)
    buf = ("\n".join("".join(rng.choice("ACGT") for _ in range(100)) for _ in range(30000)) + "\n").encode()
    eng = B.Engine(bytes([1, 2, 4, 8, 1, 2]), 1)
    with pytest.raises(RuntimeError, match="record boundary"):
        eng.scan_host(buf, SQ_FIRST | B.SQB_FASTQ)
    # one batch needs no boundary: lines 1 mod 4 of the buffer
    exp, nl, nm = expect(oracle, buf[:500000], bytes([1, 2, 4, 8, 1, 2]), 1, SQ_FIRST)
    d = DevBuf(B, buf[:500000])
    st = eng.scan_device(d.ptr, d.n, SQ_FIRST | B.SQB_FASTQ)
    assert (st.nlines, st.nmatched) == (nl, nm)
    d.free()
    eng.close()
