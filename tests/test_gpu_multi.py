"""GPU parity of the pattern-set scans (sqbMultiScanHost / sqbMultiScanDevice): several patterns
over ONE pass of the text -- K1 and the bit-plane pack run once, every pattern runs its own
matcher and finishing kernels -- must give, for every pattern, exactly what the oracle gives for
that pattern alone (records, counted lines, matched lines), in every mode."""
import random

import numpy as np
import pytest

from oracle.pyoracle import SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_FAIL, SQ_FIRST, SQ_IGNORE

from .test_gpu_large import DevBuf
from .test_gpu_parity import fastq_like, make_buffer, rand_pattern

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


def pattern_set(rng, oracle, spec):
    """spec: list of (mmin, mmax); -> [(pattern, keys, tau)]"""
    out = []
    for mmin, mmax in spec:
        p = rand_pattern(rng, mmin, mmax)
        keys, _ = oracle.parse(p)
        out.append((p, keys, rng.randint(0, min(2 + len(keys) // 12, (len(keys) - 1) // 3))))
    return out


def check_set(B, oracle, pats, buf, opt, device):
    mp = B.Multi([k for _, k, _ in pats], [t for _, _, t in pats])
    flags = B.SQB_FASTA if buf[:1] == b">" else 0
    if device:
        d = DevBuf(B, buf)
        stats = mp.scan_device(d.ptr, d.n, opt | flags)
    else:
        stats = mp.scan_host(buf, opt | flags)
    counts = []
    for i, (p, keys, tau) in enumerate(pats):
        exp, nl, nm = oracle.buffer_scan(buf, keys, tau, opt)
        exp = np.asarray(exp, dtype=np.uint64).reshape(-1, 4)
        counts.append((nl, nm))
        st = stats[i]
        assert (st.nbytes, st.nlines, st.nmatched, st.nrecs) == (len(buf), nl, nm, len(exp)), (i, p, tau, opt)
        r = mp.records(i)
        got = np.stack([r["line"].astype(np.uint64) + 1, r["start"], r["end"], r["dist"]], axis=1).astype(np.uint64) \
            if r.size else np.zeros((0, 4), np.uint64)
        assert np.array_equal(got, exp), (i, p, tau, opt)
    # counts only
    if device:
        stats = mp.scan_device(d.ptr, d.n, opt | flags | B.SQB_COUNT_ONLY)
        d.free()
    else:
        stats = mp.scan_host(buf, opt | flags | B.SQB_COUNT_ONLY)
    for i, (p, keys, tau) in enumerate(pats):
        assert (stats[i].nlines, stats[i].nmatched) == counts[i], (i, p)
    mp.close()


@pytest.mark.parametrize("device", [False, True], ids=["host", "device"])
def test_pattern_set_ragged(B, oracle, device, monkeypatch):
    monkeypatch.setenv("SEEQ_B200_CHUNK_MB", "1")
    monkeypatch.setenv("SEEQ_B200_DEVICE_CHUNK_MB", "1")
    rng = random.Random(20261017 + device)
    # the >128 pattern has no bit-sliced kernel: it rides on the leader's line starts alone,
    # and comes FIRST so that the set has to pick another leader
    pats = pattern_set(rng, oracle, [(130, 150), (8, 12), (3, 6), (20, 32), (40, 64), (90, 120)])
    for it, alphabet in enumerate(["ACGT", "ACGTNXacgu-"]):
        buf = make_buffer(rng, pats[1][1], pats[1][2], 26000, 200, alphabet, final_newline=it == 0)
        # plant a few instances of the other patterns too
        lines = buf.split(b"\n")
        for _, keys, _ in pats:
            inst = bytes(b"ACGT"[[i for i in range(4) if k >> i & 1][0]] if k & 15 else 65 for k in keys)
            for _ in range(300):
                j = rng.randrange(len(lines))
                at = rng.randrange(len(lines[j]) + 1)
                lines[j] = lines[j][:at] + inst + lines[j][at:]
        buf = b"\n".join(lines)
        assert len(buf) > (2 << 20)
        for mo, nd in ((SQ_FIRST, SQ_FAIL), (SQ_BEST, SQ_CONVERT), (SQ_ALL, SQ_IGNORE), (SQ_ALL, SQ_FAIL)):
            check_set(B, oracle, pats, buf, mo | nd, device)


@pytest.mark.parametrize("filt", ["1", "2"])
def test_pattern_set_fastq_filter(B, oracle, filt, monkeypatch):
    """FASTQ-like input: the leader filters with the smallest m - tau of the set."""
    monkeypatch.setenv("SEEQ_B200_CHUNK_MB", "1")
    monkeypatch.setenv("SEEQ_B200_FILTER", filt)
    rng = random.Random(99 + int(filt))
    pats = pattern_set(rng, oracle, [(20, 30), (4, 6), (10, 14), (40, 60)])
    fq = fastq_like(rng, pats[0][1], pats[0][2], 24000, 150)
    assert len(fq) > (2 << 20)
    for mo in (SQ_FIRST, SQ_BEST, SQ_ALL):
        check_set(B, oracle, pats, fq, mo, device=False)
    check_set(B, oracle, pats, fq, SQ_BEST, device=True)


def test_pattern_set_small_and_single(B, oracle):
    """A set of one pattern, and buffers below the size at which the bit-sliced matcher engages."""
    rng = random.Random(3)
    pats = pattern_set(rng, oracle, [(8, 12)])
    buf = make_buffer(rng, pats[0][1], pats[0][2], 500, 100, "ACGTN")
    check_set(B, oracle, pats, buf, SQ_ALL, device=False)
    pats = pattern_set(rng, oracle, [(8, 12), (30, 40), (5, 7)])
    check_set(B, oracle, pats, buf, SQ_BEST, device=False)
    check_set(B, oracle, pats, buf, SQ_FIRST | SQ_CONVERT, device=True)
    check_set(B, oracle, pats, b"", SQ_FIRST, device=False)


def test_pattern_set_leader_repeats_its_scan(B, oracle, monkeypatch):
    """Lines of ~6 bytes: more lines per byte than the engine guesses, so the leader's first scan of
    every chunk runs out of room and is repeated with exact sizes (stats.reruns) -- the other patterns
    of the set have read a front that was cut short and must go again."""
    monkeypatch.setenv("SEEQ_B200_CHUNK_MB", "1")
    monkeypatch.setenv("SEEQ_B200_MATCHER", "bitslice")
    rng = random.Random(808)
    pats = pattern_set(rng, oracle, [(4, 5), (3, 4), (6, 8)])
    lines = ["".join(rng.choice("ACGT") for _ in range(rng.randint(0, 12))) for _ in range(330000)]
    buf = ("\n".join(lines) + "\n").encode()
    assert len(buf) > (2 << 20)
    mp = B.Multi([k for _, k, _ in pats], [t for _, _, t in pats])
    stats = mp.scan_host(buf, SQ_BEST)
    assert sum(st.reruns for st in stats) > 0
    for i, (p, keys, tau) in enumerate(pats):
        exp, nl, nm = oracle.buffer_scan(buf, keys, tau, SQ_BEST)
        exp = np.asarray(exp, dtype=np.uint64).reshape(-1, 4)
        r = mp.records(i)
        got = np.stack([r["line"].astype(np.uint64) + 1, r["start"], r["end"], r["dist"]], axis=1).astype(np.uint64)
        assert (stats[i].nlines, stats[i].nmatched) == (nl, nm), (i, p)
        assert np.array_equal(got, exp), (i, p)
    mp.close()
