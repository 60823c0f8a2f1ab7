"""GPU parity: the CUDA path (through the C-ABI) against the oracle.

Bit-exact comparison of records (line, start, end, dist) and counts on seeded
inputs for every match mode x non-DNA mode, single- and multi-word patterns,
plus the known-answer vectors of the reference test-suite through the real
libseeq entry points.
"""
import random

import numpy as np
import pytest

from oracle.pyoracle import (SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_COUNTLINES, SQ_COUNTMATCH, SQ_FAIL,
                             SQ_FIRST, SQ_IGNORE, SQ_STREAM)

pytestmark = pytest.mark.gpu

MATCH = [SQ_FIRST, SQ_BEST, SQ_ALL]
NONDNA = [SQ_FAIL, SQ_CONVERT, SQ_IGNORE]


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


def rand_pattern(rng, mmin, mmax):
    m = rng.randint(mmin, mmax)
    out = []
    for _ in range(m):
        r = rng.random()
        if r < 0.08:
            out.append("N")
        elif r < 0.2:
            out.append("[" + "".join(rng.sample("ACGT", rng.randint(1, 3))) + "]")
        else:
            out.append(rng.choice("ACGTacgu"))
    return "".join(out)


def plant(rng, pattern_keys, tau):
    core = []
    for k in pattern_keys:
        opts = [c for b, c in ((1, "A"), (2, "C"), (4, "G"), (8, "T")) if k & b]
        core.append(rng.choice(opts) if opts else "A")
    for _ in range(rng.randint(0, tau + 1)):
        if not core:
            break
        p = rng.randrange(len(core))
        op = rng.random()
        if op < 0.4:
            core[p] = rng.choice("ACGT")
        elif op < 0.7:
            del core[p]
        else:
            core.insert(p, rng.choice("ACGT"))
    return "".join(core)


def make_buffer(rng, keys, tau, nlines, maxlen, alphabet, final_newline=True):
    lines = []
    for _ in range(nlines):
        n = rng.randint(0, maxlen)
        s = "".join(rng.choice(alphabet) for _ in range(n))
        if rng.random() < 0.4 and n > 0:
            at = rng.randrange(n)
            s = s[:at] + plant(rng, keys, tau) + s[at:]
        lines.append(s)
    buf = "\n".join(lines)
    if final_newline and nlines:
        buf += "\n"
    return buf.encode()


def as_tuples(recs):
    return [(int(r["line"]) + 1, int(r["start"]), int(r["end"]), int(r["dist"])) for r in recs]


def check_buffer(B, oracle, pattern, tau, buf, opt):
    sq = B.Seeq(pattern, tau)
    keys = sq.keys
    okeys, _ = oracle.parse(pattern)
    assert keys == okeys
    exp, nl, nm = oracle.buffer_scan(buf, keys, tau, opt)
    exp = [tuple(int(x) for x in row) for row in exp]
    st = B.StatsT()
    flags = B.SQB_FASTA if buf[:1] == b">" else 0
    got = as_tuples(sq.batch(buf, opt | flags, B.SQ_ANY, st))
    assert (st.nlines, st.nmatched) == (nl, nm), (pattern, tau, opt)
    assert got == exp, (pattern, tau, opt, buf[:200])
    sq.close()


@pytest.mark.parametrize("mrange", [(1, 12), (20, 32), (33, 64), (65, 128), (129, 200)])
def test_buffers_all_modes(B, oracle, mrange, matcher):
    rng = random.Random(mrange[0] * 7919)
    for it in range(12):
        pattern = rand_pattern(rng, *mrange)
        keys, _ = oracle.parse(pattern)
        tau = rng.randint(0, min(len(keys) - 1, 3 + len(keys) // 12))
        alphabet = ["ACGT", "ACGTN", "ACGTNXacgu-"][it % 3]
        buf = make_buffer(rng, keys, tau, rng.randint(1, 700), 90 + 4 * len(keys), alphabet,
                          final_newline=it % 2 == 0)
        for mo in MATCH:
            for nd in NONDNA:
                check_buffer(B, oracle, pattern, tau, buf, mo | nd)


def long_line_buffer(rng, keys, tau, nlines, maxlen, junk):
    """Ragged long lines (the matcher cuts them into segments at text offsets 256 mod
    2048), pattern instances planted at random and right across the cut offsets."""
    lines = []
    for _ in range(nlines):
        n = rng.choice([rng.randint(0, 300), rng.randint(200, 3500), rng.randint(3000, maxlen)])
        lines.append([rng.choice("ACGT") for _ in range(n)])
    buf = bytearray("\n".join("".join(l) for l in lines).encode() + b"\n")
    m = len(keys)
    for a in range(256, len(buf), 2048):
        if rng.random() < 0.7:
            inst = plant(rng, keys, tau).encode()
            at = a - rng.randint(0, m + tau + 1) + rng.randint(0, 3)
            if at >= 0 and at + len(inst) < len(buf) and b"\n" not in buf[at:at + len(inst)]:
                buf[at:at + len(inst)] = inst
    for _ in range(len(buf) // 700):
        inst = plant(rng, keys, tau).encode()
        at = rng.randrange(max(1, len(buf) - len(inst) - 1))
        if b"\n" not in buf[at:at + len(inst)]:
            buf[at:at + len(inst)] = inst
    for _ in range(junk):                       # bytes that end a line (SQ_FAIL), are N (SQ_CONVERT) or invisible
        at = rng.randrange(len(buf))
        if buf[at] != 10:
            buf[at] = rng.choice(b"XN-")
    return bytes(buf)


@pytest.mark.parametrize("cuts", ["on-demand", "always"])
@pytest.mark.parametrize("mrange", [(4, 12), (20, 32), (36, 48), (90, 110)])
def test_long_lines(B, oracle, mrange, matcher, cuts, monkeypatch):
    """Lines of up to 12 kB.  `always`: the bit-sliced matcher cuts them into segments
    (SEEQ_B200_CUTS=2); `on-demand`: the engine's default -- with the thresholds of the
    forced bit-sliced matcher lifted that is the un-cut bit-sliced scan, otherwise the
    word-parallel kernels (the buffers are below the 1 MB the engine wants)."""
    if cuts == "always":
        monkeypatch.setenv("SEEQ_B200_CUTS", "2")
    else:
        monkeypatch.delenv("SEEQ_B200_CUTS", raising=False)
    rng = random.Random(mrange[0] * 131)
    for it in range(4):
        pattern = rand_pattern(rng, *mrange)
        keys, _ = oracle.parse(pattern)
        tau = rng.randint(0, min(len(keys) - 1, 2 + len(keys) // 10))
        buf = long_line_buffer(rng, keys, tau, rng.randint(3, 60), 12000, junk=(0, 6, 40)[it % 3])
        for mo in MATCH:
            for nd in NONDNA:
                check_buffer(B, oracle, pattern, tau, buf, mo | nd)


def test_long_lines_large(B, oracle):
    """More than the 1 MB the engine wants before it goes bit-sliced on its own: 10-kb
    reads (BASELINE config 3 shape) with the engine's default choices -- the first scan
    meets the long lines and is repeated with segment cuts (stats.reruns), the later ones
    cut right away."""
    rng = random.Random(99)
    pattern = "".join(rng.choice("ACGT") for _ in range(40))
    keys, _ = oracle.parse(pattern)
    buf = long_line_buffer(rng, keys, 4, 520, 10000, junk=3)
    assert len(buf) > (1 << 20)
    for mo in MATCH:
        for nd in (SQ_FAIL, SQ_CONVERT):
            check_buffer(B, oracle, pattern, 4, buf, mo | nd)
    sq = B.Seeq(pattern, 4)
    r_all, _, _ = oracle.buffer_scan(buf, keys, 4, SQ_ALL)
    _, _, nm = oracle.buffer_scan(buf, keys, 4, SQ_FIRST)
    assert sq.batch(buf, 0, SQ_COUNTMATCH) == len(r_all)
    assert sq.batch(buf, 0, SQ_COUNTLINES) == nm
    sq.close()


def fastq_like(rng, keys, tau, nreads, maxlen):
    """4-line records (@id, sequence, +, quality); some reads empty, some shorter than the
    pattern, a few long ones; the quality alphabet '#'..'I' contains A, C and G."""
    out = []
    for r in range(nreads):
        n = rng.choice([0, rng.randint(1, 12), rng.randint(30, maxlen), rng.randint(30, maxlen)])
        seq = [rng.choice("ACGT") for _ in range(n)]
        if n > 20 and rng.random() < 0.5:
            at = rng.randrange(n)
            seq[at:at] = list(plant(rng, keys, tau))
        if rng.random() < 0.05 and seq:
            seq[rng.randrange(len(seq))] = rng.choice("NX")
        qual = "".join(chr(rng.randint(ord("#"), ord("I"))) for _ in range(len(seq)))
        out += ["@r%d" % r, "".join(seq), "+", qual]
    return ("\n".join(out) + "\n").encode()


@pytest.mark.parametrize("filt", ["0", "1", "2"])
@pytest.mark.parametrize("mrange", [(3, 9), (10, 30), (40, 60)])
def test_fastq_like_line_filter(B, oracle, mrange, filt, monkeypatch):
    """The line filter (lines with a STOP among their first m - tau bytes are not packed
    into the matcher's tiles): never / decided by the first scan / always."""
    monkeypatch.setenv("SEEQ_B200_MATCHER", "bitslice")
    monkeypatch.setenv("SEEQ_B200_FILTER", filt)
    rng = random.Random(mrange[1] * 17 + int(filt))
    for it in range(4):
        pattern = rand_pattern(rng, *mrange)
        keys, _ = oracle.parse(pattern)
        tau = rng.randint(0, min(len(keys) - 1, 3))
        if it == 3:
            monkeypatch.setenv("SEEQ_B200_CUTS", "2")          # long reads cut into segments, too
        buf = fastq_like(rng, keys, tau, rng.randint(5, 900), 300 if it < 3 else 5000)
        for mo in MATCH:
            for nd in NONDNA:
                check_buffer(B, oracle, pattern, tau, buf, mo | nd)
        sq = B.Seeq(pattern, tau)
        r_all, _, _ = oracle.buffer_scan(buf, keys, tau, SQ_ALL)
        _, _, nm = oracle.buffer_scan(buf, keys, tau, SQ_FIRST)
        assert sq.batch(buf, 0, SQ_COUNTMATCH) == len(r_all)
        assert sq.batch(buf, 0, SQ_COUNTLINES) == nm
        sq.close()


def test_counts(B, oracle, matcher):
    rng = random.Random(4)
    for it in range(10):
        pattern = rand_pattern(rng, 4, 40)
        keys, _ = oracle.parse(pattern)
        tau = rng.randint(0, min(len(keys) - 1, 4))
        buf = make_buffer(rng, keys, tau, 900, 160, "ACGTN")
        sq = B.Seeq(pattern, tau)
        r_all, _, _ = oracle.buffer_scan(buf, keys, tau, SQ_ALL)
        _, _, nm = oracle.buffer_scan(buf, keys, tau, SQ_FIRST)
        assert sq.batch(buf, 0, SQ_COUNTMATCH) == len(r_all)
        assert sq.batch(buf, 0, SQ_COUNTLINES) == nm
        sq.close()


def test_counts_of_long_lines_are_cut(B, oracle, monkeypatch):
    """Count-only scans (seeq -c, SQ_COUNTLINES / SQ_COUNTMATCH) of 10-kb lines take segment cuts too: they run as plain
    scans up to the per-tile sums, so that lines and events are counted per LINE, not per segment."""
    monkeypatch.setenv("SEEQ_B200_MATCHER", "bitslice")
    pattern = "ACGTTGCAAGCTTAGGCATCGATCGGATCAGCTAGCTAGC"
    g = B.make_gen(seed=3, line_len=10000, plant=pattern, plant_per_1024=700, max_edits=4)
    buf = bytes(B.gen_host(g, 300)).replace(b"ACGTACGT", b"ACGT\x00CGT", 40)      # a few STOP bytes: dead segments behind them
    keys, _ = oracle.parse(pattern)
    r_all, _, _ = oracle.buffer_scan(buf, keys, 4, SQ_ALL)
    _, _, nm = oracle.buffer_scan(buf, keys, 4, SQ_FIRST)
    for cuts in ("0", "2"):
        monkeypatch.setenv("SEEQ_B200_CUTS", cuts)
        sq = B.Seeq(pattern, 4)
        for it in range(2):
            st = B.StatsT()
            assert sq.batch(buf, 0, SQ_COUNTMATCH, st) == len(r_all)
            assert bool(st.path & 4) == (cuts == "2"), (cuts, st.path)
            assert sq.batch(buf, 0, SQ_COUNTLINES, st) == nm
            assert bool(st.path & 4) == (cuts == "2"), (cuts, st.path)
        sq.close()


def test_string_api_fuzz(B, oracle):
    rng = random.Random(17)
    for it in range(60):
        pattern = rand_pattern(rng, 1, 10)
        keys, _ = oracle.parse(pattern)
        tau = rng.randint(0, min(3, len(keys) - 1))
        sq = B.Seeq(pattern, tau)
        for _ in range(8):
            text = "".join(rng.choice("ACGTNX\n") for _ in range(rng.randint(0, 70)))
            for mo in MATCH:
                for nd in NONDNA:
                    for stream in (0, SQ_STREAM):
                        opt = mo | nd | stream
                        exp = [tuple(int(x) for x in r[1:]) for r in oracle.string_match(text, keys, tau, opt)]
                        assert sq.string_match(text, opt) == exp, (pattern, tau, text, opt)
        sq.close()


def test_reference_string_vectors(B):
    # /root/reference/test/testset.c:941-1030 through seeqNew/seeqStringMatch/seeqMatchIter
    sq = B.Seeq("GATC", 1)
    text = "TGACTGATGACGTAGTCTACGATCGATCAGTCA"
    assert sq.string_match(text, SQ_FIRST) == [(1, 4, 1)]
    assert sq.string_match(text, SQ_BEST) == [(20, 24, 0)]
    assert sq.string_match(text, SQ_ALL) == [(1, 4, 1), (5, 9, 1), (8, 11, 1), (14, 17, 1),
                                              (20, 24, 0), (24, 28, 0), (29, 32, 1)]
    # stored right-to-left: match[0] is the last one (testset.c:961-987)
    n = sq.L.seeqStringMatch(text.encode(), sq.sq, SQ_ALL)
    assert n == 7 and sq.sq.contents.hits == 7
    m = sq.sq.contents.match
    assert (m[6].start, m[6].end, m[6].dist) == (1, 4, 1)
    assert (m[0].start, m[0].end, m[0].dist) == (29, 32, 1)
    sq.close()
    for tau, text, exp in [(0, "GAAGAAG", [(0, 4, 0), (3, 7, 0)]), (1, "GAAGAAG", [(0, 4, 0), (3, 7, 0)]),
                           (1, "GAAGACG", [(0, 4, 0), (3, 7, 1)])]:
        sq = B.Seeq("GAAG", tau)
        assert sq.string_match(text, SQ_ALL) == exp
        sq.close()
