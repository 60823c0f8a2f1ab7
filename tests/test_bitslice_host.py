"""CPU: the bit-sliced automaton (seeq_b200/csrc/sqb_bitslice.h, the core of the CUDA
matcher K2) compiled for the host and driven exactly like the kernel drives it, against
the oracle: same events (line, end, dist) in every match mode x non-DNA mode."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle.pyoracle import SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_FAIL, SQ_FIRST, SQ_IGNORE

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "seeq_b200", "csrc")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("bs") / "host_bitslice.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I" + CSRC,
                    os.path.join(HERE, "host_bitslice.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.bs_host_scan.restype = C.c_long
    L.bs_host_scan.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int, C.c_int, C.c_int,
                               C.POINTER(C.c_uint64), C.c_long]
    L.bs_host_scan_cut.restype = C.c_long
    L.bs_host_scan_cut.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                   C.c_size_t, C.POINTER(C.c_uint64), C.c_long, C.POINTER(C.c_long)]
    L.bs_host_scan_wm.restype = C.c_long
    L.bs_host_scan_wm.argtypes = L.bs_host_scan_cut.argtypes
    L.bs_parse_cpulist.restype = C.c_int
    L.bs_parse_cpulist.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.bs_fastq_record_start.restype = C.c_long
    L.bs_fastq_record_start.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t]
    return L


def rand_pattern(rng, mmin, mmax, custom=True):
    m = rng.randint(mmin, mmax)
    brackets = [rng.sample("ACGT", rng.randint(2, 3)) for _ in range(rng.randint(1, 5))]     # up to kBsMaxCustom = 6 classes
    out = []
    for _ in range(m):
        r = rng.random()
        if r < 0.08:
            out.append("N")
        elif r < 0.2 and custom:
            out.append("[" + "".join(rng.choice(brackets)) + "]")
        else:
            out.append(rng.choice("ACGTacgu"))
    return "".join(out)


def make_buffer(rng, nlines, maxlen, alphabet, plant, final_newline):
    lines = []
    for _ in range(nlines):
        n = rng.randint(0, maxlen)
        s = [rng.choice(alphabet) for _ in range(n)]
        if rng.random() < 0.5 and n > 0:
            at = rng.randrange(n)
            q = list(plant)
            for _ in range(rng.randint(0, 3)):
                if q:
                    k = rng.randrange(len(q))
                    q[k:k + 1] = rng.choice([[], [rng.choice("ACGT")], [q[k], rng.choice("ACGT")]])
            s[at:at] = q
        lines.append("".join(s))
    buf = "\n".join(lines)
    if final_newline and nlines:
        buf += "\n"
    return buf.encode()


@pytest.mark.parametrize("mrange", [(1, 8), (9, 12), (13, 16), (17, 24), (25, 32), (33, 40), (41, 64),
                                    (65, 80), (81, 96), (97, 104), (105, 112), (113, 128)])
def test_bitsliced_events_equal_oracle(harness, oracle, mrange):
    rng = random.Random(mrange[1] * 31)
    checked = 0
    for it in range(30 if mrange[1] <= 32 else 12):
        pattern = rand_pattern(rng, *mrange)
        keys, _ = oracle.parse(pattern)
        if not keys:
            continue
        tau = rng.randint(0, min(len(keys) - 1, 3 + len(keys) // 8, 14))
        plant = "".join(rng.choice([c for b, c in ((1, "A"), (2, "C"), (4, "G"), (8, "T")) if k & b] or ["A"])
                        for k in keys)
        alphabet = ["ACGT", "ACGTN", "ACGTNXacgu-"][it % 3]
        buf = make_buffer(rng, rng.randint(1, 150), 60 + 3 * len(keys), alphabet, plant, it % 2 == 0)
        out = np.zeros((len(buf) + 64, 3), dtype=np.uint64)
        for mo in (SQ_FIRST, SQ_BEST, SQ_ALL):
            for nd in (SQ_FAIL, SQ_CONVERT, SQ_IGNORE):
                n = harness.bs_host_scan(buf, len(buf), keys, len(keys), tau, mo | nd,
                                         out.ctypes.data_as(C.POINTER(C.c_uint64)), out.shape[0])
                if n == -1:
                    continue              # more than kBsMaxCustom custom classes: not a bit-sliced pattern
                assert n >= 0
                exp, _, _ = oracle.buffer_scan(buf, keys, tau, mo | nd)
                exp = exp[:, [0, 2, 3]]
                assert np.array_equal(out[:n], exp), (pattern, tau, mo, nd, buf[:120])
                checked += 1
    assert checked > (100 if mrange[1] <= 32 else 40)


@pytest.mark.parametrize("mrange,stride,window", [((1, 8), 64, 32), ((4, 16), 96, 48), ((12, 32), 256, 128),
                                                   ((33, 48), 512, 256), ((65, 100), 1024, 512)])
def test_cut_segments_equal_oracle(harness, oracle, mrange, stride, window):
    """Long lines cut into segments with a warm-up of m + 2 tau + 2 bytes reproduce the
    events of the uncut line (the rule the GPU pipeline uses with stride 2048 / window 1024)."""
    rng = random.Random(mrange[1] * 77 + stride)
    checked = cuts = 0
    for it in range(40 if mrange[1] <= 32 else 12):
        pattern = rand_pattern(rng, *mrange)
        keys, _ = oracle.parse(pattern)
        if not keys:
            continue
        tau = rng.randint(0, min(len(keys) - 1, 3 + len(keys) // 8, 14))
        if len(keys) + 2 * tau + 2 > window:
            continue
        plant = "".join(rng.choice([c for b, c in ((1, "A"), (2, "C"), (4, "G"), (8, "T")) if k & b] or ["A"])
                        for k in keys)
        alphabet = ["ACGT", "ACGTN", "ACGTNNX"][it % 3]           # X: a STOP in the middle of a line (SQ_FAIL)
        lines = []
        for _ in range(rng.randint(1, 40)):
            n = rng.choice([rng.randint(0, window), rng.randint(window, 6 * stride)])
            s = [rng.choice(alphabet[:4] if rng.random() < 0.98 else alphabet) for _ in range(n)]
            for _ in range(rng.randint(0, 1 + n // (3 * len(keys) + 20))):
                at = rng.randrange(n + 1)
                q = list(plant)
                for _ in range(rng.randint(0, tau + 1)):
                    if q:
                        k = rng.randrange(len(q))
                        q[k:k + 1] = rng.choice([[], [rng.choice("ACGT")], [q[k], rng.choice("ACGT")]])
                s[at:at] = q
            lines.append("".join(s))
        buf = "\n".join(lines).encode() + (b"\n" if it % 2 == 0 else b"")
        out = np.zeros((len(buf) + 64, 3), dtype=np.uint64)
        ncuts = C.c_long(0)
        for mo in (SQ_FIRST, SQ_BEST, SQ_ALL):
            for nd in (SQ_FAIL, SQ_CONVERT):
                n = harness.bs_host_scan_cut(buf, len(buf), keys, len(keys), tau, mo | nd, stride, window,
                                             out.ctypes.data_as(C.POINTER(C.c_uint64)), out.shape[0], C.byref(ncuts))
                if n == -1:
                    continue
                assert n >= 0
                exp, _, _ = oracle.buffer_scan(buf, keys, tau, mo | nd)
                exp = exp[:, [0, 2, 3]]
                assert np.array_equal(out[:n], exp), (pattern, tau, mo, nd, len(buf))
                checked += 1
                cuts += ncuts.value
    assert checked > 30 and cuts > 100


@pytest.mark.parametrize("mrange,stride,window", [((1, 8), 0, 0), ((9, 12), 0, 0), ((13, 32), 0, 0),
                                                   ((2, 10), 64, 32), ((8, 32), 256, 128)])
def test_nfa_level_automaton_equals_oracle(harness, oracle, mrange, stride, window):
    """tau <= 2: the NFA-level formulation (bs_wm_step) reproduces the oracle's events, on
    whole lines (stride 0) and on lines cut into segments."""
    rng = random.Random(mrange[1] * 13 + stride)
    checked = 0
    for it in range(40):
        pattern = rand_pattern(rng, *mrange)
        keys, _ = oracle.parse(pattern)
        if not keys:
            continue
        tau = rng.randint(0, min(len(keys) - 1, 2))
        if stride and len(keys) + 2 * tau + 2 > window:
            continue
        plant = "".join(rng.choice([c for b, c in ((1, "A"), (2, "C"), (4, "G"), (8, "T")) if k & b] or ["A"])
                        for k in keys)
        alphabet = ["ACGT", "ACGTN", "ACGTNXacgu-"][it % 3]
        if stride:
            lines = []
            for _ in range(rng.randint(1, 40)):
                n = rng.choice([rng.randint(0, window), rng.randint(window, 5 * stride)])
                s = [rng.choice(alphabet[:4] if rng.random() < 0.98 else alphabet) for _ in range(n)]
                for _ in range(rng.randint(0, 1 + n // (3 * len(keys) + 20))):
                    at = rng.randrange(n + 1)
                    q = list(plant)
                    for _ in range(rng.randint(0, tau + 1)):
                        if q:
                            k = rng.randrange(len(q))
                            q[k:k + 1] = rng.choice([[], [rng.choice("ACGT")], [q[k], rng.choice("ACGT")]])
                    s[at:at] = q
                lines.append("".join(s))
            buf = "\n".join(lines).encode() + (b"\n" if it % 2 == 0 else b"")
        else:
            buf = make_buffer(rng, rng.randint(1, 150), 60 + 3 * len(keys), alphabet, plant, it % 2 == 0)
        out = np.zeros((len(buf) + 64, 3), dtype=np.uint64)
        ncuts = C.c_long(0)
        for mo in (SQ_FIRST, SQ_BEST, SQ_ALL):
            for nd in (SQ_FAIL, SQ_CONVERT, SQ_IGNORE):
                n = harness.bs_host_scan_wm(buf, len(buf), keys, len(keys), tau, mo | nd, stride, window,
                                            out.ctypes.data_as(C.POINTER(C.c_uint64)), out.shape[0], C.byref(ncuts))
                if n == -1:
                    continue
                assert n >= 0
                exp, _, _ = oracle.buffer_scan(buf, keys, tau, mo | nd)
                exp = exp[:, [0, 2, 3]]
                assert np.array_equal(out[:n], exp), (pattern, tau, mo, nd, buf[:120])
                checked += 1
    assert checked > 100


def test_fastq_chunk_cuts_are_record_starts(harness):
    """SQB_FASTQ: a chunked scan cuts the text at the last record start at or in front of a
    newline-aligned position (sqb_tables.h: fastq_record_start -- a line that starts with '@' whose
    next-but-one line starts with '+').  On well-formed records with hostile quality strings (they
    start with '@' or '+' and contain both) every cut is a record start, and the nearest one."""
    rng = random.Random(2026)
    recs = []
    for r in range(4000):
        n = rng.choice([0, 1, rng.randint(2, 150)])
        seq = "".join(rng.choice("ACGTN") for _ in range(n))
        qual = [chr(rng.randint(33, 74)) for _ in range(n)]
        if qual and rng.random() < 0.5:
            qual[0] = rng.choice("@+")
        recs.append("@%s\n%s\n+%s\n%s\n" % (rng.choice(["r%d" % r, "@@", "+", ""]), seq, rng.choice(["", "r%d" % r]), "".join(qual)))
    buf = "".join(recs).encode()
    starts, pos = [], 0
    for rec in recs:
        starts.append(pos)
        pos += len(rec)
    starts_set = set(starts)
    line_starts = [0] + [i + 1 for i, b in enumerate(buf) if b == 10 and i + 1 < len(buf)]
    checked = 0
    for _ in range(3000):
        cut = rng.choice(line_starts)
        lo = rng.choice([0, starts[rng.randrange(len(starts))]])
        if lo > cut:
            lo = 0
        q = harness.bs_fastq_record_start(buf, lo, cut, len(buf))
        want = max((s for s in starts if lo <= s <= cut), default=None)
        # the last record of the buffer has no line two on from its quality line to look at: it is
        # recognised by its '+' line like the others; only a cut inside the last TWO lines of the
        # buffer can fall back to the record before
        assert q != -1 and q in starts_set and lo <= q <= cut, (lo, cut, q)
        assert q == want, (lo, cut, q, want)
        checked += 1
    assert checked == 3000
    # text without records: no boundary
    plain = ("\n".join("".join(rng.choice("ACGT") for _ in range(80)) for _ in range(500)) + "\n").encode()
    assert harness.bs_fastq_record_start(plain, 0, 81 * 200, len(plain)) == -1


@pytest.mark.parametrize("text,want", [
    ("0-31,64-95\n", list(range(32)) + list(range(64, 96))),
    ("0\n", [0]), ("3,5,7-9", [3, 5, 7, 8, 9]), ("", []), ("\n", []), ("0-3,2-5\n", [0, 1, 2, 3, 4, 5]),
    ("1020-1030\n", [1020, 1021, 1022, 1023]), ("7-", [])])
def test_cpulist_parser_of_the_numa_placement(harness, text, want):
    """Pinned host memory is allocated on the CPUs of the GPU's NUMA node (sqb_engine.cu: NumaScope);
    the node's sysfs cpulist is parsed by sqb_tables.h: parse_cpulist."""
    buf = C.create_string_buffer(1024)
    n = harness.bs_parse_cpulist(text.encode(), buf, 1024)
    got = [i for i in range(1024) if buf.raw[i] == 1]
    assert (n, got) == (len(want), want)


@pytest.mark.parametrize("mrange", [(33, 64), (65, 128)])
def test_multi_part_pipeline_with_a_skew_of_two(harness, oracle, mrange):
    """The hand-over between the parts of a multi-part automaton does not depend on the parts running
    exactly one column apart: with part p running 2 p columns behind and reading what part p-1 produced
    two iterations ago, the events are the same (the kernels use a skew of 1)."""
    harness.bs_set_skew.argtypes = [C.c_int]
    harness.bs_set_skew.restype = None
    harness.bs_set_skew(2)
    try:
        rng = random.Random(mrange[0] * 13)
        checked = 0
        for it in range(10):
            pattern = rand_pattern(rng, *mrange)
            keys, _ = oracle.parse(pattern)
            tau = rng.randint(0, min(len(keys) - 1, 3 + len(keys) // 8, 14))
            plant = "".join(rng.choice([c for b, c in ((1, "A"), (2, "C"), (4, "G"), (8, "T")) if k & b] or ["A"])
                            for k in keys)
            alphabet = ["ACGT", "ACGTN", "ACGTNXacgu-"][it % 3]
            buf = make_buffer(rng, rng.randint(1, 150), 60 + 3 * len(keys), alphabet, plant, it % 2 == 0)
            out = np.zeros((len(buf) + 64, 3), dtype=np.uint64)
            for mo in (SQ_FIRST, SQ_BEST, SQ_ALL):
                for nd in (SQ_FAIL, SQ_CONVERT, SQ_IGNORE):
                    n = harness.bs_host_scan(buf, len(buf), keys, len(keys), tau, mo | nd,
                                             out.ctypes.data_as(C.POINTER(C.c_uint64)), out.shape[0])
                    if n == -1:
                        continue
                    exp, _, _ = oracle.buffer_scan(buf, keys, tau, mo | nd)
                    assert np.array_equal(out[:n], exp[:, [0, 2, 3]]), (pattern, tau, mo, nd)
                    checked += 1
        assert checked > 30
    finally:
        harness.bs_set_skew(1)
