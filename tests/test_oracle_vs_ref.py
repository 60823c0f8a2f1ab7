"""Differential fuzz: oracle (our restatement) vs the compiled reference.

Runs where oracle/_ref/libseeq_ref.so exists (built in this container from
/root/reference by oracle/Makefile; it travels to the GPU box as a prebuilt
artefact).  Recipe follows SURVEY.md 8(c): short random patterns with classes
and N, small tau, texts over an alphabet with N, lower case, U and illegal
bytes, all match modes x non-DNA modes (+ SQ_STREAM).  CPU only.
"""
import random

import numpy as np
import pytest

from oracle.pyoracle import (SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_COUNTLINES,
                             SQ_COUNTMATCH, SQ_FAIL, SQ_FIRST, SQ_IGNORE,
                             SQ_STREAM)

MATCH = [SQ_FIRST, SQ_BEST, SQ_ALL]
NONDNA = [SQ_FAIL, SQ_CONVERT, SQ_IGNORE]


def rand_pattern(rng, mmax):
    m = rng.randint(1, mmax)
    out = []
    for _ in range(m):
        r = rng.random()
        if r < 0.1:
            out.append("N")
        elif r < 0.25:
            out.append("[" + "".join(rng.sample("ACGT", rng.randint(1, 3))) + "]")
        else:
            out.append(rng.choice("ACGTacgu"))
    return "".join(out), m


def rand_text(rng, nmax, alphabet):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(0, nmax)))


def check(oracle, reference, pattern, tau, text, options):
    keys, err = oracle.parse(pattern)
    n_ref, r_ref = reference.string_match(pattern, tau, text, options)
    if keys is None:
        assert n_ref == -100 - err
        return
    if tau >= len(keys):
        assert n_ref == -100 - 9
        return
    r = oracle.string_match(text, keys, tau, options)
    assert n_ref == len(r), (pattern, tau, text, options)
    assert np.array_equal(r[:, 1:], r_ref), (pattern, tau, text, options, r, r_ref)


def test_fuzz_short(oracle, reference):
    rng = random.Random(20260101)
    alphabets = ["ACGT", "ACGTN", "ACGTNXacgu-", "AC"]
    for it in range(6000):
        pattern, m = rand_pattern(rng, 9)
        tau = rng.randint(0, min(3, m - 1))
        text = rand_text(rng, 60, alphabets[it % len(alphabets)])
        for mo in MATCH:
            for nd in NONDNA:
                check(oracle, reference, pattern, tau, text, mo | nd)


def test_fuzz_stream(oracle, reference):
    rng = random.Random(7)
    for it in range(1500):
        pattern, m = rand_pattern(rng, 7)
        tau = rng.randint(0, min(2, m - 1))
        text = rand_text(rng, 50, "ACGTN\nX")
        for mo in MATCH:
            for nd in NONDNA:
                check(oracle, reference, pattern, tau, text, mo | nd | SQ_STREAM)
                check(oracle, reference, pattern, tau, text, mo | nd)


def test_fuzz_long_patterns(oracle, reference):
    # m up to 130 (multi-word territory of the CUDA path), tau up to 12
    rng = random.Random(99)
    for it in range(150):
        pattern, m = rand_pattern(rng, 130)
        tau = rng.randint(0, min(12, m - 1))
        # text = noise + mutated copy of the pattern + noise
        core = [rng.choice("ACGT") if c in "N[]" else c for c in pattern if c not in "[]"]
        for _ in range(rng.randint(0, tau + 2)):
            if core:
                p = rng.randrange(len(core))
                op = rng.random()
                if op < 0.4:
                    core[p] = rng.choice("ACGT")
                elif op < 0.7:
                    del core[p]
                else:
                    core.insert(p, rng.choice("ACGT"))
        text = rand_text(rng, 300, "ACGT") + "".join(core) + rand_text(rng, 300, "ACGTN")
        for mo in MATCH:
            check(oracle, reference, pattern, tau, text, mo | rng.choice(NONDNA))


def test_buffer_scan_matches_reference(oracle, reference):
    rng = random.Random(5)
    for it in range(300):
        pattern, m = rand_pattern(rng, 8)
        tau = rng.randint(0, min(2, m - 1))
        keys, _ = oracle.parse(pattern)
        nlines = rng.randint(0, 12)
        lines = [rand_text(rng, 40, "ACGTNX") for _ in range(nlines)]
        if it % 3 == 0:            # FASTA-shaped input
            lines = [(">h%d" % k if k % 2 == 0 else s) for k, s in enumerate(lines)]
        buf = "\n".join(lines)
        if it % 2 == 0 and nlines:
            buf += "\n"
        if it % 7 == 0:
            buf = buf.replace("X", "\r", 1)
        for mo in MATCH:
            for nd in NONDNA:
                r, nl, nm = oracle.buffer_scan(buf.encode(), keys, tau, mo | nd)
                rr, rnl, rnm = reference.buffer_scan(buf.encode(), pattern, tau, mo | nd)
                assert (nl, nm) == (rnl, rnm), (pattern, tau, buf)
                assert np.array_equal(r, rr), (pattern, tau, buf, mo | nd)
        # count modes of seeqFileMatch reduce to the scan
        r_all, _, _ = oracle.buffer_scan(buf.encode(), keys, tau, SQ_ALL)
        _, _, nm = oracle.buffer_scan(buf.encode(), keys, tau, SQ_FIRST)
        assert reference.buffer_count(buf.encode(), pattern, tau, 0, SQ_COUNTMATCH) == len(r_all)
        assert reference.buffer_count(buf.encode(), pattern, tau, 0, SQ_COUNTLINES) == nm


def test_reductions_hold_on_reference(oracle, reference):
    """FIRST == ALL[0]; BEST == first minimal-distance element of ALL."""
    rng = random.Random(11)
    for it in range(3000):
        pattern, m = rand_pattern(rng, 8)
        tau = rng.randint(0, min(3, m - 1))
        text = rand_text(rng, 80, "ACGTN")
        n_all, r_all = reference.string_match(pattern, tau, text, SQ_ALL)
        n_first, r_first = reference.string_match(pattern, tau, text, SQ_FIRST)
        n_best, r_best = reference.string_match(pattern, tau, text, SQ_BEST)
        if n_all <= 0:
            assert n_first == n_all and n_best == n_all
            continue
        assert n_first == 1 and np.array_equal(r_first[0], r_all[0])
        k = int(np.argmin(r_all[:, 2]))
        assert n_best == 1 and np.array_equal(r_best[0], r_all[k])


def test_segment_warmup_rule(oracle):
    """Events are reproduced exactly when a line is cut into segments that are
    re-started m+2*tau+4 automaton-visible bytes early (basis of the
    segment-parallel long-line kernel)."""
    rng = random.Random(3)
    for it in range(1500):
        pattern, m = rand_pattern(rng, 12)
        tau = rng.randint(0, min(3, m - 1))
        keys, _ = oracle.parse(pattern)
        text = rand_text(rng, 200, "ACGTNX" if it % 2 else "ACGT")
        for nd in (SQ_CONVERT, SQ_IGNORE):
            full = oracle.string_match(text, keys, tau, SQ_ALL | nd)
            seg = rng.randint(1, 64)
            cut = oracle.string_match_segmented(text, keys, tau, SQ_ALL | nd, seg,
                                                len(keys) + 2 * tau + 4)
            assert np.array_equal(full, cut), (pattern, tau, text, seg)
