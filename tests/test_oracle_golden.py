"""Pin the oracle against the reference's own known-answer vectors.

Every expected value below is a golden vector held by the reference test-suite
(/root/reference/test/testset.c, lines cited per test) or its fixture
test/testdata.txt (three lines, reproduced here byte for byte).
CPU only.
"""
import numpy as np
import pytest

from oracle.pyoracle import (SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_FIRST, SQ_IGNORE)

# test/testdata.txt of the reference (79 bytes, od -c verified)
TESTDATA = (b"GTATGTACCACAGATGTCGATCGAC\n"
            b"TCTATCATCCGTACTCTGATCTCAT\n"
            b"RCACAGATCACAGATCACAGRATCAC\n")


def recs(oracle, text, pattern, tau, options):
    keys, err = oracle.parse(pattern)
    assert keys is not None, err
    r = oracle.string_match(text, keys, tau, options)
    return [tuple(int(x) for x in row[1:]) for row in r]


def test_distance_sequence_catg(oracle):
    # testset.c:546-616: CATG, tau=1 over ATCCTCATGA
    d, mtm = oracle.distances("ATCCTCATGA", "CATG", 1)
    assert d == [2, 2, 2, 2, 2, 2, 2, 1, 0, 1]
    assert mtm == [2, 1, 2, 2, 1, 2, 1, 0, 0, 0]


def test_distance_sequence_aaaa(oracle):
    # testset.c:626-685: AAAA, tau=1 over ATTAAAT (memory-capped DFA, same values)
    d, mtm = oracle.distances("ATTAAAT", "AAAA", 1)
    assert d == [2, 2, 2, 2, 2, 1, 1]
    assert mtm == [2, 2, 3, 2, 1, 0, 0]


@pytest.mark.parametrize("pattern,keys", [
    ("AaAaAaAa", [1] * 8), ("CcCcCcCc", [2] * 8), ("GgGgGgGg", [4] * 8),
    ("TtTtTtTt", [8] * 8), ("NnNnNnNn", [31] * 8),
    ("Nn[]Nn[]NnN[]n", [31] * 8),
    ("[GATC][gatc][GaTc][gAtC]", [15] * 4),
    ("ACTGA", [1, 2, 8, 4, 1]),          # testset.c:776-780
    ("ACG[AT]", [1, 2, 4, 9]),           # testset.c:794-797
])
def test_parse_golden(oracle, pattern, keys):
    # testset.c:719-752
    k, err = oracle.parse(pattern)
    assert err == 0 and list(k) == keys


@pytest.mark.parametrize("pattern,err", [
    ("[GATCgatc", 5), ("A]", 3), ("[ATG[C]]", 2), ("Z", 4),      # testset.c:754-757
    ("ACT[A[AG]", 2), ("ACT[AG]T]A", 3), ("ACHT[AG]", 4), ("ACT[AG]A[TG", 5),  # :815-826
    ("CACAG[AT", 5),                                              # :1211-1212
])
def test_parse_errors(oracle, pattern, err):
    k, e = oracle.parse(pattern)
    assert k is None and e == err


def test_string_first_best_all(oracle):
    # testset.c:941-987
    text = "TGACTGATGACGTAGTCTACGATCGATCAGTCA"
    assert recs(oracle, text, "GATC", 1, SQ_FIRST) == [(1, 4, 1)]
    assert recs(oracle, text, "GATC", 1, SQ_BEST) == [(20, 24, 0)]
    assert recs(oracle, text, "GATC", 1, SQ_ALL) == [
        (1, 4, 1), (5, 9, 1), (8, 11, 1), (14, 17, 1), (20, 24, 0), (24, 28, 0),
        (29, 32, 1)]


def test_overlapping(oracle):
    # testset.c:991-1030
    assert recs(oracle, "GAAGAAG", "GAAG", 0, SQ_ALL) == [(0, 4, 0), (3, 7, 0)]
    assert recs(oracle, "GAAGAAG", "GAAG", 1, SQ_ALL) == [(0, 4, 0), (3, 7, 0)]
    assert recs(oracle, "GAAGACG", "GAAG", 1, SQ_ALL) == [(0, 4, 0), (3, 7, 1)]


def scan(oracle, pattern, tau, options):
    keys, _ = oracle.parse(pattern)
    r, nl, nm = oracle.buffer_scan(TESTDATA, keys, tau, options)
    return [tuple(int(x) for x in row) for row in r], nl, nm


def test_file_first(oracle):
    # testset.c:835-856: ATCG tau=1, SQ_FIRST / SQ_MATCH
    r, nl, nm = scan(oracle, "ATCG", 1, SQ_FIRST)
    assert r == [(1, 2, 5, 1), (2, 3, 7, 1)] and nl == 3 and nm == 2


def test_file_best(oracle):
    # testset.c:862-880: TGTC tau=1, SQ_BEST
    r, _, _ = scan(oracle, "TGTC", 1, SQ_BEST)
    assert r[:2] == [(1, 14, 18, 0), (2, 2, 6, 1)]
    # testset.c:903-914: CACAGAT tau=1, SQ_BEST, line 1 hit, line 2 none
    r, _, _ = scan(oracle, "CACAGAT", 1, SQ_BEST)
    assert r[0] == (1, 8, 15, 0) and all(x[0] != 2 for x in r)


def test_file_nomatch_lines(oracle):
    # testset.c:886-896: CACAGAT tau=1, default options: lines 2 and 3 do not match
    r, nl, nm = scan(oracle, "CACAGAT", 1, 0)
    assert sorted({x[0] for x in r}) == [1] and nl == 3 and nm == 1


def test_file_counts(oracle):
    # testset.c:922-931: ATC tau=0: COUNTLINES (FIRST) = 2, COUNTMATCH (ALL) = 4
    _, _, nm = scan(oracle, "ATC", 0, SQ_FIRST)
    assert nm == 2
    r, _, _ = scan(oracle, "ATC", 0, SQ_ALL)
    assert len(r) == 4


def test_cli_vectors(oracle):
    """The record content behind the 15 CLI strings of testset.c:1077-1207."""
    # Test 2 (:1098-1102)  "1 8-14 0 ..." : printed end is end-1
    r, _, _ = scan(oracle, "CACAGAT", 0, SQ_FIRST)
    assert r == [(1, 8, 15, 0)]
    # Test 3 (:1105-1111) compact tau=3 "1:8-14:0\n2:8-11:3\n"
    r, _, _ = scan(oracle, "CACAGAT", 3, SQ_FIRST)
    assert r == [(1, 8, 15, 0), (2, 8, 12, 3)]
    # Test 7.1 (:1146-1150) -x 1 adds line 3 "CACAGAT"
    r, _, _ = scan(oracle, "CACAGAT", 3, SQ_FIRST | SQ_CONVERT)
    assert [x[0] for x in r] == [1, 2, 3]
    assert TESTDATA.split(b"\n")[2][r[2][1]:r[2][2]] == b"CACAGAT"
    # Test 7.2 (:1153-1164) CTCAT tau=1: best "CTCAT" vs first "CTAT" on line 2
    line2 = TESTDATA.split(b"\n")[1]
    r, _, _ = scan(oracle, "CTCAT", 1, SQ_BEST)
    assert [line2[x[1]:x[2]] for x in r if x[0] == 2] == [b"CTCAT"]
    r, _, _ = scan(oracle, "CTCAT", 1, SQ_FIRST)
    assert [line2[x[1]:x[2]] for x in r if x[0] == 2] == [b"CTAT"]
    # Test 8/9 (:1167-1181) prefix / suffix at tau=3
    r, _, _ = scan(oracle, "CACAGAT", 3, SQ_FIRST)
    lines = TESTDATA.split(b"\n")
    assert [lines[x[0] - 1][:x[1]] for x in r] == [b"GTATGTAC", b"TCTATCAT"]
    assert [lines[x[0] - 1][x[2]:] for x in r] == [b"GTCGATCGAC", b"ACTCTGATCTCAT"]
    # Test 11 (:1184-1192) -x 2 tau=0 first: lines 1 and 3
    r, _, _ = scan(oracle, "CACAGAT", 0, SQ_FIRST | SQ_IGNORE)
    assert [(x[0], lines[x[0] - 1][x[1]:x[2]]) for x in r] == [
        (1, b"CACAGAT"), (3, b"CACAGAT")]
    # Test 10 (:1195-1200) -a -x 1
    r, _, _ = scan(oracle, "CACAGAT", 0, SQ_ALL | SQ_CONVERT)
    assert [(x[0], lines[x[0] - 1][x[1]:x[2]]) for x in r] == [
        (1, b"CACAGAT"), (3, b"CACAGAT"), (3, b"CACAGAT")]
    # Test 12 (:1203-1207) -a -x 2: the ignored 'R' sits inside the last match
    r, _, _ = scan(oracle, "CACAGAT", 0, SQ_ALL | SQ_IGNORE)
    assert [(x[0], lines[x[0] - 1][x[1]:x[2]]) for x in r] == [
        (1, b"CACAGAT"), (3, b"CACAGAT"), (3, b"CACAGAT"), (3, b"CACAGRAT")]


def test_python_module_vectors(oracle):
    # test/python_lib_test.py:13-35 (matchPrefix / matchSuffix use SQ_BEST)
    keys, _ = oracle.parse("CGCTAATTAATGGAAT")
    assert len(oracle.string_match("ATGCTGATGCTGGGGG", keys, 3, SQ_BEST | SQ_IGNORE)) == 0
    text = "GGGGCGCTAATAATGGAATGGGG"
    r = oracle.string_match(text, keys, 3, SQ_BEST | SQ_IGNORE)
    assert len(r) == 1
    s, e = int(r[0][1]), int(r[0][2])
    assert text[:e] == "GGGGCGCTAATAATGGAAT" and text[:s] == "GGGG"
    assert text[s:] == "CGCTAATAATGGAATGGGG" and text[e:] == "GGGG"


def test_line_rules(oracle):
    # SURVEY 3.4 [probed]: '\r' is an ordinary illegal byte, empty lines count,
    # a final line without '\n' is still a line.
    keys, _ = oracle.parse("CACAGAT")
    r, nl, nm = oracle.buffer_scan(b"ACACAGAT\r\n\nCACAGAT", keys, 0, SQ_FIRST)
    assert [int(x[0]) for x in r] == [1, 3] and nl == 3 and nm == 2
    # FASTA: headers are neither counted nor matched (seeq.c:367-377)
    r, nl, nm = oracle.buffer_scan(b">CACAGAT\nCACAGAT\n>x\nTTTT\nCACAGAT\n", keys, 0, SQ_FIRST)
    assert [int(x[0]) for x in r] == [1, 3] and nl == 3
    # not FASTA when the first byte is not '>' : '>' lines are plain illegal text
    r, nl, nm = oracle.buffer_scan(b"CACAGAT\n>CACAGAT\n", keys, 0, SQ_FIRST)
    assert [int(x[0]) for x in r] == [1] and nl == 2


def test_free_end_reverse_pass(oracle):
    # SURVEY 7 "hard parts": GATC tau=1 on TTTGAATCTTTT -> 5-7 ("ATC"), not 3-7
    assert recs(oracle, "TTTGAATCTTTT", "GATC", 1, SQ_FIRST)[0][:2] == (5, 8)


def test_aaaa_quirk(oracle):
    # AA on AAAA reports 2 matches, not 3 (match-flag rule)
    assert len(recs(oracle, "AAAA", "AA", 0, SQ_ALL)) == 2
