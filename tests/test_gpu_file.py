"""GPU: the file driver (seeqOpen / seeqFileMatch / seeqClose) and the CLI
formatter seeq() through the C-ABI, against the reference's known-answer vectors
(/root/reference/test/testset.c, lines cited) and against the oracle on seeded
multi-chunk inputs."""
import ctypes as C
import json
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle.pyoracle import SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_FIRST, SQ_IGNORE

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))

# test/testdata.txt of the reference (79 bytes)
TESTDATA = (b"GTATGTACCACAGATGTCGATCGAC\n"
            b"TCTATCATCCGTACTCTGATCTCAT\n"
            b"RCACAGATCACAGATCACAGRATCAC\n")

SQ_ANY, SQ_MATCH, SQ_NOMATCH, SQ_COUNTLINES, SQ_COUNTMATCH = 0, 1, 2, 3, 4


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


@pytest.fixture()
def testdata(tmp_path):
    p = tmp_path / "testdata.txt"
    p.write_bytes(TESTDATA)
    return str(p)


class File:
    def __init__(self, B, path):
        self.L = B.lib()
        self.f = self.L.seeqOpen(path.encode() if path else None)
        assert self.f, B.seeqerr()

    def match(self, sq, mopt, fopt):
        return self.L.seeqFileMatch(self.f, sq.sq, mopt, fopt)

    @property
    def line(self):
        return self.f.contents.line

    def close(self):
        assert self.L.seeqClose(self.f) == 0


def hits(sq):
    out = []
    while True:
        m = sq.L.seeqMatchIter(sq.sq)
        if not m:
            return out
        out.append((m.contents.start, m.contents.end, m.contents.dist))


def test_filematch_vectors(B, testdata):
    """testset.c:829-939 (test_seeqFileMatch)."""
    L = B.lib()
    # :835-856  ATCG tau=1, SQ_FIRST, SQ_MATCH: lines 1 and 2, then end of file
    sq = B.Seeq("ATCG", 1)
    f = File(B, testdata)
    assert f.match(sq, SQ_FIRST, SQ_MATCH) == 1
    assert f.line == 1 and sq.sq.contents.hits == 1
    assert L.seeqGetString(sq.sq) == b"GTATGTACCACAGATGTCGATCGAC"
    assert hits(sq) == [(2, 5, 1)]
    assert f.match(sq, SQ_FIRST, SQ_MATCH) == 1
    assert f.line == 2 and hits(sq) == [(3, 7, 1)]
    assert L.seeqGetString(sq.sq) == b"TCTATCATCCGTACTCTGATCTCAT"
    assert f.match(sq, SQ_FIRST, SQ_MATCH) == 0
    f.close()
    sq.close()
    # :862-880  TGTC tau=1 SQ_BEST
    sq = B.Seeq("TGTC", 1)
    f = File(B, testdata)
    assert f.match(sq, SQ_BEST, SQ_MATCH) == 1 and f.line == 1 and hits(sq) == [(14, 18, 0)]
    assert f.match(sq, SQ_BEST, SQ_MATCH) == 1 and f.line == 2 and hits(sq) == [(2, 6, 1)]
    f.close()
    sq.close()
    # :886-896  CACAGAT tau=1, SQ_NOMATCH: lines 2 and 3
    sq = B.Seeq("CACAGAT", 1)
    f = File(B, testdata)
    assert f.match(sq, SQ_FIRST, SQ_NOMATCH) == 1 and f.line == 2
    assert L.seeqGetString(sq.sq) == b"TCTATCATCCGTACTCTGATCTCAT"
    assert f.match(sq, SQ_FIRST, SQ_NOMATCH) == 1 and f.line == 3
    assert f.match(sq, SQ_FIRST, SQ_NOMATCH) == 0
    f.close()
    # :903-914  SQ_ANY walks line by line
    f = File(B, testdata)
    assert f.match(sq, SQ_BEST, SQ_ANY) == 1 and f.line == 1 and hits(sq) == [(8, 15, 0)]
    assert f.match(sq, SQ_BEST, SQ_ANY) == 1 and f.line == 2 and sq.sq.contents.hits == 0
    f.close()
    sq.close()
    # :922-931  ATC tau=0: COUNTLINES = 2, COUNTMATCH = 4
    sq = B.Seeq("ATC", 0)
    f = File(B, testdata)
    assert f.match(sq, 0, SQ_COUNTLINES) == 2
    f.close()
    f = File(B, testdata)
    assert f.match(sq, 0, SQ_COUNTMATCH) == 4
    assert f.line == 3
    # :933-939  a NULL file pointer is error 10
    f.f.contents.fdi = None
    assert f.match(sq, 0, SQ_ANY) == -1 and B.seeqerr() == 10
    L.seeqClose(f.f)
    sq.close()


DEFAULT_ARGS = dict(showdist=0, showpos=0, showline=0, printline=1, matchonly=0, count=0, compact=0, dist=0,
                    verbose=0, endline=0, prefix=0, split=0, invert=0, best=0, non_dna=0, all=0, memory=0)


def run_seeq(pattern, path, **kw):
    args = dict(DEFAULT_ARGS)
    args.update(kw)
    r = subprocess.run([sys.executable, os.path.join(HERE, "cli_helper.py"),
                        json.dumps({"pattern": pattern, "input": path, "args": args})],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, stdin=subprocess.DEVNULL, timeout=300)
    tail = r.stderr.decode().strip().splitlines()[-1]
    assert tail.startswith("RC="), r.stderr.decode()[-800:]
    rc, err = (int(x.split("=")[1]) for x in tail.split())
    return r.stdout.decode("latin-1"), rc, err


# (kwargs, pattern, expected stdout): testset.c:1077-1207
CLI_VECTORS = [
    (dict(), "CACAGAT", "GTATGTACCACAGATGTCGATCGAC\n"),                                              # test 1
    (dict(showdist=1, showpos=1, showline=1), "CACAGAT", "1 8-14 0 GTATGTACCACAGATGTCGATCGAC\n"),    # test 2
    (dict(compact=1, dist=3), "CACAGAT", "1:8-14:0\n2:8-11:3\n"),                                    # test 3
    (dict(count=1), "CACAGAT", "1\n"),                                                               # test 4
    (dict(invert=1), "CACAGAT", "TCTATCATCCGTACTCTGATCTCAT\nRCACAGATCACAGATCACAGRATCAC\n"),          # test 5
    (dict(invert=1, showline=1), "CACAGAT", "2 TCTATCATCCGTACTCTGATCTCAT\n3 RCACAGATCACAGATCACAGRATCAC\n"),
    (dict(matchonly=1, dist=3), "CACAGAT", "CACAGAT\nCCGT\n"),                                       # test 7
    (dict(matchonly=1, dist=3, non_dna=1), "CACAGAT", "CACAGAT\nCCGT\nCACAGAT\n"),                   # test 7.1
    (dict(matchonly=1, dist=1, best=1), "CTCAT", "CTCAT\n"),                                         # test 7.2
    (dict(matchonly=1, dist=1), "CTCAT", "CTAT\n"),
    (dict(printline=0, prefix=1, dist=3), "CACAGAT", "GTATGTAC\nTCTATCAT\n"),                        # test 8
    (dict(printline=0, endline=1, dist=3), "CACAGAT", "GTCGATCGAC\nACTCTGATCTCAT\n"),                # test 9
    (dict(printline=0, showline=1, non_dna=2, matchonly=1), "CACAGAT", "1 CACAGAT\n3 CACAGAT\n"),    # test 11
    (dict(printline=0, showline=1, non_dna=1, matchonly=1, all=1), "CACAGAT",
     "1 CACAGAT\n3 CACAGAT\n3 CACAGAT\n"),                                                           # test 10
    (dict(printline=0, showline=1, non_dna=2, matchonly=1, all=1), "CACAGAT",
     "1 CACAGAT\n3 CACAGAT\n3 CACAGAT\n3 CACAGRAT\n"),                                               # test 12
]


@pytest.mark.parametrize("k", range(len(CLI_VECTORS)))
def test_seeq_cli_vectors(testdata, k):
    kw, pattern, expected = CLI_VECTORS[k]
    out, rc, _ = run_seeq(pattern, testdata, **kw)
    assert rc == 0
    assert out == expected


def test_seeq_cli_errors(testdata):
    # testset.c:1209-1227
    assert run_seeq("CACAG[AT", testdata)[1:] == (1, 5)
    assert run_seeq("CACAGAT", testdata, dist=7)[1:] == (1, 9)
    assert run_seeq("CACAGAT", "invented.txt")[1:] == (1, 2)
    assert run_seeq("CACAGAT", testdata, dist=-1)[1:] == (1, 1)


def test_iterator_over_many_chunks(B, oracle, tmp_path, monkeypatch, matcher):
    """seeqFileMatch hands out lines one call at a time although the file is matched
    in batches: force tiny chunks so that the walk crosses many chunk boundaries."""
    monkeypatch.setenv("SEEQ_B200_FILE_CHUNK_MB", "1")
    rng = random.Random(5)
    g = B.make_gen(seed=11, line_len=97, plant="GATCGGAAGAGC", plant_per_1024=300, max_edits=2, n_per_1024=8)
    buf = B.gen_host(g, 40000)                      # 3.9 MB -> 4 chunks
    path = tmp_path / "reads.txt"
    buf.tofile(path)
    keys, _ = oracle.parse("GATCGGAAGAGC")
    for opt in (SQ_FIRST, SQ_BEST | SQ_CONVERT, SQ_ALL | SQ_IGNORE):
        exp, nl, nm = oracle.buffer_scan(buf, keys, 2, opt)
        exp = [tuple(int(x) for x in r) for r in exp]
        sq = B.Seeq("GATCGGAAGAGC", 2)
        f = File(B, str(path))
        got = []
        fopt = rng.choice([SQ_ANY, SQ_MATCH])
        while f.match(sq, opt, fopt) > 0:
            got += [(f.line, *h) for h in hits(sq)]
        assert f.line == nl
        f.close()
        assert got == exp, opt
        # counts over the same file
        f = File(B, str(path))
        assert f.match(sq, opt, SQ_COUNTLINES) == nm
        f.close()
        f = File(B, str(path))
        n_all = len(oracle.buffer_scan(buf, keys, 2, (opt & 0xC) | SQ_ALL)[0])
        assert f.match(sq, opt, SQ_COUNTMATCH) == n_all
        f.close()
        sq.close()


def test_sharded_scan_equals_unsharded(B, oracle, matcher):
    """Newline-aligned byte ranges scanned independently (one per GPU on the box; here
    one after another on cuda:0) + line-base prefix == the unsharded scan."""
    from seeq_b200 import shard
    g = B.make_gen(seed=21, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=200, max_edits=2)
    buf = B.gen_host(g, 30000)
    sq = B.Seeq("GATCGGAAGAGC", 2)
    whole = sq.batch(buf, SQ_ALL, SQ_ANY)
    for world in (2, 3, 8):
        parts, base = [], 0
        for r in range(world):
            b, e = B.shard_range(buf, r, world)
            assert (b, e) == shard.shard_ranges(buf, world)[r]
            st = B.StatsT()
            recs = sq.batch(buf[b:e], SQ_ALL, SQ_ANY, st)
            parts.append(shard.gather_records(recs, base))
            base += st.nlines
        got = np.concatenate(parts, axis=0)
        exp = np.stack([whole["line"], whole["start"], whole["end"], whole["dist"]], axis=1).astype(np.int64)
        assert np.array_equal(got, exp), world
    sq.close()


def test_long_lines_through_the_line_keeping_scan_are_cut(B, oracle, monkeypatch):
    """seeqFileMatch scans with SQB_KEEP_LINES (it serves the lines one call at a time).  Lines of 10 kb are cut into
    segments there too: the line starts handed back are the entries of the segment list that open a line."""
    import numpy as np
    monkeypatch.setenv("SEEQ_B200_MATCHER", "bitslice")      # (lifts the size thresholds: 400 lines take the production kernels)
    monkeypatch.setenv("SEEQ_B200_CUTS", "2")                 # ... and with them the on-demand trigger: cut from the first scan
    pattern = "ACGTTGCAAGCTTAGGCATCGATCGGATCAGCTAGCTAGC"
    g = B.make_gen(seed=3, line_len=10000, plant=pattern, plant_per_1024=1024, max_edits=4)
    buf = B.gen_host(g, 400)
    sq = B.Seeq(pattern, 4)
    for mo in (B.SQ_FIRST, B.SQ_BEST, B.SQ_ALL):
        for it in range(2):                                   # the first scan of an engine finds the long lines and repeats itself
            st = B.StatsT()
            recs = sq.batch(buf, mo | B.SQB_KEEP_LINES, B.SQ_ANY, st)
            exp, nl, nm = oracle.buffer_scan(buf, sq.keys, 4, mo)
            got = [(int(r["line"]) + 1, int(r["start"]), int(r["end"]), int(r["dist"])) for r in recs]
            assert got == [tuple(int(x) for x in row) for row in exp]
            assert (st.nlines, st.nmatched) == (nl, nm)
            assert st.path & 4, st.path                        # SQB_PATH_CUTS
            starts = B.Engine.borrowed(sq.engine()).host_line_starts()
            nlpos = np.flatnonzero(np.frombuffer(buf, np.uint8) == 10)
            assert np.array_equal(starts, np.concatenate([[0], nlpos[:-1] + 1]).astype(np.uint64))
    sq.close()
