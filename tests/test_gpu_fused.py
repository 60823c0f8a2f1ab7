"""GPU parity of the fused tokenise + pack kernel (k12_scan_pack, sqb_k12_fused.cuh) and of the matcher that
reads its group planes: ragged lines across tile boundaries, lines beyond the staged overlap (one re-run with
the wide overlap, then the two-kernel path), more line starts per tile than the kernel lists, the line
filter, buffers without a final newline, NUL bytes -- each against the oracle, and the fused path against
the two-kernel path on the same bytes."""
import random

import numpy as np
import pytest

from oracle.pyoracle import SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_FAIL, SQ_FIRST, SQ_IGNORE

pytestmark = pytest.mark.gpu

MATCH = [SQ_FIRST, SQ_BEST, SQ_ALL]
NONDNA = [SQ_FAIL, SQ_CONVERT, SQ_IGNORE]
FUSED, BITSLICE = 2, 1


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


@pytest.fixture(autouse=True)
def lifted_thresholds(monkeypatch):
    monkeypatch.setenv("SEEQ_B200_MATCHER", "bitslice")      # small buffers take the production kernels
    monkeypatch.delenv("SEEQ_B200_FUSED", raising=False)


def ragged(rng, nlines, lo, hi, pattern, alphabet="ACGT", junk=0.0, final_newline=True):
    lines = []
    for _ in range(nlines):
        n = rng.randint(lo, hi)
        s = [rng.choice(alphabet) for _ in range(n)]
        if n > len(pattern) + 2 and rng.random() < 0.5:
            at = rng.randrange(n - len(pattern))
            inst = list(pattern)
            if rng.random() < 0.6:
                inst[rng.randrange(len(inst))] = rng.choice("ACGT")
            s[at:at + len(inst)] = inst
        if junk:
            for k in range(n):
                if rng.random() < junk:
                    s[k] = rng.choice("NnXa-u\x00")
        lines.append("".join(s))
    buf = "\n".join(lines)
    if final_newline:
        buf += "\n"
    return buf.encode("latin-1")


def scan(B, oracle, pattern, tau, buf, opt, extra=0):
    sq = B.Seeq(pattern, tau)
    st = B.StatsT()
    recs = sq.batch(buf, opt | extra, B.SQ_ANY, st)
    exp, nl, nm = oracle.buffer_scan(buf, sq.keys, tau, opt)
    got = [(int(r["line"]) + 1, int(r["start"]), int(r["end"]), int(r["dist"])) for r in recs]
    assert (st.nlines, st.nmatched) == (nl, nm), (pattern, tau, opt, st.path)
    assert got == [tuple(int(x) for x in row) for row in exp], (pattern, tau, opt, st.path)
    sq.close()
    return st


@pytest.mark.parametrize("pattern,tau", [("A[CG]TNNGATC", 1), ("GATCGGAAGAGC", 2), ("ACGTTGCAAGCTTAGGCATCGATC", 3),
                                         ("GATTACAGATTACAGATTACAGATTACAGA", 0)])
def test_ragged_lines_all_modes(B, oracle, pattern, tau):
    rng = random.Random(len(pattern) * 31 + tau)
    plain = pattern.replace("[CG]", "C").replace("N", "A")
    buf = ragged(rng, 2600, 0, 330, plain, junk=0.004)            # ~ 430 KB: a dozen tiles, lines across every boundary
    for mo in MATCH:
        for nd in NONDNA:
            st = scan(B, oracle, pattern, tau, buf, mo | nd)
            assert st.path & BITSLICE and st.path & FUSED, (mo, nd, st.path)
    buf = ragged(rng, 900, 100, 160, plain, final_newline=False)
    st = scan(B, oracle, pattern, tau, buf, SQ_BEST)
    assert st.path & FUSED


def test_lines_beyond_the_overlap(B, oracle):
    """Lines of up to 3000 bytes run past the 512 bytes staged behind a tile: the scan is repeated once with the
    4096-byte overlap; lines of 9000 bytes send the engine to the two-kernel path for good.  Same records."""
    rng = random.Random(5)
    pattern = "GATCGGAAGAGC"
    buf = ragged(rng, 600, 0, 3000, pattern)
    sq = B.Seeq(pattern, 2)
    for it in range(2):
        st = B.StatsT()
        recs = sq.batch(buf, SQ_ALL, B.SQ_ANY, st)
        exp, nl, nm = oracle.buffer_scan(buf, sq.keys, 2, SQ_ALL)
        assert [(int(r["line"]) + 1, int(r["start"]), int(r["end"]), int(r["dist"])) for r in recs] == \
            [tuple(int(x) for x in row) for row in exp]
        # (first pass: the overlap re-run, and possibly one for the plane capacity guess)
        assert st.path & FUSED and (1 <= st.reruns <= 2 if it == 0 else st.reruns == 0), (it, st.path, st.reruns)
    sq.close()
    buf = ragged(rng, 200, 0, 9000, pattern)
    for mo in MATCH:
        st = scan(B, oracle, pattern, 2, buf, mo)
        assert not (st.path & FUSED) and st.reruns >= 1


def test_more_line_starts_than_a_tile_lists(B, oracle):
    rng = random.Random(6)
    buf = ragged(rng, 60000, 0, 12, "ACGTA")                      # ~ 5000 line starts per 32 KiB tile (1024 are listed)
    for mo in MATCH:
        st = scan(B, oracle, "ACGTA", 1, buf, mo | SQ_CONVERT)
        assert not (st.path & FUSED)
    buf = ragged(rng, 30000, 30, 60, "ACGTACG")                   # ~ 700 per tile: listed
    for mo in MATCH:
        st = scan(B, oracle, "ACGTACG", 1, buf, mo)
        assert st.path & FUSED


def test_line_filter(B, oracle, monkeypatch):
    monkeypatch.setenv("SEEQ_B200_FILTER", "2")
    monkeypatch.setenv("SEEQ_B200_FUSED", "2")               # filtered scans take the two-kernel path by default (it is faster)
    g = B.make_gen(seed=5, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=307, max_edits=2, fastq=True)
    buf = B.gen_host(g, 6000)
    for mo in MATCH:
        st = scan(B, oracle, "GATCGGAAGAGC", 2, buf, mo)
        assert st.path & FUSED and st.path & 8, st.path
    rng = random.Random(8)
    mixed = ragged(rng, 3000, 0, 200, "GATCGGAAGAGC", alphabet="ACGT", junk=0.02)      # dead lines scattered about
    for mo in MATCH:
        st = scan(B, oracle, "GATCGGAAGAGC", 2, mixed, mo)
        assert st.path & FUSED


def test_fused_equals_two_kernel_path(B, monkeypatch):
    g = B.make_gen(seed=2, line_len=150, n_per_1024=5)
    buf = B.gen_host(g, 200000)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("SEEQ_B200_FUSED", fused)
        for pattern, tau, opt in (("A[CG]TNNGATC", 1, SQ_BEST), ("GATCGGAAGAGC", 2, SQ_ALL), ("TTGACAGCTAGCTCAGTCCT", 2, SQ_FIRST)):
            sq = B.Seeq(pattern, tau)
            st = B.StatsT()
            recs = sq.batch(buf, opt, B.SQ_ANY, st)
            assert bool(st.path & FUSED) == (fused == "1")
            out.setdefault((pattern, opt), []).append((recs.copy(), st.nlines, st.nmatched, st.nrecs))
            n = sq.batch(buf, opt, B.SQ_COUNTLINES)
            assert n == st.nmatched
            sq.close()
    for key, (a, b) in out.items():
        assert a[1:] == b[1:], key
        assert np.array_equal(a[0], b[0]), key


def test_device_chunks_with_an_unaligned_start(B, oracle):
    """sqbScanDeviceLarge hands K12 chunks that start `skip` < 16 bytes into their aligned address."""
    import ctypes as C
    g = B.make_gen(seed=9, line_len=97, plant="GATCGGAAGAGC", plant_per_1024=200, max_edits=2)
    buf = B.gen_host(g, 40000)
    L = B.lib()
    d = L.sqbDeviceAlloc(buf.size + 64)
    assert L.sqbMemcpyH2D(d, buf.ctypes.data, buf.size) == 0
    sq = B.Seeq("GATCGGAAGAGC", 2)
    eng = B.Engine.borrowed(sq.engine())
    import os
    os.environ["SEEQ_B200_DEVICE_CHUNK_MB"] = "1"
    try:
        st = eng.scan_device_large(d, buf.size, SQ_BEST)
    finally:
        del os.environ["SEEQ_B200_DEVICE_CHUNK_MB"]
    recs = eng.host_records()
    exp, nl, nm = oracle.buffer_scan(buf, sq.keys, 2, SQ_BEST)
    assert (st.nlines, st.nmatched) == (nl, nm)
    assert [(int(r["line"]) + 1, int(r["start"]), int(r["end"]), int(r["dist"])) for r in recs] == \
        [tuple(int(x) for x in row) for row in exp]
    assert st.path & FUSED
    L.sqbDeviceFree(d)
    sq.close()


def test_many_bracket_classes_stay_bit_sliced(B, oracle):
    """Up to six bracket classes (other than single bases, N and "any") have an Eq slot of their own: a pattern
    with four or six of them runs on the bit-sliced kernels like any other (it used to fall back to the
    word-parallel ones beyond two); a seventh class does fall back -- same records either way."""
    rng = random.Random(11)
    buf = ragged(rng, 3000, 20, 200, "ACGTTAGGCATT", junk=0.003)
    for pattern, sliced in (("A[CG]T[AT]NG[ACG]TC[GT]A", True), ("[AC][AG][AT][CG][CT][GT]ACGT", True),
                            ("[AC][AG][AT][CG][CT][GT][ACG]ACG", False)):
        for mo in MATCH:
            for nd in (SQ_FAIL, SQ_CONVERT):
                st = scan(B, oracle, pattern, 2, buf, mo | nd)
                assert bool(st.path & BITSLICE) == sliced, (pattern, st.path)


def test_tile_and_buffer_boundaries(B, oracle):
    """The fused kernel works on 32 KiB tiles with 512 bytes of overlap and masks everything at and behind the end of
    the buffer by hand: newlines on the last byte of a tile, lines that start on a tile's first byte, buffers that end
    exactly on a tile boundary / one byte before / one byte after it, with and without a final newline, and a line that
    ends on the very last byte of the staged overlap."""
    rng = random.Random(17)
    pattern, tau = "GATCGGAAGAGC", 2
    tile = 32768

    def lines_to(total, final_newline):
        """ragged lines whose bytes add up to exactly `total` (newlines included)"""
        out, left = [], total
        while left > 0:
            n = min(rng.randint(0, 120), left - 1)
            s = [rng.choice("ACGT") for _ in range(n)]
            if n > 14 and rng.random() < 0.5:
                at = rng.randrange(n - 12)
                s[at:at + 12] = pattern
            out.append("".join(s) + "\n")
            left -= n + 1
        buf = "".join(out)
        assert len(buf) == total
        return (buf if final_newline else buf[:-1] + "A").encode()

    sizes = [tile - 1, tile, tile + 1, 2 * tile - 1, 2 * tile, 2 * tile + 1, 3 * tile + 511, 3 * tile + 512, 3 * tile + 513, 5 * tile]
    for total in sizes:
        for final_newline in (True, False):
            buf = lines_to(total, final_newline)
            for mo in (SQ_BEST, SQ_ALL):
                st = scan(B, oracle, pattern, tau, buf, mo)
                assert st.path & FUSED, (total, st.path)
    # a newline exactly on the last byte of the first tile, the next line starting on the first byte of the second; a line
    # of exactly 512 bytes (terminator included) that starts on the last byte of a tile: it ends on the last staged byte
    body = lines_to(tile, True)
    assert body[tile - 1:tile] == b"\n"
    exact = ("A" * 200 + pattern + "C" * (511 - 200 - 12)).encode() + b"\n"          # 512 bytes with the newline
    head = lines_to(tile - 1, True)
    for buf in (body + lines_to(1000, True), head + exact + lines_to(3000, True), head + exact):
        for mo in MATCH:
            st = scan(B, oracle, pattern, tau, buf, mo)
            assert st.path & FUSED, st.path


def test_fused_path_against_the_compiled_reference(B, reference):
    """The records of the CUDA path against the UNMODIFIED reference (oracle/_ref, compiled from /root/reference in the
    build container and carried to the GPU box), not only against the oracle port: 60 000 reads, three modes."""
    g = B.make_gen(seed=12, line_len=150, plant="TTGACAGCTAGCTCAGTCCT", plant_per_1024=200, max_edits=2, n_per_1024=3)
    buf = B.gen_host(g, 60000)
    for pattern, tau in (("TTGACAGCTAGCTCAGTCCT", 2), ("A[CG]TNNGATC", 1)):
        sq = B.Seeq(pattern, tau)
        for mo in MATCH:
            st = B.StatsT()
            recs = sq.batch(buf, mo, B.SQ_ANY, st)
            exp, nl, nm = reference.buffer_scan(buf, pattern, tau, mo, cap=max(1024, 4 * 60000))
            got = np.stack([recs["line"].astype(np.uint64) + 1, recs["start"], recs["end"], recs["dist"]], axis=1).astype(np.uint64)
            assert st.path & FUSED
            assert (st.nlines, st.nmatched) == (nl, nm)
            assert np.array_equal(got, exp), (pattern, mo)
        sq.close()
