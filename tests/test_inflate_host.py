"""CPU: the DEFLATE decoder of the device inflater (seeq_b200/csrc/sqb_inflate.h: bit reader, block headers, canonical
decoder, the three-literal table, the symbol loop) compiled for the host and run the way one warp of k0_inflate_bgzf
runs it (tests/host_inflate.cpp), against zlib: every block type, every compression level, DNA / FASTQ / binary /
repetitive text, members of every size up to 64 KiB, and damaged streams (which must be refused, not crash)."""
import ctypes as C
import os
import random
import subprocess
import zlib

import numpy as np
import pytest

from seeq_b200 import bgzf

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "seeq_b200", "csrc")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("inflate") / "host_inflate.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-w", "-shared", "-fPIC", "-pthread", "-I" + CSRC,
                    os.path.join(HERE, "host_inflate.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.host_bgzf_inflate_pair.restype = C.c_longlong
    L.host_bgzf_inflate_pair.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    L.host_set_queue.argtypes = [C.c_uint32]
    L.host_bgzf_inflate.restype = C.c_longlong
    L.host_bgzf_inflate.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    L.host_bgzf_index.restype = C.c_longlong
    L.host_bgzf_index.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.host_lit_entry_of.restype = C.c_uint32
    L.host_lit_entry_of.argtypes = [C.c_uint32, C.c_uint32]
    L.host_dist_entry_of.restype = C.c_uint32
    L.host_dist_entry_of.argtypes = [C.c_uint32, C.c_uint32]
    return L


@pytest.fixture(params=[32, 16], ids=["queue32", "queue16"])
def H(host_lib, request):
    """matches per queue: 32 = one member per warp (k0_inflate_bgzf), 16 = two members per warp (k0_inflate_bgzf_pair)"""
    L = host_lib
    L.host_set_queue(request.param)
    return L


def inflate(L, gz, cap=None, stats=None):
    n = C.c_uint64(0)
    assert L.host_bgzf_index(gz, len(gz), C.byref(n)) >= 0
    cap = n.value if cap is None else cap
    guard = 64
    buf = (C.c_ubyte * (cap + 2 * guard))()
    C.memset(buf, 0xA5, cap + 2 * guard)
    rc = L.host_bgzf_inflate(gz, len(gz), C.addressof(buf) + guard, cap, stats)
    raw = bytes(buf)
    assert raw[:guard] == b"\xa5" * guard and raw[guard + cap:] == b"\xa5" * guard, "wrote outside the text buffer"
    return rc, raw[guard:guard + cap]


def dna(rng, n, line=150):
    s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    s[line::line + 1] = 10
    return s.tobytes()


def fastq(rng, nrec, line=100):
    out = []
    for i in range(nrec):
        seq = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=line, p=[.245, .245, .245, .245, .02]).tobytes()
        qual = (rng.integers(0, 41, size=line) + 33).astype(np.uint8).tobytes()
        out.append(b"@SRR0000001.%d HWI-ST1234:100:C0ABCACXX:1:1101:%d:%d length=%d\n" % (i, 1000 + i % 977, 2000 + i % 1511, line)
                   + seq + b"\n+\n" + qual + b"\n")
    return b"".join(out)


TEXTS = {
    "dna": lambda rng: dna(rng, 300000),
    "fastq": lambda rng: fastq(rng, 1500),
    "binary": lambda rng: rng.integers(0, 256, size=150000, dtype=np.uint8).tobytes(),
    "runs": lambda rng: b"".join(bytes([int(rng.integers(65, 70))]) * int(rng.integers(1, 600)) for _ in range(800)),
    "repeats": lambda rng: (dna(rng, 700, line=10 ** 9) * 400)[:250000],
    "skewed": lambda rng: rng.choice(np.arange(256, dtype=np.uint8), size=200000,
                                     p=np.array([2.0 ** -(1 + i % 40) for i in range(256)]) /
                                     sum(2.0 ** -(1 + i % 40) for i in range(256))).tobytes(),
}


def test_length_and_distance_symbols_follow_rfc1951():
    """base values and extra bits of RFC 1951 §3.2.5, computed arithmetically in the header, against the tables"""
    lbase = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
    lext = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
    dbase = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
             6145, 8193, 12289, 16385, 24577]
    dext = [0, 0, 0, 0] + [i // 2 for i in range(2, 28)]
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        so = os.path.join(d, "h.so")
        subprocess.run(["g++", "-std=c++17", "-O2", "-w", "-shared", "-fPIC", "-I" + CSRC,
                        os.path.join(HERE, "host_inflate.cpp"), "-o", so], check=True)
        L = C.CDLL(so)
        L.host_lit_entry_of.restype = C.c_uint32
        L.host_dist_entry_of.restype = C.c_uint32
        for s in range(257, 286):
            e = L.host_lit_entry_of(s, 9)
            assert e & 15 == 9 and (e >> 4) & 3 == 0 and (e >> 6) & 3 == 0
            assert (e >> 8) & 0x1ff == lbase[s - 257] and (e >> 17) & 7 == lext[s - 257], s
        for s in range(256):
            assert L.host_lit_entry_of(s, 7) == 7 | 1 << 4 | s << 8
        assert (L.host_lit_entry_of(256, 7) >> 6) & 3 == 1
        assert (L.host_lit_entry_of(286, 8) >> 6) & 3 == 3 and (L.host_lit_entry_of(287, 8) >> 6) & 3 == 3
        for s in range(30):
            e = L.host_dist_entry_of(s, 5)
            assert e & 15 == 5 and (e >> 4) & 15 == dext[s] and e >> 16 == dbase[s] and not e & 0x300, s
        assert L.host_dist_entry_of(30, 5) & 0x200 and L.host_dist_entry_of(31, 5) & 0x200


@pytest.mark.parametrize("kind", sorted(TEXTS))
@pytest.mark.parametrize("level", [1, 6, 9])
def test_text_kinds_and_levels(H, kind, level):
    rng = np.random.default_rng(1000 * sorted(TEXTS).index(kind) + level)
    text = TEXTS[kind](rng)
    gz = bgzf.compress(text, level=level)
    assert bgzf.decompress_cpu(gz) == text
    stats = (C.c_uint32 * 16)()
    rc, out = inflate(H, gz, stats=stats)
    assert rc == len(text), rc
    assert out == text
    if kind in ("dna", "fastq"):
        # [2] dynamic blocks, [6] table entries holding two literals, [9] matches copied by one lane each, [10] rounds
        assert stats[2] > 0 and (stats[6] > 0 or level == 1 or kind == "fastq") and stats[9] > 0
        assert stats[10] < 8 * stats[11], "the queue resolves in a few rounds, not match by match"


def test_three_literals_per_look_up(H):
    """DNA without matches (Z_HUFFMAN_ONLY): code words of 2-3 bits, so the 9 index bits hold three bases"""
    rng = np.random.default_rng(9)
    text = dna(rng, 200000)
    gz = bgzf.compress(text, level=6, strategy=zlib.Z_HUFFMAN_ONLY)
    stats = (C.c_uint32 * 16)()
    rc, out = inflate(H, gz, stats=stats)
    assert rc == len(text) and out == text
    assert stats[7] > 300 * stats[2] and stats[9] == 0 and stats[8] == 0


def test_block_types_and_strategies(H):
    rng = np.random.default_rng(7)
    text = fastq(rng, 600)
    for strategy, want in ((zlib.Z_FIXED, 1), (zlib.Z_HUFFMAN_ONLY, 2), (zlib.Z_RLE, 2), (zlib.Z_FILTERED, 2)):
        gz = bgzf.compress(text, level=6, strategy=strategy)
        stats = (C.c_uint32 * 16)()
        rc, out = inflate(H, gz, stats=stats)
        assert rc == len(text) and out == text, strategy
        assert stats[want] > 0
    gz = bgzf.compress(text, level=0)                           # stored blocks
    stats = (C.c_uint32 * 16)()
    rc, out = inflate(H, gz, stats=stats)
    assert rc == len(text) and out == text and stats[0] > 0
    noise = rng.integers(0, 256, size=200000, dtype=np.uint8).tobytes()   # does not deflate: zlib stores
    rc, out = inflate(H, bgzf.compress(noise, level=6))
    assert rc == len(noise) and out == noise


def test_member_sizes_and_flush_points(H):
    """members of 1 byte .. 64 KiB, several deflate blocks per member (full flushes), empty members in between"""
    rng = np.random.default_rng(11)
    text = fastq(rng, 900)
    parts = []
    sizes = [1, 2, 3, 7, 31, 32, 33, 255, 256, 257, 258, 259, 4095, 65535, 65536, 65280, 1000]
    off = 0
    want = b""
    for i, sz in enumerate(sizes):
        piece = (text * 2)[off:off + sz]
        off += 13
        parts.append(bgzf.member(piece, level=1 + i % 9))
        if i % 5 == 0:
            parts.append(bgzf.EOF_MEMBER)
        want += piece
    # one member made of several deflate blocks, of all three types
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    piece = text[:40000]
    cdata = co.compress(piece[:10000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(piece[10000:10001]) + \
        co.flush(zlib.Z_SYNC_FLUSH) + co.compress(piece[10001:]) + co.flush()
    import struct
    head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 0x42, 0x43, 2, len(cdata) + 25)
    parts.append(head + cdata + struct.pack("<II", zlib.crc32(piece), len(piece)))
    want += piece
    gz = b"".join(parts) + bgzf.EOF_MEMBER
    assert bgzf.decompress_cpu(gz) == want
    stats = (C.c_uint32 * 16)()
    rc, out = inflate(H, gz, stats=stats)
    assert rc == len(want) and out == want
    assert stats[0] > 0, "the flush points are empty stored blocks"


def test_long_code_words_take_the_canonical_path(H):
    """a skewed alphabet gives code words longer than the 9 index bits of the table"""
    rng = np.random.default_rng(5)
    p = np.array([0.5 ** min(i + 1, 22) for i in range(200)])
    text = rng.choice(np.arange(200, dtype=np.uint8), size=400000, p=p / p.sum()).tobytes()
    gz = bgzf.compress(text, level=6, strategy=zlib.Z_HUFFMAN_ONLY)
    rc, out = inflate(H, gz)
    assert rc == len(text) and out == text


def test_damaged_streams_are_refused(H):
    """bit flips, truncation and wrong ISIZE: an error or (for flips that keep the stream valid) different text --
    never a write outside the buffer (the guard bytes of inflate() see to that), never a hang"""
    rng = random.Random(3)
    nrng = np.random.default_rng(3)
    text = fastq(nrng, 300)
    gz = bytearray(bgzf.compress(text, level=6))
    refused = accepted = 0
    for trial in range(400):
        bad = bytearray(gz)
        for _ in range(rng.randint(1, 3)):
            pos = rng.randrange(18, len(bad) - 28 - 8)
            bad[pos] ^= 1 << rng.randrange(8)
        n = C.c_uint64(0)
        if H.host_bgzf_index(bytes(bad), len(bad), C.byref(n)) < 0:
            refused += 1
            continue
        rc, out = inflate(H, bytes(bad), cap=n.value)
        if rc < 0:
            refused += 1
            continue
        # accepted: the flip left a valid stream of the announced size (a changed literal, say; CRC-32 is not checked).
        # zlib, told to ignore the CRC (raw inflate of every member), must read the same text out of it.
        assert rc == n.value
        accepted += 1
        off, want = 0, b""
        raw = bytes(bad)
        while off < len(raw):
            bsize = int.from_bytes(raw[off + 16:off + 18], "little") + 1
            want += zlib.decompressobj(-15).decompress(raw[off + 18:off + bsize - 8])
            off += bsize
        assert out == want
    assert refused > 150 and accepted > 50
    # ISIZE too small / too large
    import struct
    m = bytearray(bgzf.member(text[:5000]))
    for isize in (4999, 5001, 0, 65536):
        m[-4:] = struct.pack("<I", isize)
        if isize == 0:
            continue                                           # an "empty" member is skipped, not decoded
        rc, _ = inflate(H, bytes(m), cap=isize)
        assert rc < 0, isize
    # truncated deflate stream inside a member that claims the full size
    whole = bgzf.member(text[:20000])
    cut = bytearray(whole[:18] + whole[18:-8][:-40] + whole[-8:])
    struct.pack_into("<H", cut, 16, len(cut) - 1)
    rc, _ = inflate(H, bytes(cut), cap=20000)
    assert rc < 0
    # not BGZF: plain gzip
    assert H.host_bgzf_index(zlib.compress(text, 6, 31), 100, C.byref(C.c_uint64())) == -1


def inflate_pair(L, gz, cap):
    guard = 64
    buf = (C.c_ubyte * (cap + 2 * guard))()
    C.memset(buf, 0xA5, cap + 2 * guard)
    status = (C.c_uint32 * 4096)()
    rc = L.host_bgzf_inflate_pair(gz, len(gz), C.addressof(buf) + guard, cap, status)
    raw = bytes(buf)
    assert raw[:guard] == b"\xa5" * guard and raw[guard + cap:] == b"\xa5" * guard, "wrote outside the text buffer"
    return rc, raw[guard:guard + cap], status


def test_two_members_per_warp(host_lib):
    """sqb_bgzf_warp.h -- the body of k0_inflate_bgzf_pair, the same source -- run by 32 host threads that meet at a
    barrier wherever the lanes of the warp exchange something: block headers by the two decoding lanes, tables by the
    half-warps, queues of 16 matches, stored blocks, members that end at different times, an odd number of members,
    a half with a damaged member next to a sound one."""
    rng = np.random.default_rng(23)
    pieces = [dna(rng, 9000), fastq(rng, 25), rng.integers(0, 256, size=5000, dtype=np.uint8).tobytes(),
              b"".join(bytes([int(rng.integers(65, 70))]) * int(rng.integers(1, 300)) for _ in range(60)),
              dna(rng, 3000, line=60) * 3, b"A", dna(rng, 20000)]
    parts, want = [], b""
    for i, piece in enumerate(pieces):
        if i == 3:
            parts.append(bgzf.EOF_MEMBER)
        strategy = zlib.Z_FIXED if i == 4 else zlib.Z_DEFAULT_STRATEGY
        parts.append(bgzf.member(piece, level=(0 if i == 2 else 6), strategy=strategy))
        want += piece
    # several deflate blocks of all types in one member, next to a member of one block
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    piece = fastq(rng, 40)
    cdata = co.compress(piece[:3000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(piece[3000:3001]) + \
        co.flush(zlib.Z_SYNC_FLUSH) + co.compress(piece[3001:]) + co.flush()
    import struct
    head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 0x42, 0x43, 2, len(cdata) + 25)
    parts.append(head + cdata + struct.pack("<II", zlib.crc32(piece), len(piece)))
    want += piece
    gz = b"".join(parts) + bgzf.EOF_MEMBER
    assert bgzf.decompress_cpu(gz) == want and len(pieces) % 2 == 1
    rc, out, status = inflate_pair(host_lib, gz, len(want))
    assert rc == len(want), rc
    assert out == want
    assert list(status[:len(pieces) + 1]) == [0] * (len(pieces) + 1)
    # an odd number of members: the upper half of the last warp has none
    odd = b"".join(parts[:-1]) + bgzf.EOF_MEMBER
    rc, out, status = inflate_pair(host_lib, odd, len(want) - len(piece))
    assert rc == len(want) - len(piece) and out == want[:-len(piece)]
    assert list(status[:len(pieces)]) == [0] * len(pieces)
    # member 1 damaged (ISIZE one short), its neighbour in the warp and every other member are inflated all the same
    bad = bytearray(gz)
    off = len(parts[0])
    end = off + len(parts[1])
    bad[end - 4:end] = struct.pack("<I", len(pieces[1]) - 1)
    rc, out, status = inflate_pair(host_lib, bytes(bad), len(want))
    assert rc == -(16 * (1 + 1) + 2), rc                          # member 1: more text than ISIZE announces
    assert status[0] == 0 and status[1] == 2 and list(status[2:len(pieces) + 1]) == [0] * (len(pieces) - 1)
    assert out[:len(pieces[0])] == pieces[0]


def test_two_members_per_warp_under_thread_sanitizer(tmp_path):
    """The same host run of sqb_bgzf_warp.h under ThreadSanitizer: between two barriers of the warp no lane reads or
    writes a byte -- of the tables, the queues or the TEXT -- that another lane writes.  (compute-sanitizer's racecheck
    on the device sees shared memory only; the text lives in global memory.)"""
    import glob
    import sys
    tsan = glob.glob("/usr/lib/gcc/x86_64-linux-gnu/*/libtsan.so") + glob.glob("/usr/lib/x86_64-linux-gnu/libtsan.so*")
    if not tsan:
        pytest.skip("no libtsan in this image")
    so = str(tmp_path / "host_inflate_tsan.so")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-w", "-shared", "-fPIC", "-pthread", "-I" + CSRC,
                        os.path.join(HERE, "host_inflate.cpp"), "-o", so], capture_output=True, text=True)
    if r.returncode:
        pytest.skip("g++ -fsanitize=thread does not build here: " + r.stderr[-200:])
    code = """
import ctypes as C, sys
sys.path.insert(0, %r)
import numpy as np
from seeq_b200 import bgzf
from tests.test_inflate_host import dna, fastq
L = C.CDLL(%r)
L.host_bgzf_inflate_pair.restype = C.c_longlong
L.host_bgzf_inflate_pair.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
rng = np.random.default_rng(2)
text = dna(rng, 6000) + fastq(rng, 20) + rng.integers(0, 256, size=3000, dtype=np.uint8).tobytes() + dna(rng, 5000)
gz = bgzf.compress(text, level=6, block=4000)
out = (C.c_ubyte * len(text))()
rc = L.host_bgzf_inflate_pair(gz, len(gz), out, len(text), None)
print("RESULT", rc == len(text) and bytes(out) == text)
""" % (os.path.dirname(HERE), so)
    env = dict(os.environ, LD_PRELOAD=tsan[0], TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    if "RESULT" not in r.stdout:
        pytest.skip("the interpreter does not run under LD_PRELOAD=libtsan here: " + r.stderr[-300:])
    assert "RESULT True" in r.stdout
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-3000:]


def test_fuzz_against_zlib(host_lib):
    """hypothesis: texts of every texture (alphabet size, repeat structure, length up to a member and a bit), every
    level and strategy, members cut anywhere -- the decoder reads what zlib reads, for both queue sizes"""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @st.composite
    def texts(draw):
        seed = draw(st.integers(0, 2 ** 32 - 1))
        rng = np.random.default_rng(seed)
        n = draw(st.sampled_from([0, 1, 2, 5, 300, 4000, 30000, 66000, 140000]))
        alphabet = draw(st.sampled_from([1, 2, 4, 5, 20, 64, 256]))
        base = rng.integers(0, alphabet, size=max(n, 1), dtype=np.uint8) + (0 if alphabet == 256 else 48)
        period = draw(st.sampled_from([0, 1, 3, 17, 260, 5000, 40000]))
        if period and n > period:                                    # repeats at a fixed distance, with a few edits
            reps = -(-n // period)
            base = np.tile(base[:period], reps)[:n].copy()
            for at in rng.integers(0, n, size=n // 97 + 1):
                base[at] ^= 1
        return base[:n].tobytes()

    @settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck))
    @given(text=texts(), level=st.integers(0, 9),
           strategy=st.sampled_from([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]),
           block=st.sampled_from([1, 7, 500, 4097, 0xff00, 65536]), queue=st.sampled_from([16, 32]))
    def run(text, level, strategy, block, queue):
        if block < 500 and len(text) > 4000:
            text = text[:4000]                                       # thousands of tiny members: keep it quick
        try:
            gz = bgzf.compress(text, level=level, block=block, strategy=strategy)
        except AssertionError:                                       # 64 KiB that do not deflate do not fit a member
            return
        host_lib.host_set_queue(queue)
        rc, out = inflate(host_lib, gz)
        assert rc == len(text) and out == text

    run()
    host_lib.host_set_queue(32)
