"""BGZF (bgzip) read sets inflated on the device (the reference reads plain text only,
seeq.c:201-256).  k0_inflate_bgzf through the C-ABI (sqbBgzfIndex / sqbBgzfInflateDevice / sqbScanHostBgzf) against
zlib for the text and against the oracle for the records: every block type, compression level and strategy, members
of every size, several slices, damaged streams (refused with the member named), plain gzip (refused)."""
import random
import struct
import zlib

import numpy as np
import pytest

from oracle.pyoracle import SQ_ALL, SQ_BEST, SQ_FIRST

from seeq_b200 import bgzf

from .test_gpu_fastq import rows
from .test_inflate_host import TEXTS, dna, fastq

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


@pytest.fixture(params=["pair", "single"], autouse=True)
def kernel(request, monkeypatch):
    """k0_inflate_bgzf_pair (two members per warp, the default) and k0_inflate_bgzf (one member per warp)"""
    monkeypatch.setenv("SEEQ_B200_BGZF_KERNEL", request.param)
    return request.param


def inflated(B, gz):
    out, ms = B.bgzf_inflate_device(gz)
    return out.tobytes()


@pytest.mark.parametrize("kind", sorted(TEXTS))
def test_text_kinds_and_levels(B, kind):
    for level in (1, 6, 9):
        rng = np.random.default_rng(1000 * sorted(TEXTS).index(kind) + level)
        text = TEXTS[kind](rng)
        gz = bgzf.compress(text, level=level)
        members, cnt, tb = B.bgzf_index(gz)
        assert tb == len(text) and cnt == (len(text) + bgzf.BLOCK_TEXT - 1) // bgzf.BLOCK_TEXT
        assert inflated(B, gz) == text, (kind, level)


def test_block_types_strategies_and_member_sizes(B):
    rng = np.random.default_rng(7)
    text = fastq(rng, 600)
    for strategy in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
        assert inflated(B, bgzf.compress(text, level=6, strategy=strategy)) == text, strategy
    assert inflated(B, bgzf.compress(text, level=0)) == text                        # stored blocks
    noise = rng.integers(0, 256, size=200000, dtype=np.uint8).tobytes()             # does not deflate: zlib stores
    assert inflated(B, bgzf.compress(noise, level=6)) == noise
    p = np.array([0.5 ** min(i + 1, 22) for i in range(200)])                       # code words beyond the table
    skew = rng.choice(np.arange(200, dtype=np.uint8), size=400000, p=p / p.sum()).tobytes()
    assert inflated(B, bgzf.compress(skew, level=6, strategy=zlib.Z_HUFFMAN_ONLY)) == skew
    # members of 1 byte .. 64 KiB at odd offsets, empty members in between, several deflate blocks in one member
    parts, want, off = [], b"", 0
    for i, sz in enumerate([1, 2, 3, 7, 31, 32, 33, 255, 256, 257, 258, 259, 4095, 65535, 65536, 65280, 1000]):
        piece = (text * 2)[off:off + sz]
        off += 13
        parts.append(bgzf.member(piece, level=1 + i % 9))
        if i % 5 == 0:
            parts.append(bgzf.EOF_MEMBER)
        want += piece
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    piece = text[:40000]
    cdata = co.compress(piece[:10000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(piece[10000:10001]) + \
        co.flush(zlib.Z_SYNC_FLUSH) + co.compress(piece[10001:]) + co.flush()
    head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 0x42, 0x43, 2, len(cdata) + 25)
    parts.append(head + cdata + struct.pack("<II", zlib.crc32(piece), len(piece)))
    want += piece
    gz = b"".join(parts) + bgzf.EOF_MEMBER
    assert bgzf.decompress_cpu(gz) == want
    assert inflated(B, gz) == want
    # nothing but the end-of-file member; nothing at all
    assert inflated(B, bgzf.EOF_MEMBER) == b"" and inflated(B, b"") == b""


def test_damaged_streams_are_refused(B):
    """an error that names the member, or (flips that leave a valid stream; CRC-32 is not checked) zlib's text"""
    rng = random.Random(3)
    text = fastq(np.random.default_rng(3), 300)
    gz = bgzf.compress(text, level=6)
    refused = accepted = 0
    for trial in range(120):
        bad = bytearray(gz)
        for _ in range(rng.randint(1, 3)):
            pos = rng.randrange(18, len(bad) - 28 - 8)
            bad[pos] ^= 1 << rng.randrange(8)
        bad = bytes(bad)
        try:
            out = inflated(B, bad)
        except (ValueError, RuntimeError) as err:
            assert "BGZF member" in str(err) or "sqbBgzfIndex" in str(err), err
            refused += 1
            continue
        accepted += 1
        off, want = 0, b""
        while off < len(bad):
            bsize = int.from_bytes(bad[off + 16:off + 18], "little") + 1
            want += zlib.decompressobj(-15).decompress(bad[off + 18:off + bsize - 8])
            off += bsize
        assert out == want
    assert refused > 30 and accepted > 10
    m = bytearray(bgzf.member(text[:5000]))
    for isize in (4999, 5001, 65536):
        m[-4:] = struct.pack("<I", isize)
        with pytest.raises(RuntimeError, match="BGZF member 0"):
            inflated(B, bytes(m))
    with pytest.raises(ValueError, match="not a BGZF member"):
        B.bgzf_index(zlib.compress(text, 6, 31))                                    # plain gzip


@pytest.mark.parametrize("kind", ["lines", "fastq"])
def test_scan_of_a_bgzf_buffer_matches_the_oracle(B, oracle, kind, monkeypatch):
    """sqbScanHostBgzf = sqbScanHost of the inflated text: records against the oracle, in several slices"""
    monkeypatch.setenv("SEEQ_B200_BGZF_SLICE_MB", "1")
    monkeypatch.setenv("SEEQ_B200_DEVICE_CHUNK_MB", "4")
    nrng = np.random.default_rng(21)
    pattern, tau = "GATTACAGATTACA", 2
    sq = B.Seeq(pattern, tau)
    eng = B.Engine.borrowed(sq.engine())
    if kind == "lines":
        text = bytearray(dna(nrng, 9_000_000, line=150))
        hit = b"GATTACGATTACA"                                                      # one deletion
        for at in range(500, len(text) - 200, 1510):
            if 10 not in text[at:at + len(hit)]:
                text[at:at + len(hit)] = hit
        text = bytes(text) + b"\n"
        options = [SQ_FIRST, SQ_BEST, SQ_ALL]
    else:
        text = fastq(nrng, 30000, line=100)
        options = [SQ_BEST | B.SQB_FASTQ]
    gz = bgzf.compress(text, level=6, processes=8)
    assert len(gz) > (2 << 20), "several slices"
    for opt in options:
        exp, nl, _ = oracle.buffer_scan(text, sq.keys, tau, opt & ~B.SQB_FASTQ)
        exp = np.asarray(exp, dtype=np.uint64).reshape(-1, 4)
        if opt & B.SQB_FASTQ:
            exp = exp[(exp[:, 0] - 1) % 4 == 1]
        st = eng.scan_host_bgzf(gz, opt)
        assert st.nbytes == len(text) and st.nlines == nl
        assert np.array_equal(rows(eng.host_records()), exp), (kind, opt)
        assert st.nmatched == len(np.unique(exp[:, 0])) and st.nmatched > 20
    assert eng.bgzf_text().tobytes() == text
    plain = eng.scan_host(text, options[-1])
    assert (plain.nlines, plain.nmatched, plain.nrecs) == (st.nlines, st.nmatched, st.nrecs)
    # a damaged member in the middle: the scan is refused, the member is named
    bad = bytearray(gz)
    members, cnt, _ = B.bgzf_index(gz)
    mid = members[cnt // 2]
    bad[mid.in_off + mid.in_len // 2:mid.in_off + mid.in_len // 2 + 64] = bytes(64)
    with pytest.raises(RuntimeError, match="BGZF member"):
        eng.scan_host_bgzf(bytes(bad), options[0])
    sq.close()
