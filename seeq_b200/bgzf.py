"""BGZF (bgzip) buffers: a writer for tests and benchmarks (Python's zlib makes the deflate streams; the image has no
bgzip binary) and the ctypes face of the device inflater (seeq_b200.h: sqbBgzfIndex / sqbBgzfInflateDevice /
sqbScanHostBgzf).  Format: SAM/BAM specification §4.1 -- gzip members (RFC 1952) of at most 64 KiB of text, an extra
sub-field "BC" holding the member's total size minus one, an empty member at the end of the file."""
from __future__ import annotations

import struct
import zlib

BLOCK_TEXT = 0xff00            # bytes of text per member, as bgzip cuts them
EOF_MEMBER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def member(text: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    """one BGZF member holding `text` (at most 64 KiB)"""
    assert len(text) <= 65536
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    cdata = co.compress(text) + co.flush()
    if len(cdata) + 26 > 65536:                       # does not deflate: stored
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        cdata = co.compress(text) + co.flush()
    bsize = len(cdata) + 25
    assert bsize < 65536, "a member must fit 64 KiB: cut the text smaller"
    head = struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 0x42, 0x43, 2, bsize)
    return head + cdata + struct.pack("<II", zlib.crc32(text) & 0xffffffff, len(text) & 0xffffffff)


def _job(args):
    text, level, strategy = args
    return member(text, level, strategy)


def compress(text: bytes, level: int = 6, block: int = BLOCK_TEXT, strategy: int = zlib.Z_DEFAULT_STRATEGY,
             eof: bool = True, processes: int = 1) -> bytes:
    """`text` as a BGZF buffer.  processes > 1 deflates the members side by side."""
    view = memoryview(text)
    parts = [(bytes(view[i:i + block]), level, strategy) for i in range(0, len(text), block)]
    if processes > 1 and len(parts) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(processes) as pool:
            out = pool.map(_job, parts, chunksize=max(1, len(parts) // (processes * 8)))
    else:
        out = [_job(p) for p in parts]
    if eof:
        out.append(EOF_MEMBER)
    return b"".join(out)


def decompress_cpu(gz: bytes) -> bytes:
    """zlib's answer, member by member (the checker of the tests)"""
    out = []
    off = 0
    while off < len(gz):
        d = zlib.decompressobj(31)
        out.append(d.decompress(gz[off:]))
        off = len(gz) - len(d.unused_data)
        assert d.eof
    return b"".join(out)
