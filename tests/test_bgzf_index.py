"""CPU: sqbBgzfIndex (host only).  The members of a BGZF buffer are found by walking the headers; buffers of 8 MiB or
more are walked by several threads from speculative starting points (sqb_bgzf.cu: index_members) -- the result must
be the serial walk's whatever the bytes look like: look-alike headers inside stored data, damage, truncation."""
import struct
import zlib

import numpy as np
import pytest

from seeq_b200 import bgzf


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


def walk(gz):
    """the members by the format's own rule (SAM specification 4.1), in Python"""
    out, off, o = [], 0, 0
    while off < len(gz):
        assert gz[off:off + 4] == b"\x1f\x8b\x08\x04" and gz[off + 12:off + 16] == b"BC\x02\x00"
        bsize = int.from_bytes(gz[off + 16:off + 18], "little") + 1
        isize = int.from_bytes(gz[off + bsize - 4:off + bsize], "little")
        if isize:
            out.append((off + 18, bsize - 26, isize, o))
        o += isize
        off += bsize
    return out, o


def index(B, gz, threads, monkeypatch):
    monkeypatch.setenv("SEEQ_B200_BGZF_INDEX_THREADS", str(threads))
    members, cnt, tb = B.bgzf_index(gz)
    return [(m.in_off, m.in_len, m.isize, m.out_off) for m in members[:cnt]], tb


def big_buffer(fakes):
    rng = np.random.default_rng(17)
    noise = rng.integers(0, 256, size=20 << 20, dtype=np.uint8)
    if fakes:
        # two empty members back to back parse as a member followed by a member: a perfect look-alike, planted
        # wherever a thread may start looking (the text is stored, not deflated: the bytes survive as they are)
        fake = np.frombuffer(bgzf.EOF_MEMBER * 2, dtype=np.uint8)
        for at in range(1 << 18, noise.size - 100, 1 << 18):
            noise[at:at + fake.size] = fake
    dna = rng.choice(np.frombuffer(b"ACGT\n", dtype=np.uint8), size=24 << 20).tobytes()
    return bgzf.compress(noise.tobytes(), level=0) + bgzf.compress(dna, level=1, processes=8)


@pytest.mark.parametrize("fakes", [False, True], ids=["plain", "look-alike headers"])
def test_threads_find_what_the_serial_walk_finds(B, fakes, monkeypatch):
    gz = big_buffer(fakes)
    assert len(gz) > (24 << 20)
    want, text = walk(gz)
    for threads in (1, 2, 3, 8):
        got, tb = index(B, gz, threads, monkeypatch)
        assert tb == text and got == want, threads
    # the end-of-file member in the middle and at the end, members of odd sizes
    parts = [bgzf.member(bytes([65 + i % 4]) * (1 + 977 * i % 60000), level=1 + i % 3) for i in range(900)]
    parts[300:300] = [bgzf.EOF_MEMBER]
    gz2 = b"".join(parts) + bgzf.EOF_MEMBER
    assert len(gz2) < (8 << 20)
    gz2 = gz2 + gz
    want, text = walk(gz2)
    got, tb = index(B, gz2, 8, monkeypatch)
    assert tb == text and got == want


def test_damage_is_refused_by_every_walk(B, monkeypatch):
    gz = bytearray(big_buffer(False))
    want, _ = walk(bytes(gz))
    mid = want[len(want) // 2][0] - 18
    for threads in (1, 8):
        for kind in ("magic", "bsize", "truncated", "isize"):
            bad = bytearray(gz)
            if kind == "magic":
                bad[mid] = 0
            elif kind == "bsize":
                struct.pack_into("<H", bad, mid + 16, 17)
            elif kind == "isize":
                end = mid + int.from_bytes(bad[mid + 16:mid + 18], "little") + 1
                struct.pack_into("<I", bad, end - 4, 70000)
            else:
                bad = bad[:-5]
            monkeypatch.setenv("SEEQ_B200_BGZF_INDEX_THREADS", str(threads))
            with pytest.raises(ValueError, match="BGZF member"):
                B.bgzf_index(bytes(bad))
    monkeypatch.setenv("SEEQ_B200_BGZF_INDEX_THREADS", "8")
    with pytest.raises(ValueError, match="not a BGZF member"):
        B.bgzf_index(zlib.compress(bytes(10 << 20), 1, 31) * 40)                    # plain gzip, 10 MB of it
    assert B.bgzf_index(b"")[1:] == (0, 0) and B.bgzf_index(bgzf.EOF_MEMBER)[1:] == (0, 0)


def test_inflating_fails_loudly_without_a_device(B):
    """no CPU fallback: the inflater is CUDA only (zlib is the CHECKER of the tests, never the product)"""
    if B.lib().sqbDeviceCount() > 0:
        pytest.skip("a CUDA device is present")
    gz = bgzf.compress(b"ACGT\n" * 1000)
    with pytest.raises(RuntimeError, match="failed"):
        B.bgzf_inflate_device(gz)
    import ctypes as C
    members, cnt, tb = B.bgzf_index(gz)
    text = (C.c_ubyte * tb)()
    rc = B.lib().sqbBgzfInflateDevice(0, gz, members, cnt, text, None, None)
    assert rc == -1 and "CUDA" in B.last_error()
    assert B.lib().sqbScanHostBgzf(None, gz, len(gz), 0, None) == -1
