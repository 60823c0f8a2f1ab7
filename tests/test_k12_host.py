"""CPU: the arithmetic of the fused tokenise + pack kernel (seeq_b200/csrc/sqb_k12_arith.h: newline flags of four bytes,
the flag word of a 32-byte chunk and its rank masks, the 32-bit class table, the PRMT assembly of plane words, the lead
(NULL) columns of lines that do not start on a 4-byte boundary) compiled for the host -- exhaustively where the domain is
small -- and one group of 32 lines built the way the kernel builds it, against class codes computed byte by byte."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle.pyoracle import SQ_CONVERT, SQ_FAIL, SQ_IGNORE

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "seeq_b200", "csrc")


@pytest.fixture(scope="module")
def K(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("k12") / "host_k12.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I" + CSRC,
                    os.path.join(HERE, "host_bitslice.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    for name in ("k12_nl_flags", "k12_chunk_before", "k12_chunk_byte_of"):
        getattr(L, name).restype = C.c_uint32
        getattr(L, name).argtypes = [C.c_uint32]
    L.k12_chunk_flags.restype = C.c_uint32
    L.k12_chunk_flags.argtypes = [C.c_char_p]
    L.k12_table32.restype = None
    L.k12_table32.argtypes = [C.c_int, C.POINTER(C.c_uint32)]
    L.k12_host_group.restype = C.c_uint32
    L.k12_host_group.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.c_int,
                                 C.POINTER(C.c_uint32), C.c_uint32]
    return L


def test_newline_flags_are_exact(K):
    """bit 7 of a byte is set iff the byte is 0x0A -- for every pair of neighbouring byte values in every position (the
    classic (x - 0x01..) & ~x test flags a 0x0B behind a newline; this one must not), and for random words."""
    for lo in range(256):
        for hi in range(256):
            for shift in (0, 8, 16):
                w = (lo << shift) | (hi << (shift + 8)) | (0x41414141 & ~(0xffff << shift))
                got = K.k12_nl_flags(w & 0xffffffff)
                exp = 0
                for b in range(4):
                    if (w >> (8 * b)) & 0xff == 0x0A:
                        exp |= 0x80 << (8 * b)
                assert got == exp, (hex(w), hex(got), hex(exp))
    rng = random.Random(1)
    for _ in range(20000):
        w = rng.getrandbits(32)
        exp = sum(0x80 << (8 * b) for b in range(4) if (w >> (8 * b)) & 0xff == 0x0A)
        assert K.k12_nl_flags(w) == exp


def test_chunk_flag_word_and_rank_masks(K):
    """the flag of byte v of a 32-byte chunk sits at bit u with chunk_byte_of(u) == v; chunk_before(v) is the set of the
    flags of the bytes in front of v -- what the kernel ranks the line starts of a chunk with."""
    pos_of = {}
    for u in range(32):
        v = K.k12_chunk_byte_of(u)
        assert 0 <= v < 32 and v not in pos_of
        pos_of[v] = u
    for v in range(33):
        exp = sum(1 << pos_of[x] for x in range(min(v, 32)))
        assert K.k12_chunk_before(v) == exp, v
    rng = random.Random(2)
    for _ in range(5000):
        chunk = bytearray(rng.choice(b"ACGTN\x0b\x00*") for _ in range(32))
        nls = sorted(rng.sample(range(32), rng.randint(0, 6)))
        for p in nls:
            chunk[p] = 0x0A
        word = K.k12_chunk_flags(bytes(chunk))
        assert word == sum(1 << pos_of[p] for p in nls), (chunk, nls)
        for rank, p in enumerate(nls):                          # the kernel's rank of a start inside its chunk
            assert bin(word & K.k12_chunk_before(p)).count("1") == rank


@pytest.mark.parametrize("nondna", [SQ_FAIL, SQ_CONVERT, SQ_IGNORE])
def test_class_table_and_one_group_of_planes(K, oracle, nondna):
    """the 32-bit class table against the oracle's byte classes, and the planes of a group -- four columns per lane, table
    look-up and multiply-add per byte, PRMT assembly, NULL lead columns -- against the class code of every (line, column)."""
    tab = (C.c_uint32 * 256)()
    K.k12_table32(nondna, tab)
    code = []
    for b in range(256):
        e = tab[b]
        c = (e & 1) | ((e >> 8) & 1) << 1 | ((e >> 16) & 1) << 2
        assert bool(e >> 24) == (b == 0x0A)
        code.append(c)
        # classes: 0..3 ACGT(U), 4 N (and other bytes with -x 1), 5 STOP, 6 SKIP (other bytes with -x 2)
        ch = chr(b)
        if ch in "Aa": exp = 0
        elif ch in "Cc": exp = 1
        elif ch in "Gg": exp = 2
        elif ch in "TtUu": exp = 3
        elif ch in "Nn": exp = 4
        elif b in (0, 0x0A): exp = 5
        else: exp = {SQ_FAIL: 5, SQ_CONVERT: 4, SQ_IGNORE: 6}[nondna]
        assert c == exp, (b, c, exp)
    rng = random.Random(3 + nondna)
    for it in range(60):
        nlines = rng.randint(1, 32)
        lines = ["".join(rng.choice("ACGTNacgtuX-\x00") if rng.random() < 0.1 else rng.choice("ACGT") for _ in range(rng.randint(0, 90)))
                 for _ in range(nlines)]
        pre = "".join(rng.choice("ACGT\n") for _ in range(rng.randint(0, 7)))
        text = (pre + "\n".join(lines) + "\n").encode("latin-1")
        starts, lens, at = [], [], len(pre)
        for s in lines:
            starts.append(at)
            lens.append(len(s) + 1)                              # with the terminator
            at += len(s) + 1
        buf = text + bytes(rng.choice(b"ACGT\n") for _ in range(256))          # what follows in the text: any bytes
        a_s = (C.c_uint32 * 32)(*starts)
        a_l = (C.c_uint32 * 32)(*lens)
        planes = (C.c_uint32 * 4096)()
        ncols = K.k12_host_group(buf, a_s, a_l, nlines, nondna, planes, 4096)
        assert ncols == max(l + (s & 3) for s, l in zip(starts, lens))
        for r in range(nlines):
            lead = starts[r] & 3
            for c in range(lens[r] + lead):
                blk, col = divmod(c, 32)
                got = sum(((planes[blk * 96 + p * 32 + col] >> r) & 1) << p for p in range(3))
                exp = 7 if c < lead else code[buf[starts[r] - lead + c]]
                assert got == exp, (it, r, c, got, exp)
