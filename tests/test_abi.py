"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, keeps the
struct layouts callers depend on, and reproduces the reference's constructor / parser /
error behaviour.  No matching is computed here (no GPU in this container): the matching
entry points must FAIL LOUDLY without a CUDA device, never fall back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


def declared_functions():
    names = set()
    for h in ("libseeq.h", "seeq.h", "seeq_b200.h"):
        src = open(os.path.join(INC, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(([^;{}()]*)\)\s*;", src):
            names.add(m.group(1))
    return names


def test_every_declared_symbol_is_exported(B):
    names = declared_functions()
    assert {"seeqNew", "seeqStringMatch", "seeqFileMatch", "seeqMatchIter", "seeq", "sqbScanDevice",
            "sqbScanHost", "seeqBatchMatch"} <= names
    out = subprocess.run(["nm", "-D", "--defined-only", B._build.LIB], stdout=subprocess.PIPE, text=True,
                         check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    missing = sorted(n for n in names if n not in exported)
    assert not missing, missing
    assert "seeqerr" in exported
    # and the binding types every one of them
    assert names <= set(B.SYMBOLS), sorted(names - set(B.SYMBOLS))


def test_no_oracle_or_torch_in_the_product(B):
    """The shipped library links CUDA runtime + libc only; nothing under oracle/."""
    out = subprocess.run(["ldd", B._build.LIB], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out and "torch" not in out and "libseeq_ref" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "seeq_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f), errors="replace").read()
                assert "pyoracle" not in src and "liboracle" not in src and "seeq_oracle" not in src, f


def test_struct_layouts(B):
    # SURVEY 8a [probed on the reference]: seeq_t is 80 bytes with these offsets
    T = B.SeeqT
    assert C.sizeof(T) == 80
    assert [getattr(T, n).offset for n in ("hits", "stacksize", "match", "bufsz", "string", "tau", "wlen",
                                           "keys", "rkeys", "dfa", "rdfa")] == [0, 8, 16, 24, 32, 40, 44, 48, 56, 64, 72]
    assert C.sizeof(B.MatchT) == 24
    F = B.SeeqFileT
    assert [getattr(F, n).offset for n in ("flags", "line", "info", "fdi")] == [0, 8, 16, 24]
    assert C.sizeof(B.StatsT) == 4 * 8 + 8 + 8 * 8 + 16 and C.sizeof(B.GenT) == 8 + 7 * 4 + 256 + 4


def test_seeqnew_fields_and_errors(B):
    """testset.c:763-827 (test_seeqNew)."""
    L = B.lib()
    sq = B.Seeq("CACAGAT", 3)
    s = sq.sq.contents
    assert (s.hits, s.stacksize, s.tau, s.wlen) == (0, 16, 3, 7)
    assert sq.keys == bytes([2, 1, 2, 1, 4, 1, 8])
    assert bytes(s.rkeys[i][0] for i in range(7)) == bytes([8, 1, 4, 1, 2, 1, 2])
    assert s.dfa and s.rdfa and s.match and not s.string
    sq.close()
    for pattern, tau, err in [("CACAGAT", -1, 1), ("CACAGAT", 7, 9), ("CACAGAT", 8, 9), ("CAC[[AT]", 1, 2),
                              ("CAC]AT", 1, 3), ("CACXAT", 1, 4), ("CAC[AT", 1, 5)]:
        p = L.seeqNew(pattern.encode(), tau, 0)
        assert not p and B.seeqerr() == err, (pattern, tau)
    assert L.seeqPrintError() == b"Incorrect pattern (missing closing bracket)"


@pytest.mark.parametrize("pattern,keys", [
    ("ACGTUNacgtun", [1, 2, 4, 8, 8, 31, 1, 2, 4, 8, 8, 31]),
    ("A[CG]TNNGATC", [1, 6, 8, 31, 31, 4, 1, 8, 2]),
    ("Nn[]Nn[]NnN[]n", [31] * 8),               # testset.c:713-761: "[]" adds no position
    ("[ACGT][acgu]", [15, 15]),
])
def test_parser_matches_oracle(B, oracle, pattern, keys):
    sq = B.Seeq(pattern, 0)
    assert list(sq.keys) == keys
    assert oracle.parse(pattern)[0] == bytes(keys)
    sq.close()


def test_open_close_and_errors(B, tmp_path):
    L = B.lib()
    assert not L.seeqOpen(b"/nonexistent/invented.txt") and B.seeqerr() == 2     # testset.c:1221-1222
    p = tmp_path / "x.fa"
    p.write_bytes(b">hdr\nACGT\n")
    f = L.seeqOpen(str(p).encode())
    assert f and f.contents.flags == 1 and f.contents.line == 0
    sq = B.Seeq("ACG", 0)
    f.contents.fdi = None
    assert L.seeqFileMatch(f, sq.sq, 0, 0) == -1 and B.seeqerr() == 10            # testset.c:933-939
    assert L.seeqClose(f) == 0
    sq.close()


def test_iterator_and_legacy_helpers(B):
    L = B.lib()
    sq = B.Seeq("ACGT", 1)
    for k in range(40):                                    # grows past INITIAL_MATCH_STACK_SIZE
        assert L.seeqAddMatch(sq.sq, B.MatchT(k, k + 4, 0)) == 0
    assert sq.sq.contents.hits == 40 and sq.sq.contents.stacksize >= 40
    m = L.seeqMatchIter(sq.sq)
    assert (m.contents.start, m.contents.end) == (39, 43) and sq.sq.contents.hits == 39
    sq.close()
    st = L.stackNew(2)
    assert st


def test_matching_fails_loudly_without_a_device(B):
    """No CPU fallback: on a box without CUDA the matcher returns -1 (ENODEV)."""
    if B.lib().sqbDeviceCount() > 0:
        pytest.skip("a CUDA device is present")
    sq = B.Seeq("ACGT", 1)
    assert B.lib().seeqStringMatch(b"ACGTACGT", sq.sq, 0) == -1
    assert "no CUDA device" in B.last_error()
    with pytest.raises(RuntimeError):
        sq.batch(b"ACGT\n", 0, 0)
    with pytest.raises(RuntimeError):
        B.Engine(b"\x01\x02", 0)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        B.Multi([b"\x01\x02\x04", b"\x08\x01"], [1, 0])        # a pattern set needs its engines
    sq.close()


def test_shard_range_agrees_with_numpy_statement(B):
    import numpy as np
    from seeq_b200 import shard
    rng = np.random.default_rng(3)
    for n in (0, 1, 5, 257, 10_000):
        buf = rng.choice(np.frombuffer(b"ACGT\n", dtype=np.uint8), n, p=[.24, .24, .24, .24, .04])
        for world in (1, 2, 3, 8):
            want = shard.shard_ranges(buf, world)
            got = [B.shard_range(buf, r, world) for r in range(world)] if n else [(0, 0)] * world
            assert got == want, (n, world)
            assert want[0][0] == 0 and want[-1][1] == n
            for (b0, e0), (b1, e1) in zip(want, want[1:]):
                assert e0 == b1 and (b1 == 0 or b1 == n or buf[b1 - 1] == 0x0A)


def test_relinked_module_has_the_batched_method():
    """seeq_b200/_relink/seeq*.so (the reference's seeqmodule.c compiled in place + matchBatch): imports,
    keeps the reference's surface (seeqmodule.c:986-1014, :1097-1103) and fails loudly without a device."""
    import glob
    import importlib.util
    from seeq_b200 import build
    build.build_library()
    out = build.build_relinks()
    if "module" not in out:
        pytest.skip("reference tree absent: the module cannot be re-linked here")
    mods = glob.glob(os.path.join(os.path.dirname(out["module"]), "seeq*.so"))
    spec = importlib.util.spec_from_file_location("seeq", mods[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    m = mod.compile("CGCTAATTAATGGAAT", 3)
    for name in ("match", "matchBest", "matchAll", "matchIter", "matchPrefix", "matchSuffix", "matchBatch"):
        assert callable(getattr(m, name)), name
    with pytest.raises(mod.exception):
        m.matchBatch(b"ACGT\n", mode="fastest")
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(mod.clibexception):
            m.matchBatch(b"ACGT\nGGGGCGCTAATAATGGAATGGGG\n")
