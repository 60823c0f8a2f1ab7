"""The reference's CPython module, re-linked against libseeq_b200.so, on the GPU.

seeq_b200/_relink/seeq*.so = /root/reference/src/seeqmodule.c compiled in place (through the wrapper
translation unit seeq_b200/csrc/seeqmodule_b200.c, which adds SeeqObject.matchBatch) and linked against
our library; it travels to the GPU box prebuilt.  Reproduces /root/reference/test/python_lib_test.py:13-35
through it and checks the batched method against the oracle.
"""
import glob
import importlib.util
import os
import random

import numpy as np
import pytest

from oracle.pyoracle import SQ_ALL, SQ_BEST, SQ_CONVERT, SQ_FIRST, SQ_IGNORE

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
RELINK = os.path.join(os.path.dirname(HERE), "seeq_b200", "_relink")


@pytest.fixture(scope="module")
def seeq():
    from seeq_b200 import build
    build.build_library()
    build.build_relinks()                      # a no-op where /root/reference is absent
    mods = glob.glob(os.path.join(RELINK, "seeq*.so"))
    if not mods:
        pytest.skip("re-linked CPython module not built (reference tree absent at build time)")
    spec = importlib.util.spec_from_file_location("seeq", mods[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_python_lib_test_vectors(seeq):
    # /root/reference/test/python_lib_test.py:13-35
    matcher = seeq.compile("CGCTAATTAATGGAAT", 3)
    nomatch = "ATGCTGATGCTGGGGG"
    match = "GGGGCGCTAATAATGGAATGGGG"
    assert matcher.matchPrefix(nomatch, True) is None
    assert matcher.matchPrefix(nomatch, False) is None
    assert matcher.matchPrefix(match, True) == "GGGGCGCTAATAATGGAAT"
    assert matcher.matchPrefix(match, False) == "GGGG"
    assert matcher.matchSuffix(nomatch, True) is None
    assert matcher.matchSuffix(nomatch, False) is None
    assert matcher.matchSuffix(match, True) == "CGCTAATAATGGAATGGGG"
    assert matcher.matchSuffix(match, False) == "GGGG"
    assert seeq.__version__ == "1.2"


def test_match_objects(seeq, oracle):
    # match / matchBest / matchAll read sq->match through seeqMatchIter (seeqmodule.c:838-900)
    matcher = seeq.compile("GATCGGAAGAGC", 2)
    keys, _ = oracle.parse("GATCGGAAGAGC")
    rng = random.Random(7)
    for _ in range(20):
        s = "".join(rng.choice("ACGT") for _ in range(rng.randint(20, 90)))
        at = rng.randrange(len(s))
        s = s[:at] + "GATCGGAAGAGC"[:rng.randint(9, 12)] + s[at:]
        for meth, opt in ((matcher.match, SQ_FIRST), (matcher.matchBest, SQ_BEST), (matcher.matchAll, SQ_ALL)):
            exp = oracle.string_match(s.encode(), keys, 2, opt | SQ_CONVERT)
            got = meth(s)
            if len(exp) == 0:
                assert got is None
            else:
                assert [tuple(int(x) for x in t) for t in got.matchlist] == [tuple(int(x) for x in e[1:]) for e in exp]


@pytest.mark.parametrize("mode,opt", [("first", SQ_FIRST), ("best", SQ_BEST), ("all", SQ_ALL)])
@pytest.mark.parametrize("nondna", [0, 1])
def test_match_batch_against_oracle(seeq, oracle, mode, opt, nondna):
    pattern, tau = "A[CG]TNNGATC", 1
    matcher = seeq.compile(pattern, tau, nondna)
    keys, _ = oracle.parse(pattern)
    rng = random.Random(11 + nondna)
    lines = []
    for _ in range(3000):
        n = rng.randint(0, 120)
        lines.append("".join(rng.choice("ACGTNacgtRY-") if rng.random() < 0.05 else rng.choice("ACGT") for _ in range(n)))
    text = ("\n".join(lines) + "\n").encode()
    nd = SQ_CONVERT if nondna == 0 else SQ_IGNORE          # seeqmodule.c:1071-1074
    exp, nl, nm = oracle.buffer_scan(np.frombuffer(text, np.uint8), keys, tau, opt | nd)
    exp_t = [(int(a) - 1, int(b), int(c), int(d)) for a, b, c, d in exp]
    got = matcher.matchBatch(text, mode=mode)
    assert got == exp_t
    raw = np.frombuffer(matcher.matchBatch(bytearray(text), mode, True), dtype=np.uint32).reshape(-1, 4)
    assert [tuple(int(x) for x in r) for r in raw] == exp_t
    assert matcher.matchBatch(lines, mode=mode) == exp_t                     # a list of str, joined here
    assert matcher.matchBatch(tuple(s.encode() for s in lines), mode) == exp_t


def test_match_batch_argument_errors(seeq):
    matcher = seeq.compile("ACGTACGT", 1)
    with pytest.raises(seeq.exception):
        matcher.matchBatch(b"ACGT\n", mode="fastest")
    with pytest.raises(TypeError):
        matcher.matchBatch(42)
    with pytest.raises(TypeError):
        matcher.matchBatch([1, 2])
    assert matcher.matchBatch(b"") == []
