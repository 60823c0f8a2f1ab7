"""Golden fixtures (tests/golden/*.npz, written from the UNMODIFIED reference by
tests/golden/make_golden.py) against the oracle (CPU) and the CUDA path (GPU).

The inputs are regenerated with the pinned host generator and checked against the
SHA-256 stored in the fixture, so a drifting generator cannot silently change the
test.  Bit-exact comparison of (line, start, end, dist) and of the line counts.
"""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests.golden.cases import CASES, CLI_CASES, make_input

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def B():
    from seeq_b200 import binding
    binding.lib()
    return binding


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def case_input(B, name):
    buf = make_input(B, CASES[name])
    gold = load(name)
    assert buf.size == int(gold["nbytes"])
    assert hashlib.sha256(buf.tobytes()).digest() == gold["sha256"].tobytes(), "generator drifted: " + name
    return buf, gold


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_fixture(B, oracle, name):
    buf, gold = case_input(B, name)
    case = CASES[name]
    keys, err = oracle.parse(case["pattern"])
    assert keys is not None, err
    for opt in case["options"]:
        recs, nl, nm = oracle.buffer_scan(buf, keys, case["tau"], opt)
        assert (nl, nm) == (int(gold["nlines_%d" % opt]), int(gold["nmatched_%d" % opt])), (name, opt)
        assert np.array_equal(recs.astype(np.uint32), gold["recs_%d" % opt]), (name, opt)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_reference_fixture(B, name, matcher):
    buf, gold = case_input(B, name)
    case = CASES[name]
    sq = B.Seeq(case["pattern"], case["tau"])
    fasta = B.SQB_FASTA if buf[0] == ord(">") else 0
    for opt in case["options"]:
        st = B.StatsT()
        r = sq.batch(buf, opt | fasta, B.SQ_ANY, st)
        got = np.stack([r["line"] + 1, r["start"], r["end"], r["dist"]], axis=1).astype(np.uint32) \
            if r.size else np.zeros((0, 4), np.uint32)
        exp = gold["recs_%d" % opt]
        assert (st.nlines, st.nmatched) == (int(gold["nlines_%d" % opt]), int(gold["nmatched_%d" % opt])), (name, opt)
        assert np.array_equal(got, exp), (name, opt)
        # the count-only entry points agree with the record path
        assert sq.batch(buf, (opt & 0xC) | fasta, B.SQ_COUNTLINES) == int(gold["nmatched_%d" % opt])
        if opt & 3 == 2:
            assert sq.batch(buf, (opt & 0xC) | fasta, B.SQ_COUNTMATCH) == exp.shape[0]
    sq.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CLI_CASES))
def test_relinked_cli_text_matches_reference(B, tmp_path, name):
    """The reference's own front-end (seeq-main.c, compiled from the mount, linked
    against libseeq_b200.so) prints byte-identical text to the reference CLI."""
    from seeq_b200 import build
    cli = os.path.join(build.RELINK, "seeq")
    if not os.path.exists(cli):
        pytest.skip("re-linked CLI not built (needs /root/reference at build time)")
    gold = json.load(open(os.path.join(GOLD, "cli.json")))[name]
    case_name, flags = CLI_CASES[name]
    buf, _ = case_input(B, case_name)
    path = tmp_path / (case_name + ".txt")
    buf.tofile(path)
    r = subprocess.run([cli, *flags, CASES[case_name]["pattern"], str(path)], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, stdin=subprocess.DEVNULL, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-500:]
    assert len(r.stdout) == gold["bytes"], (name, r.stdout[:300])
    assert hashlib.sha256(r.stdout).hexdigest() == gold["sha256"], (name, r.stdout[:300], gold["head"])
