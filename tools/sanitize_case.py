#!/usr/bin/env python
"""A reduced tour of the kernels for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
every kernel family runs at least once on small inputs and is checked against the oracle.

  compute-sanitizer --tool racecheck python tools/sanitize_case.py      (tools/gpu_sanitize.sh)

K1's TMA stage is overwritten in place and released with fence.proxy.async + a bulk store; the fused
tokenise+pack kernel reuses its text stage for nibbles and planes: exactly the code these tools exist for.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SEEQ_B200_CHUNK_MB", "1")          # several chunks through the two slots

from oracle import pyoracle                                 # noqa: E402
from seeq_b200 import binding as B                          # noqa: E402


def check(orc, text, pattern, tau, opt, label, extra=0):
    sq = B.Seeq(pattern, tau)
    st = B.StatsT()
    recs = sq.batch(text, opt | extra, B.SQ_ANY, st)
    exp, nl, nm = orc.buffer_scan(text, sq.keys, tau, opt)
    got = np.stack([recs["line"].astype(np.uint64) + 1, recs["start"], recs["end"], recs["dist"]],
                   axis=1).astype(np.uint64) if recs.size else np.zeros((0, 4), np.uint64)
    ok = (st.nlines, st.nmatched) == (nl, nm) and np.array_equal(got, exp)
    print("%-44s lines %7d records %7d path %x launches %3d %s" % (label, st.nlines, len(got), st.path, st.launches,
                                                                   "ok" if ok else "MISMATCH"))
    sq.close()
    return ok


def main():
    pyoracle.build()
    orc = pyoracle.Oracle()
    nreads = int(os.environ.get("SANITIZE_READS", "30000"))
    ok = True
    short = B.gen_host(B.make_gen(seed=2, line_len=150, n_per_1024=5), nreads)
    fastq = B.gen_host(B.make_gen(seed=5, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=307, max_edits=2, fastq=True),
                       nreads // 3)
    long_ = B.gen_host(B.make_gen(seed=3, line_len=10000, plant="ACGTTGCAAGCTTAGGCATCGATCGGATCAGCTAGCTAGC", plant_per_1024=1024,
                                  max_edits=4), 300)
    for matcher in ("bitslice", "word"):
        os.environ["SEEQ_B200_MATCHER"] = matcher
        ok &= check(orc, short, "A[CG]TNNGATC", 1, B.SQ_BEST, matcher + ": short reads, best, NFA levels")
        ok &= check(orc, short, "GATCGGAAGAGC", 2, B.SQ_ALL | B.SQ_IGNORE, matcher + ": short reads, all, -x 2")
        ok &= check(orc, short, "ACGTACGTTGCATGCAAGCTTAGCTAGGATCCATGGCATGCAAGCTTGGCACTGGCCGTCGTTTTACAAC", 6,
                    B.SQ_FIRST | B.SQ_CONVERT, matcher + ": 70-mer, first, -x 1 (two parts)")
        ok &= check(orc, fastq, "GATCGGAAGAGC", 2, B.SQ_FIRST, matcher + ": FASTQ-like, first (line filter)")
        ok &= check(orc, long_, "ACGTTGCAAGCTTAGGCATCGATCGGATCAGCTAGCTAGC", 4, B.SQ_ALL, matcher + ": 10-kb lines, all (cuts)")
    print("sanitize tour", "ok" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
