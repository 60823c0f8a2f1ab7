#!/usr/bin/env python
"""The BGZF inflate kernels for compute-sanitizer (memcheck / racecheck): both kernels over small buffers of every
block type -- dynamic and fixed codes, stored blocks, several blocks per member, literal-only and match-heavy text,
an odd number of members -- checked against zlib, and one damaged member (refused).

  compute-sanitizer --tool racecheck python tools/sanitize_bgzf.py      (tools/gpu_bgzf.sh TAG ... sanitize)
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from seeq_b200 import bgzf, binding as B                    # noqa: E402


def main():
    rng = np.random.default_rng(4)
    dna = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=400000)
    dna[150::151] = 10
    dna = dna.tobytes()
    qual = (rng.integers(0, 41, size=200000) + 33).astype(np.uint8).tobytes()
    runs = b"".join(bytes([65 + i % 5]) * (1 + (i * 37) % 500) for i in range(600))
    noise = rng.integers(0, 256, size=70000, dtype=np.uint8).tobytes()
    cases = [("dna level 6", bgzf.compress(dna, level=6), dna),
             ("dna level 1", bgzf.compress(dna[:200000], level=1), dna[:200000]),
             ("dna huffman only", bgzf.compress(dna[:150000], strategy=zlib.Z_HUFFMAN_ONLY), dna[:150000]),
             ("quality strings", bgzf.compress(qual, level=6), qual),
             ("fixed code", bgzf.compress(qual[:60000], strategy=zlib.Z_FIXED), qual[:60000]),
             ("runs", bgzf.compress(runs, level=9), runs),
             ("stored", bgzf.compress(noise, level=6) , noise),
             ("small members", bgzf.compress(dna[:30000], level=6, block=777), dna[:30000])]
    bad = 0
    for kernel in ("pair", "single"):
        os.environ["SEEQ_B200_BGZF_KERNEL"] = kernel
        for label, gz, text in cases:
            out, ms = B.bgzf_inflate_device(gz)
            ok = out.tobytes() == text
            bad += not ok
            print("%-8s %-18s %7d -> %7d bytes  %s" % (kernel, label, len(gz), len(text), "ok" if ok else "MISMATCH"))
        broken = bytearray(cases[0][1])
        broken[4000:4064] = bytes(64)
        try:
            B.bgzf_inflate_device(bytes(broken))
            print(kernel, "damaged member: NOT refused")
            bad += 1
        except RuntimeError as err:
            print("%-8s damaged member refused: %s" % (kernel, err))
    print("sanitize bgzf tour:", "all ok" if not bad else "%d MISMATCH" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
