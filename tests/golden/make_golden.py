#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference -> oracle/_ref/libseeq_ref.so):

    python tests/golden/make_golden.py

Each fixture is a scaled-down BASELINE.json config: the input bytes come from the
pinned counter-based generator (seeq_b200/csrc/sqb_gen.h, host side, no GPU), the
expected records (1-based line, start, end, dist), line count and matched-line
count come from the reference's own seeqFileMatch(SQ_ANY) loop through
oracle/ref_driver.c.  The fixture stores the generator parameters, a SHA-256 of
the input bytes (so a drifting generator is detected) and the reference output.
The GPU box has no /root/reference: tests read these files only.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from seeq_b200 import binding as B  # noqa: E402
from tests.golden.cases import CASES, CLI_CASES, make_input  # noqa: E402


def main():
    pyoracle.build()
    assert pyoracle.have_reference(), "oracle/_ref is not built (no /root/reference?)"
    ref = pyoracle.Reference()
    for name, case in CASES.items():
        buf = make_input(B, case)
        out = {}
        for opt in case["options"]:
            recs, nl, nm = ref.buffer_scan(buf, case["pattern"], case["tau"], opt)
            out["recs_%d" % opt] = recs.astype(np.uint32)
            out["nlines_%d" % opt] = np.uint64(nl)
            out["nmatched_%d" % opt] = np.uint64(nm)
        out["sha256"] = np.frombuffer(hashlib.sha256(buf.tobytes()).digest(), dtype=np.uint8)
        out["nbytes"] = np.uint64(buf.size)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, buf.size, "bytes;", {k: (v.shape if hasattr(v, "shape") and v.ndim else int(v))
                                          for k, v in out.items() if k != "sha256"})


def cli_golden():
    """stdout of the reference CLI (oracle/_ref/seeq_ref) -> tests/golden/cli.json"""
    import json
    import subprocess
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (case_name, flags) in CLI_CASES.items():
            case = CASES[case_name]
            path = os.path.join(tmp, case_name + ".txt")
            if not os.path.exists(path):
                make_input(B, case).tofile(path)
            r = subprocess.run([pyoracle.REF_CLI, *flags, case["pattern"], path], stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, stdin=subprocess.DEVNULL, check=True)
            out[name] = {"case": case_name, "flags": flags, "bytes": len(r.stdout),
                         "sha256": hashlib.sha256(r.stdout).hexdigest(),
                         "head": r.stdout[:300].decode("latin-1")}
            print(name, len(r.stdout), "bytes of stdout")
    with open(os.path.join(HERE, "cli.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
    cli_golden()
