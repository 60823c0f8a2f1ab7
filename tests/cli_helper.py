"""Runs seeq() (the CLI formatter of libseeq_b200.so) in this process so that the
parent test can capture the C-level stdout.  argv[1] = JSON {pattern, input, args}.
Prints `\\nRC=<rc> SEEQERR=<n>` on stderr."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from seeq_b200 import binding as B  # noqa: E402


def main():
    job = json.loads(sys.argv[1])
    L = B.lib()
    a = B.SeeqArgT()
    for k, v in job["args"].items():
        setattr(a, k, v)
    sys.stdout.flush()
    rc = L.seeq(job["pattern"].encode(), job["input"].encode(), a)
    import ctypes
    ctypes.CDLL(None).fflush(None)
    sys.stderr.write("\nRC=%d SEEQERR=%d\n" % (rc, B.seeqerr()))


if __name__ == "__main__":
    main()
