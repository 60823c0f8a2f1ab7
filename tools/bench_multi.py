#!/usr/bin/env python
"""Pattern sets: P patterns over one pass of the text against P separate scans.
Device-resident and end to end (host buffers), cfg1-shaped reads (150 nt), 12-mers at d = 2."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seeq_b200 import binding as B  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
L = B.lib()
g = B.make_gen(seed=11, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=102, max_edits=2)
rec = L.sqbGenBytes(C.byref(g), 0, 1)
nbytes = rec * reads
d_text = L.sqbDeviceAlloc(nbytes + 64)
assert L.sqbGenDevice(C.byref(g), 0, reads, d_text, None) == 0
h_text = L.sqbHostAlloc(nbytes + 64)
assert L.sqbMemcpyD2H(h_text, d_text, nbytes) == 0
rng = np.random.default_rng(5)
code = {"A": 1, "C": 2, "G": 4, "T": 8}
pats = ["GATCGGAAGAGC"] + ["".join("ACGT"[i] for i in rng.integers(0, 4, 12)) for _ in range(15)]
keys = [bytes(code[c] for c in p) for p in pats]
out = []
for P in (1, 2, 4, 8, 16):
    mp = B.Multi(keys[:P], [2] * P)
    engs = [B.Engine(k, 2) for k in keys[:P]]
    for e in engs:
        e.scan_device_large(d_text, nbytes, B.SQ_BEST)
    res = {}
    for name, fn_multi, fn_sep in (
            ("device", lambda: mp.scan_device(d_text, nbytes, B.SQ_BEST),
             lambda: [e.scan_device_large(d_text, nbytes, B.SQ_BEST) for e in engs]),
            ("host", lambda: mp.scan_host_ptr(h_text, nbytes, B.SQ_BEST),
             lambda: [e.scan_host_ptr(h_text, nbytes, B.SQ_BEST) for e in engs])):
        for fn, tag in ((fn_multi, "set"), (fn_sep, "separate")):
            fn(); fn()
            t0 = time.perf_counter()
            for _ in range(5):
                st = fn()
            t = (time.perf_counter() - t0) / 5
            res[name + "_" + tag + "_ms"] = t * 1e3
            res[name + "_" + tag + "_GBps_x_patterns"] = nbytes * P / t / 1e9
        if name == "device":
            a = mp.scan_device(d_text, nbytes, B.SQ_BEST)
            b = [e.scan_device_large(d_text, nbytes, B.SQ_BEST) for e in engs]
            assert [(x.nlines, x.nmatched, x.nrecs) for x in a] == [(x.nlines, x.nmatched, x.nrecs) for x in b]
    res["patterns"] = P
    res["matched"] = [int(x.nmatched) for x in a]
    out.append(res)
    print(json.dumps(res), flush=True)
    mp.close()
    for e in engs:
        e.close()
