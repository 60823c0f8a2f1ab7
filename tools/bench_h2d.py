#!/usr/bin/env python
"""Raw pinned H2D copy bandwidth of the box next to sqbScanHost at several chunk sizes (cfg2 reads)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from seeq_b200 import binding as B  # noqa: E402

L = B.lib()
reads = 10_000_000
g = B.make_gen(seed=2, line_len=150, n_per_1024=5)
nbytes = L.sqbGenBytes(C.byref(g), 0, 1) * reads
d = L.sqbDeviceAlloc(nbytes + 64)
assert L.sqbGenDevice(C.byref(g), 0, reads, d, None) == 0
h = L.sqbHostAlloc(nbytes + 64)
assert L.sqbMemcpyD2H(h, d, nbytes) == 0
out = {}
# raw copy
for _ in range(2):
    L.sqbMemcpyH2D(d, h, nbytes)
t0 = time.perf_counter()
for _ in range(5):
    L.sqbMemcpyH2D(d, h, nbytes)
out["raw_h2d_GBps"] = nbytes * 5 / (time.perf_counter() - t0) / 1e9
keys = bytes([1, 6, 8, 31, 31, 4, 1, 8, 2])          # A[CG]TNNGATC
for mb in (16, 32, 64, 128, 256, 512):
    os.environ["SEEQ_B200_CHUNK_MB"] = str(mb)
    eng = B.Engine(keys, 1)
    for _ in range(2):
        eng.scan_host_ptr(h, nbytes, B.SQ_BEST)
    t0 = time.perf_counter()
    for _ in range(5):
        st = eng.scan_host_ptr(h, nbytes, B.SQ_BEST)
    out["e2e_chunk_%dMB_GBps" % mb] = nbytes * 5 / (time.perf_counter() - t0) / 1e9
    eng.close()
print(json.dumps(out))
