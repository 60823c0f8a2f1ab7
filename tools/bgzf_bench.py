#!/usr/bin/env python
"""tools/bgzf_bench.py -- the metric's scan fed with a BGZF (bgzip) buffer in pinned host memory (sqbScanHostBgzf:
compressed bytes over the PCIe link, k0_inflate_bgzf, scan of the inflated text where it lies), next to sqbScanHost
of the same text uncompressed.  GB/s are of TEXT bytes.  ctypes only (no torch): run on the GPU box.

  python tools/bgzf_bench.py [--mb 64] [--copies 16] [--steps 5] [--level 6] [--out gpurun_out/x.json]

The text is the synthetic read stream of bench.py's `metric` workload: --mb MiB of it deflated by Python's zlib
member by member (bgzip's own cut of 0xff00 bytes; the image holds no bgzip binary), the members repeated --copies
times (members are independent: the text repeats with them).  Results are checked: same line / match / record counts
as the scan of the plain text, and the inflated text is compared byte for byte."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(mb=64, copies=16, steps=5, level=6, workload="metric", kernel="", inflate_only=False, sweep=True):
    """One BGZF buffer of the workload's reads through sqbScanHostBgzf next to sqbScanHost of the same text;
    returns the dict of tools/bgzf_bench.py's JSON line.  Raises on any difference in text or counts."""
    import bench
    from seeq_b200 import bgzf, binding as B
    L = B.lib()
    w = bench.WORKLOADS[workload]
    g = B.make_gen(**w["gen"])
    rec = L.sqbGenBytes(C.byref(g), 0, 1)
    reads = (mb << 20) // rec
    text1 = B.gen_host(g, reads)
    t0 = time.perf_counter()
    gz1 = bgzf.compress(text1.tobytes(), level=level, processes=min(32, os.cpu_count() or 1), eof=False)
    t_deflate = time.perf_counter() - t0
    gz = gz1 * copies + bgzf.EOF_MEMBER
    n_text = text1.size * copies
    n_gz = len(gz)

    h_gz = L.sqbHostAlloc(n_gz + 64)
    h_text = L.sqbHostAlloc(n_text + 64)
    assert h_gz and h_text, B.last_error()
    C.memmove(h_gz, gz, n_gz)
    for c in range(copies):
        C.memmove(h_text + c * text1.size, text1.ctypes.data, text1.size)
    del gz

    members, cnt, tb = B.bgzf_index(np.ctypeslib.as_array((C.c_uint8 * n_gz).from_address(h_gz)))
    assert tb == n_text

    # the inflate kernel alone, compressed bytes resident in HBM
    d_gz = L.sqbDeviceAlloc(n_gz + 64)
    d_text = L.sqbDeviceAlloc(n_text + 64)
    assert d_gz and d_text, B.last_error()
    assert L.sqbMemcpyH2D(d_gz, h_gz, n_gz) == 0
    kms, kms_single = [], []
    env0 = os.environ.get("SEEQ_B200_BGZF_KERNEL")
    for name, dst in (("single", kms_single), ("pair", kms)):
        if inflate_only and name == "single":
            continue
        os.environ["SEEQ_B200_BGZF_KERNEL"] = name
        for _ in range(1 if inflate_only else 4):
            ms = C.c_double(0)
            assert L.sqbBgzfInflateDevice(0, d_gz, members, cnt, d_text, None, C.byref(ms)) == 0, B.last_error()
            dst.append(ms.value)
    if kernel:
        os.environ["SEEQ_B200_BGZF_KERNEL"] = kernel
    elif env0 is None:
        del os.environ["SEEQ_B200_BGZF_KERNEL"]
    else:
        os.environ["SEEQ_B200_BGZF_KERNEL"] = env0
    back = np.empty(n_text, dtype=np.uint8)
    assert L.sqbMemcpyD2H(back.ctypes.data, d_text, n_text) == 0
    same_text = bool(np.array_equal(back[:text1.size], text1) and np.array_equal(back[-text1.size:], text1))
    del back
    L.sqbDeviceFree(d_gz)
    L.sqbDeviceFree(d_text)
    if inflate_only:
        L.sqbHostFree(h_gz)
        L.sqbHostFree(h_text)
        return {"inflate_ms": kms, "same_text": same_text}

    sq = B.Seeq(w["pattern"], w["tau"])
    eng = B.Engine.borrowed(sq.engine())
    opt = w["options"] | (B.SQB_COUNT_ONLY if w["count"] else 0)

    def timed(fn):
        for _ in range(2):
            st = fn()
        t = []
        for _ in range(steps):
            t0 = time.perf_counter()
            st = fn()
            t.append(time.perf_counter() - t0)
        return st, t

    st_p, t_p = timed(lambda: eng.scan_host_ptr(h_text, n_text, opt))
    by_slice = {}
    slice0 = os.environ.get("SEEQ_B200_BGZF_SLICE_MB")
    for smb in ((8, 16) if sweep else ()):
        os.environ["SEEQ_B200_BGZF_SLICE_MB"] = str(smb)
        _, t = timed(lambda: eng.scan_host_bgzf_ptr(h_gz, n_gz, opt))
        by_slice[str(smb)] = n_text / (sum(t) / len(t)) / 1e9
    if slice0 is None:
        os.environ.pop("SEEQ_B200_BGZF_SLICE_MB", None)
    else:
        os.environ["SEEQ_B200_BGZF_SLICE_MB"] = slice0
    st_z, t_z = timed(lambda: eng.scan_host_bgzf_ptr(h_gz, n_gz, opt))          # the library's default slices
    same = (st_p.nlines, st_p.nmatched, st_p.nrecs) == (st_z.nlines, st_z.nmatched, st_z.nrecs)
    raw = []
    d = L.sqbDeviceAlloc(n_gz + 64)
    for _ in range(3):
        t0 = time.perf_counter()
        assert L.sqbMemcpyH2D(d, h_gz, n_gz) == 0
        raw.append(time.perf_counter() - t0)
    L.sqbDeviceFree(d)
    gbps = lambda t: n_text / (sum(t) / len(t)) / 1e9
    out = {
        "workload": w["desc"], "text_bytes": n_text, "bgzf_bytes": n_gz, "ratio": n_gz / n_text, "members": cnt,
        "zlib_level": level, "deflate_s_for_one_copy": t_deflate, "copies": copies, "steps": steps,
        "e2e_plain_GBps": gbps(t_p), "e2e_bgzf_GBps_of_text": gbps(t_z), "speedup": gbps(t_z) / gbps(t_p),
        "e2e_bgzf_GBps_by_slice_mb": by_slice,
        "e2e_plain_ms": [x * 1e3 for x in t_p], "e2e_bgzf_ms": [x * 1e3 for x in t_z],
        "inflate_kernel_ms": kms, "inflate_kernel_GBps_of_text": n_text / (min(kms) * 1e-3) / 1e9,
        "inflate_single_kernel_ms": kms_single,
        "inflate_single_kernel_GBps_of_text": n_text / (min(kms_single) * 1e-3) / 1e9,
        "e2e_kernel": os.environ.get("SEEQ_B200_BGZF_KERNEL", "pair"),
        "h2d_of_the_bgzf_bytes_ms": min(raw) * 1e3, "h2d_GBps": n_gz / min(raw) / 1e9,
        "link_bound_GBps_of_text": n_text / min(raw) / 1e9,
        "same_counts": same, "same_text": same_text,
        "nlines": int(st_z.nlines), "nmatched": int(st_z.nmatched), "nrecs": int(st_z.nrecs),
        "launches_per_scan": int(st_z.launches),
        "api": "sqbScanHostBgzf (pinned BGZF buffer -> H2D in slices -> k0_inflate_bgzf_pair -> scan of the text in HBM -> records D2H)",
    }
    L.sqbHostFree(h_gz)
    L.sqbHostFree(h_text)
    sq.close()
    assert same and same_text, out
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=64)
    ap.add_argument("--copies", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--workload", default="metric")
    ap.add_argument("--out", default="")
    ap.add_argument("--kernel", default="", help="pair | single for the end-to-end runs (default: the library's)")
    ap.add_argument("--inflate-only", action="store_true", help="one inflate of the whole buffer (for ncu)")
    a = ap.parse_args()
    out = measure(a.mb, a.copies, a.steps, a.level, a.workload, a.kernel, a.inflate_only)
    line = json.dumps(out)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
