#!/usr/bin/env python
"""Wall time of the re-linked CLI (seeq_b200/_relink/seeq = the reference's own seeq-main.c linked
against libseeq_b200.so) against the unmodified reference CLI (oracle/_ref/seeq_ref) on files of
BASELINE.json's five configurations, with the flags the configurations name.  Run on the GPU box:

  python tools/cli_times.py [--scale 1.0] [--configs cfg1,cfg2,...] > gpurun_out/<tag>_cli_times.json

The files are written to /dev/shm (page cache: no disk in the numbers).  Both programs write to a file in
/dev/shm as well; the outputs are compared byte for byte (sha256) -- a row whose hashes differ is a FAILED
row, whatever the speed-up.  The reference runs on ONE core (it is single-threaded; the N-process numbers are
bench.py's `--impl reference`).  --scale shrinks the read counts (the reference needs minutes on cfg4).
SEEQ_B200_DEVICES=all lets the re-linked CLI use every visible GPU (one host thread + engine per device).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402  (workload table)
from seeq_b200 import binding as B  # noqa: E402

OURS = os.path.join(ROOT, "seeq_b200", "_relink", "seeq")
REF = os.path.join(ROOT, "oracle", "_ref", "seeq_ref")

# CLI flags of the five configurations (BASELINE.json `configs`)
FLAGS = {
    "cfg1": lambda w: ["-c", "-d", "2", w["pattern"]],
    "cfg2": lambda w: ["-b", "-l", "-p", "-k", "-d", "1", w["pattern"]],
    "cfg3": lambda w: ["-a", "-f", "-d", "4", w["pattern"]],
    "cfg4": lambda w: ["-b", "-x", "1", "-d", "8", w["pattern"]],
    "cfg5": lambda w: ["-e", "-d", "2", w["pattern"]],
    "metric": lambda w: ["-b", "-l", "-p", "-k", "-d", "2", w["pattern"]],
}


def sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def run(exe: str, flags, infile: str, outfile: str, env=None, repeat: int = 1):
    best = None
    for _ in range(repeat):
        with open(outfile, "wb") as out:
            t0 = time.perf_counter()
            p = subprocess.run([exe, *flags, infile], stdout=out, stderr=subprocess.PIPE, env=env)
            dt = time.perf_counter() - t0
        if p.returncode not in (0, 1):
            raise RuntimeError("%s %s: rc %d %s" % (exe, flags, p.returncode, p.stderr[-300:]))
        best = dt if best is None else min(best, dt)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--configs", default="cfg1,cfg2,cfg3,cfg4,cfg5")
    ap.add_argument("--ref-scale", type=float, default=None,
                    help="the reference runs on the first ref_scale of the file's reads (time extrapolated linearly; the "
                         "outputs are then compared on that prefix, with our CLI run on it as well)")
    ap.add_argument("--tmp", default="/dev/shm")
    a = ap.parse_args()
    rows = []
    # what every CUDA process pays before its first kernel (context creation: 1.1 - 3.5 s on the boxes measured,
    # see DESIGN.md) -- measured on a file of 1000 reads, best of 3
    g0 = B.make_gen(seed=1, line_len=150)
    tiny = os.path.join(a.tmp, "sqb_tiny.txt")
    B.gen_host(g0, 1000).tofile(tiny)
    startup = run(OURS, ["-c", "-d", "1", "GATTACA"], tiny, tiny + ".out", repeat=3)
    os.remove(tiny)
    os.remove(tiny + ".out")
    print(json.dumps({"startup_s": startup, "note": "re-linked CLI on 1000 reads: CUDA context creation + library load"}), flush=True)
    for name in a.configs.split(","):
        w = dict(bench.WORKLOADS[name])
        reads = max(1000, int(w["reads"] * a.scale))
        g = B.make_gen(**w["gen"])
        text = B.gen_host(g, reads)
        infile = os.path.join(a.tmp, "sqb_%s.txt" % name)
        text.tofile(infile)
        flags = FLAGS[name](w)
        out_o, out_r = infile + ".ours", infile + ".ref"
        row = {"config": name, "flags": " ".join(flags[:-1]) + " <pattern m=%d>" % len(w["pattern"].replace("[CG]", "C")),
               "reads": reads, "bytes": int(text.size)}
        # ours: first run includes CUDA context creation and the first-scan capacity guesses; report both
        row["ours_first_s"] = run(OURS, flags, infile, out_o)
        row["ours_s"] = run(OURS, flags, infile, out_o, repeat=3)
        row["startup_s"] = startup
        row["ours_net_s"] = max(row["ours_s"] - startup, 1e-3)
        env_all = dict(os.environ, SEEQ_B200_DEVICES="all")
        try:
            import torch
            ngpu = torch.cuda.device_count()
        except Exception:
            ngpu = 1
        if ngpu > 1:
            row["ours_all_gpus_s"] = run(OURS, flags, infile, out_o + ".all", env=env_all, repeat=2)
            row["gpus"] = ngpu
            row["all_gpus_same_output"] = sha(out_o + ".all") == sha(out_o)
            os.remove(out_o + ".all")
        # the reference, on the whole file or on a prefix of it
        if a.ref_scale is not None and a.ref_scale < 1.0:
            nref = max(1000, int(reads * a.ref_scale))
            sub = B.gen_host(g, nref)
            subfile = infile + ".sub"
            sub.tofile(subfile)
            row["ref_reads"] = nref
            row["ref_s_measured"] = run(REF, flags, subfile, out_r)
            row["ref_s"] = row["ref_s_measured"] * reads / nref
            run(OURS, flags, subfile, out_o)
            os.remove(subfile)
        else:
            row["ref_s"] = run(REF, flags, infile, out_r)
        row["same_output"] = sha(out_o) == sha(out_r)
        row["output_bytes"] = os.path.getsize(out_r)
        row["speedup"] = row["ref_s"] / row["ours_s"]
        row["speedup_net_of_startup"] = row["ref_s"] / row["ours_net_s"]
        row["ours_net_GBps"] = text.size / row["ours_net_s"] / 1e9
        row["ours_GBps"] = text.size / row["ours_s"] / 1e9
        row["ref_GBps"] = text.size / row["ref_s"] / 1e9
        for f in (infile, out_o, out_r):
            if os.path.exists(f):
                os.remove(f)
        rows.append(row)
        print(json.dumps(row), flush=True)
    ok = all(r["same_output"] for r in rows)
    print(json.dumps({"summary": True, "all_outputs_identical": ok, "rows": len(rows)}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
