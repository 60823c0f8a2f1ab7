#!/usr/bin/env python
"""Dynamic SASS opcode histogram of one kernel of an `ncu --set full --import-source on` report.

  python tools/ncu_dynhist.py gpurun_out/r1x_cfg2_full.ncu-rep k1_scan_classify [--top 24] [--lines]

Sums the "Instructions Executed" column of the source page per opcode (warp instructions, all
launches of the first kernel whose name matches).  --lines prints the hottest SASS lines instead.
Runs here (no GPU needed): it only reads the report with `ncu -i`.
"""
from __future__ import annotations

import argparse
import collections
import csv
import re
import subprocess


def blocks_of(rep: str, pat: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    i = 0
    while i < len(lines):
        if lines[i].startswith('"Kernel Name"'):
            name = next(csv.reader([lines[i]]))[1]
            hdr = next(csv.reader([lines[i + 1]]))
            blk = []
            i += 2
            while i < len(lines) and not lines[i].startswith('"Kernel Name"'):
                if lines[i].strip():
                    blk.append(next(csv.reader([lines[i]])))
                i += 1
            yield name, hdr, blk
        else:
            i += 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=24)
    ap.add_argument("--lines", action="store_true")
    args = ap.parse_args()
    for name, hdr, blk in blocks_of(args.report, args.kernel):
        ie, src = hdr.index("Instructions Executed"), hdr.index("Source")
        smp = hdr.index("# Samples") if "# Samples" in hdr else None
        tot = sum(int(r[ie]) for r in blk)
        print(name[:110], "| warp instructions", tot)
        if args.lines:
            rows = sorted(blk, key=lambda r: -int(r[smp] if smp is not None else r[ie]))[:args.top]
            for r in rows:
                print("  %10d instr %8s samples  %s" % (int(r[ie]), r[smp] if smp is not None else "-", r[src].strip()[:100]))
        else:
            h = collections.Counter()
            for r in blk:
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[src])
                h[m.group(2) if m else "?"] += int(r[ie])
            for op, n in h.most_common(args.top):
                print("  %-10s %12d %5.1f%%" % (op, n, 100.0 * n / max(tot, 1)))
        break


if __name__ == "__main__":
    main()
