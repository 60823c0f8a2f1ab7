#!/usr/bin/env python
"""profiles/<tag>_sass_summary.txt: per kernel of seeq_b200/libseeq_b200.so the static SASS instruction mix and the
instructions that prove the mechanisms DESIGN.md names (UBLKCP = 1-D bulk / TMA copy, SYNCS = mbarrier, ...).
Runs here (no GPU): cuobjdump -sass + c++filt.

  python tools/sass_summary.py profiles/r4_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

WANT = ("k12_scan_pack", "k2_bitslice<20, 1, 1, false, 3, true>", "k2_bitslice<10, 1, 1, false, 2, true>",
        "k2_bitslice<26, 4, 1, false, 0, false>", "k1_scan_classify<true, false, false>", "k15_pack", "k34_finish_lines<1, 4>",
        "k1_gather", "k1_scan_tiles", "k0_inflate_bgzf_pair", "k0_inflate_bgzf(")
EVIDENCE = ("UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "PRMT", "LOP3", "IMAD", "SHFL", "VOTE", "ATOMG", "ATOMS", "REDUX",
            "BAR", "NANOSLEEP")


def main():
    so = "seeq_b200/libseeq_b200.so"
    out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    fn, per = None, collections.OrderedDict()
    arch = re.search(r"arch = (sm_\w+)", out)
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            per[fn][m.group(1)] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(per.keys()), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    lines = ["# SASS of %s (cuobjdump -sass), %s: per kernel the static instruction mix (top opcodes) and the" % (so, arch.group(1) if arch else "?"),
             "# instructions that prove the mechanisms DESIGN.md names: UBLKCP (1-D bulk / TMA copies), SYNCS (mbarrier), LDS / STS,",
             "# PRMT, LOP3, IMAD, SHFL, VOTE.  No tensor-core instruction anywhere: the path is integer bit logic.  tools/sass_summary.py", ""]
    for name, d in zip(per.keys(), dem):
        if not any(w in d for w in WANT):
            continue
        c = per[name]
        lines.append(d[:150])
        lines.append("   %d instructions: " % sum(c.values()) + ", ".join("%s %d" % kv for kv in c.most_common(14)))
        lines.append("   evidence: " + ", ".join("%s %d" % (k, c[k]) for k in EVIDENCE if c.get(k)))
    tc = sum(c.get(k, 0) for c in per.values() for k in ("HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA", "BMMA", "DMMA"))
    lines.append("")
    lines.append("tensor-core instructions in the whole library: %d" % tc)
    open(sys.argv[1] if len(sys.argv) > 1 else "profiles/sass_summary.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
