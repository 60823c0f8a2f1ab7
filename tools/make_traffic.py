#!/usr/bin/env python
"""profiles/k2_traffic.json from ncu summaries: DRAM bytes (read + written) per launch
of the kernels bench.py reports a roofline for.

  python tools/make_traffic.py cfg2=profiles/r1g_cfg2_ncu.json cfg4=profiles/r1f_cfg4_ncu.json ...
"""
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
NAMES = (("k2_bitslice", "k2_matcher"), ("k2_forward", "k2_matcher"), ("k15_pack", "k15_pack"),
         ("k1_scan_classify", "k1_scan_classify"), ("k12_scan_pack", "k12_scan_pack"), ("k34_finish", "k34_finish"))


def main():
    out = {}
    for arg in sys.argv[1:]:
        wl, path = arg.split("=")
        per = {}
        pipes = {}
        for k in json.load(open(path))["kernels"]:
            for needle, name in NAMES:
                if needle in k["kernel"]:
                    b = (k["dram_read"] * UNIT[k.get("dram_read_unit", "byte")] +
                         k["dram_write"] * UNIT[k.get("dram_write_unit", "byte")])
                    if b > per.get(name, 0):          # the gated no-op launch of the other matcher moves nothing
                        per[name] = int(b)
                        pipes[name] = {"alu_pipe_pct": round(k.get("alu_pipe_pct", 0.0), 1),
                                       "issue_active_pct": round(k.get("issue_active_pct", 0.0), 1),
                                       "dram_pct_of_peak": round(k.get("dram_pct_of_peak", 0.0), 1),
                                       "warp_instructions": int(k.get("warp_instructions", 0))}
        out[wl] = per
        out.setdefault("_pipes", {})[wl] = pipes
    out["_source"] = {a.split("=")[0]: a.split("=")[1] for a in sys.argv[1:]}
    json.dump(out, open("profiles/k2_traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
