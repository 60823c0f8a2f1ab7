#!/usr/bin/env python
"""Summarise an `ncu --set full` report (and optionally a launch list) into a
small tracked file under profiles/.

  python tools/summarise_ncu.py gpurun_out/r1a_cfg2_full.ncu-rep profiles/r1a_cfg2 \
         [--launches gpurun_out/r1a_cfg2_launches.csv] [--note "..."]

Writes <out>_ncu.json (per kernel: duration, DRAM bytes, pipe utilisation,
instruction counts, stall reasons) and <out>_launches.csv (kernel, duration).
Runs here (no GPU needed): it only reads the report with `ncu -i`.
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "launch__shared_mem_per_block_static": "static_smem_per_block",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__inst_executed_op_branch.sum": "branch_instructions",
    "smsp__sass_inst_executed_op_shared_ld.sum": "shared_loads",
    "smsp__sass_inst_executed_op_shared_st.sum": "shared_stores",
    "smsp__sass_inst_executed_op_global_ld.sum": "global_loads",
    "smsp__sass_inst_executed_op_global_st.sum": "global_stores",
    "smsp__sass_inst_executed_op_local_ld.sum": "local_loads",
    "smsp__sass_inst_executed_op_local_st.sum": "local_stores",
    "smsp__sass_inst_executed_op_tma_ld.sum": "tma_loads",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "sm__cycles_elapsed.max": "sm_cycles",
}
STALL = "smsp__average_warps_issue_stalled_"


def raw_page(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def num(x: str):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--launches")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    hdr, units, rows = raw_page(a.report)
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows:
        k = {"kernel": r[idx["Kernel Name"]]}
        for h, name in KEEP.items():
            if h in idx:
                k[name] = num(r[idx[h]])
                if units[idx[h]]:
                    k[name + "_unit"] = units[idx[h]]
        stalls = {}
        for h, i in idx.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                stalls[h[len(STALL):-len("_per_issue_active.ratio")]] = num(r[i])
        k["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1] if isinstance(kv[1], float) else 0)[:6])
        kernels.append(k)
    doc = {"source": a.report, "tool": "ncu --set full --clock-control none --import-source on", "note": a.note,
           "kernels": kernels}
    with open(a.out + "_ncu.json", "w") as f:
        json.dump(doc, f, indent=1)
    if a.launches:
        lines = [l for l in open(a.launches) if l.startswith('"')]
        rd = list(csv.reader(lines))
        h = {n: i for i, n in enumerate(rd[0])}
        with open(a.out + "_launches.csv", "w") as f:
            f.write("id,kernel,grid,block,duration_ns\n")
            for r in rd[1:]:
                f.write("%s,\"%s\",\"%s\",\"%s\",%s\n" % (r[h["ID"]], r[h["Kernel Name"]], r[h["Grid Size"]],
                                                        r[h["Block Size"]], r[h["Metric Value"]]))
    print("wrote", a.out + "_ncu.json")


if __name__ == "__main__":
    sys.exit(main())
