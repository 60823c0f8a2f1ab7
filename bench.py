#!/usr/bin/env python
"""bench.py -- throughput of the seeq matching path on B200 (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W [--workload metric] [--reads R] [--impl reference]

A *step* is one pass of the hot path (K12: line scan + class coding + bit-plane pack in one kernel -- or K1
and the pack kernel for multi-part automata, cut lines and filtered scans -> K2 forward matcher -> K3
reverse pass / K4 ordered compaction) over one batch of synthetic reads.

Default workload = `metric`: BASELINE.json's metric is quoted on "d=2, 20-nt pattern": `seeq -b -l -p -k
-d 2 <20-mer>` (best match + positions, the flags of configs[1]) over 10 M synthetic 150-nt reads, 1.51 GB
per GPU, the 20-mer planted in 10 % of the reads with 0-2 edits.  Weak scaling: every rank scans its own
10 M reads of one global read stream.  `--workload cfg1..cfg5` are BASELINE.json's configs; at one GPU the
default run also measures every one of them and reports them under "configs" (a shard of 2 GiB or more
goes through sqbScanDeviceLarge).

One JSON line is printed by rank 0:
  value        GB/s of reads scanned, whole job, input resident in HBM: CUDA events around K steps on the
               stream the scans are queued on, max over ranks; two scans in flight, a repeated scan
               replays as a CUDA graph
  e2e          the same metric through the C-ABI with HOST buffers (sqbScanHost: pinned host text -> H2D
               -> kernels -> records D2H inside the timed region); h2d_raw_GBps next to it = a bare
               cudaMemcpy of the same pinned buffer on all ranks at once (the ceiling the box allows)
  roofline     kernel "step": SURVEY 8(d) algorithmic bytes of the step / its duration against the
               measured HBM copy peak; under "kernels" every kernel of the step with ITS OWN algorithmic
               bytes (what its interface makes it read and write once), from CUDA events the engine
               records around its kernels INSIDE the timed region (every other step); DRAM traffic and
               pipe utilisation from the committed ncu capture (`traffic_source`; not measured in this run)
  cpu_baseline the unmodified reference (oracle/_ref) on the host cores, on a bounded sample of the same
               workload; "parity": the GPU's records of that very sample against the reference's count and
               record checksum (line, start, end, dist of every record)
  bgzf         (one GPU) the same scan fed with a BGZF (bgzip) buffer in pinned host memory through sqbScanHostBgzf:
               GB/s of TEXT end to end, the inflate kernel alone, the bound of the link at this compression ratio
`--impl reference` times the reference's own CPU implementation instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SQ_FIRST, SQ_BEST, SQ_ALL, SQ_CONVERT = 0, 1, 2, 4


def fixed_pattern(seed: int, n: int) -> str:
    rng = np.random.default_rng(seed)
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


# BASELINE.json's metric shape and its five configs; `count` selects the count-only path (CLI -c)
WORKLOADS = {
    "metric": dict(desc="seeq -b -l -p -k -d 2 <20-mer>, 10M x 150 nt (the metric's d=2 / 20-nt shape)",
                   pattern=fixed_pattern(20, 20), tau=2, options=SQ_BEST, count=False, reads=10_000_000,
                   gen=dict(seed=20, line_len=150, plant=fixed_pattern(20, 20), plant_per_1024=102, max_edits=2)),
    "cfg1": dict(desc="seeq -c -d 2 GATCGGAAGAGC, 1M x 150 nt", pattern="GATCGGAAGAGC", tau=2, options=SQ_FIRST,
                 count=True, reads=1_000_000, gen=dict(seed=1, line_len=150, plant="GATCGGAAGAGC",
                                                       plant_per_1024=102, max_edits=2)),
    "cfg2": dict(desc="seeq -b -l -p -k -d 1 A[CG]TNNGATC, 10M x 150 nt", pattern="A[CG]TNNGATC", tau=1,
                 options=SQ_BEST, count=False, reads=10_000_000,
                 gen=dict(seed=2, line_len=150, n_per_1024=5)),
    "cfg3": dict(desc="seeq -a -f -d 4 <40-mer>, 100k x 10 kb", pattern=fixed_pattern(3, 40), tau=4,
                 options=SQ_ALL, count=False, reads=100_000,
                 gen=dict(seed=3, line_len=10_000, plant=fixed_pattern(3, 40), plant_per_1024=1024, max_edits=4)),
    "cfg4": dict(desc="seeq -b -x 1 -d 8 <100-mer>, 10M x 250 nt", pattern=fixed_pattern(4, 100), tau=8,
                 options=SQ_BEST | SQ_CONVERT, count=False, reads=10_000_000,
                 gen=dict(seed=4, line_len=250, plant=fixed_pattern(4, 100), plant_per_1024=102, max_edits=8,
                          junk_per_1024=1)),
    "cfg5": dict(desc="seeq -e -d 2 GATCGGAAGAGC, FASTQ-like 4-line records", pattern="GATCGGAAGAGC", tau=2,
                 options=SQ_FIRST, count=False, reads=4_000_000,
                 gen=dict(seed=5, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=307, max_edits=2, fastq=True)),
    # beyond BASELINE.json: config 5 with -x 1 (every byte is scannable, the dead-on-arrival filter drops nothing),
    # as the reference would scan it, and record-aware (SQB_FASTQ: sequence lines only)
    "cfg5x1": dict(desc="seeq -e -x 1 -d 2 GATCGGAAGAGC, FASTQ-like 4-line records, every line scanned",
                   pattern="GATCGGAAGAGC", tau=2, options=SQ_FIRST | SQ_CONVERT, count=False, reads=4_000_000,
                   gen=dict(seed=5, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=307, max_edits=2, fastq=True)),
    "cfg5x1q": dict(desc="the same, record-aware (SQB_FASTQ: sequence lines only)",
                    pattern="GATCGGAAGAGC", tau=2, options=SQ_FIRST | SQ_CONVERT | 0x4000, count=False, reads=4_000_000,
                    gen=dict(seed=5, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=307, max_edits=2, fastq=True)),
}
BASELINE_CONFIGS = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]


# --------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# --------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# reference arm / cpu baseline
# --------------------------------------------------------------------------
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(w, host_text: np.ndarray, nproc: int, seconds_target: float, steps: int = 1, warmup: int = 0):
    """Time the reference (oracle/_ref) -- or, if it is not built, the oracle port --
    on a bounded sample of host_text.  -> dict(value GB/s, reads/s, kind, cores, sample, result, checksum)."""
    from oracle import pyoracle
    pyoracle.build()
    rec_bytes = w["rec_bytes"]
    mode = 0 if w["count"] else 2
    if pyoracle.have_reference():
        ref = pyoracle.Reference()
        kind = "reference"

        def run(buf, n):
            return ref.bench_ck(buf, w["pattern"], w["tau"], w["options"] & 0x1F, mode, n)
    else:
        orc = pyoracle.Oracle()
        keys, _ = orc.parse(w["pattern"])
        kind = "port"
        nproc = 1

        def run(buf, n):
            t0 = time.perf_counter()
            r, nl, nm = orc.buffer_scan(buf, keys, w["tau"], w["options"] & 0x1F)
            t = time.perf_counter() - t0
            if w["count"]:
                return t, nm, nm
            return t, len(r), pyoracle.records_checksum(r[:, 0], r[:, 1], r[:, 2], r[:, 3])
    # calibrate on ~8 MB, one core
    cal_reads = max(1, min(host_text.size // rec_bytes, (8 << 20) // rec_bytes))
    t_cal = run(host_text[:cal_reads * rec_bytes], 1)[0]
    rate1 = cal_reads * rec_bytes / max(t_cal, 1e-6)                 # bytes/s on one core
    want = int(rate1 * nproc * seconds_target * 0.7)                 # imperfect scaling margin
    reads = max(nproc, min(host_text.size // rec_bytes, want // rec_bytes))
    sample = host_text[:reads * rec_bytes]
    for _ in range(warmup):
        run(sample, nproc)
    times = []
    result = checksum = None
    for _ in range(max(1, steps)):
        t, result, checksum = run(sample, nproc)
        times.append(t)
    # a sample the host cores finish in a fraction of a second is repeated (about 2 s of wall time)
    while steps <= 1 and sum(times) < 2.0 and len(times) < 8:
        t, result, checksum = run(sample, nproc)
        times.append(t)
    t = float(np.mean(times))
    return dict(value=sample.size / t / 1e9, reads_per_s=reads / t, unit="GB/s", cores=nproc, kind=kind,
                one_core_value=rate1 / 1e9,
                sample="%d reads (%.1f MB) of the same workload, %d processes over newline-aligned shards, "
                       "%.2f s per pass, mean of %d passes" % (reads, sample.size / 1e6, nproc, t, len(times)),
                seconds=t, result=int(result), checksum=int(checksum), sample_reads=int(reads))


# --------------------------------------------------------------------------
# the kernels' own algorithmic bytes (what each one's interface makes it move once)
# --------------------------------------------------------------------------
def own_bytes(n_in: int, n_lines: int, n_recs: int, path: int, count_only: bool) -> dict:
    fused = bool(path & 2)
    if fused:
        # k12_scan_pack: text in, three bit-planes (3 bits per byte) and the line starts out
        tok = {"k12_scan_pack": n_in + 3 * n_in // 8 + 4 * n_lines}
        matcher = 3 * n_in // 8
    else:
        # K1: text in, class nibbles and line starts out; pack: nibbles in, planes {p0,p1,p2,-} out
        tok = {"k1_scan_classify": n_in + n_in // 2 + 4 * n_lines, "k15_pack": n_in // 2 + n_in // 2}
        matcher = n_in // 2
    out = dict(tok)
    out["k2_matcher"] = matcher + (0 if count_only else 8 * n_recs)
    # K3/K4: candidates in (8 B per line), one 32-byte sector of text per record, records out
    out["k34_finish"] = 0 if count_only else 8 * n_lines + 32 * n_recs + 16 * n_recs
    return out


# --------------------------------------------------------------------------
# one workload on this rank's GPU (all ranks run it together)
# --------------------------------------------------------------------------
class Ctx:
    pass


def device_run(cx, wname: str, reads_override: int, steps: int, warmup: int, want_e2e: bool, sample_clocks: bool):
    B, L, torch, dist = cx.B, cx.L, cx.torch, cx.dist
    w = dict(WORKLOADS[wname])
    reads = int(reads_override) if reads_override else w["reads"]
    g = B.make_gen(**w["gen"])
    rec_bytes = L.sqbGenBytes(C.byref(g), 0, 1)
    w["rec_bytes"] = rec_bytes
    nbytes = rec_bytes * reads
    sq = B.Seeq(w["pattern"], w["tau"])
    eng = B.Engine.borrowed(sq.engine())
    stream = cx.stream

    d_text = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    first_read = cx.rank * reads
    assert L.sqbGenDevice(C.byref(g), first_read, reads, d_text.data_ptr(), stream.cuda_stream) == 0, B.last_error()
    torch.cuda.synchronize()

    opt = (w["options"]) | (B.SQB_COUNT_ONLY if w["count"] else 0)
    # a shard of 2 GiB or more (cfg4: 2.51 GB; cfg5 at 12.5 GB per GPU: --reads 39800000) is scanned in
    # newline-aligned chunks by sqbScanDeviceLarge
    big = nbytes >= (1 << 31)
    big_stats = {}

    # One step = one scan.  Two scans may be in flight (sqbScanDeviceIssue / Wait, one slot of result
    # arrays each): step i+1 is queued on the same stream before the host waits for step i, so the
    # device never idles between steps while the host reads the counters back.
    def issue(i, timing=False):
        o = opt | (B.SQB_TIMING if timing else 0)
        if big:      # like the one-batch scans of this arm, the records stay in HBM
            big_stats[i] = eng.scan_device_large(d_text.data_ptr(), nbytes, o | B.SQB_DEVICE_RESULTS, stream.cuda_stream)
        else:
            eng.scan_device_issue(i & 1, d_text.data_ptr(), nbytes, o, stream.cuda_stream)

    def wait(i):
        return big_stats.pop(i) if big else eng.scan_device_wait(i & 1)

    def barrier():
        if cx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        issue(i, timing=(i & 1) == 0)         # as in the timed region: the scans of slot 0 carry the events
        st = wait(i)
    clocks = Clocks(cx.local_rank) if (sample_clocks and cx.rank == 0) else None
    barrier()
    if clocks:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = {k: [] for k in ("k1", "k2", "fin", "match", "pack", "k1c")}
    launches = reruns = 0

    def account(st, timed):
        nonlocal launches, reruns
        if timed:
            for k, idx in (("k1", 0), ("k2", 1), ("fin", 2), ("match", 3), ("pack", 4), ("k1c", 5)):
                ms[k].append(st.kernel_ms[idx])
        launches += st.launches
        reruns += st.reruns

    # Per-kernel times come from INSIDE the timed region: the steps of slot 0 (every other step) carry
    # SQB_TIMING -- the engine's CUDA events around its single kernels, as nodes of the replayed graph.
    # Nine event records cost a step ~37 us (r1v), so the steps of slot 1 run bare.
    def timed(i):
        return (i & 1) == 0

    barrier()
    ev0.record(stream)
    issue(0, timing=timed(0))
    for i in range(1, steps):
        issue(i, timing=timed(i))
        account(wait(i - 1), timed(i - 1))
    st = wait(steps - 1)
    account(st, timed(steps - 1))
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if clocks else None
    nlines, nmatched, nrecs, path = st.nlines, st.nmatched, st.nrecs, st.path

    # the tiny exchanges: global line base, totals; time = max over ranks
    from seeq_b200 import shard
    line_base, tot_lines, tot_matched, tot_recs = shard.exchange(nlines, nmatched, nrecs, dist if cx.world > 1 else None,
                                                                  device="cuda")
    if cx.world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / steps
    total_bytes = nbytes * cx.world

    r = dict(w=w, reads=reads, nbytes=nbytes, rec_bytes=rec_bytes, gen=g, big=big,
             value=total_bytes / (ms_per_step * 1e-3) / 1e9, reads_per_s=reads * cx.world / (ms_per_step * 1e-3),
             ms_per_step=ms_per_step, nlines=int(nlines), nmatched=int(nmatched), nrecs=int(nrecs), path=int(path),
             tot_lines=int(tot_lines), tot_recs=int(tot_recs), launches=int(launches), reruns=int(reruns), clocks=clk,
             ms={k: (float(np.mean(v)) if v else 0.0) for k, v in ms.items()}, e2e=None, eng=eng, sq=sq, opt=opt)

    # ---------------- end to end through the C-ABI with host buffers ---------
    host_ok = True
    if big:
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except (OSError, IndexError, ValueError):
            avail = 0
        host_ok = avail > 3 * nbytes * max(1, cx.world)
    if want_e2e and host_ok:
        h_ptr = L.sqbHostAlloc(nbytes + 64)
        assert h_ptr, B.last_error()
        torch.cuda.synchronize()
        # same bytes as on the device (copied back once, outside any timed region)
        assert L.sqbMemcpyD2H(h_ptr, d_text.data_ptr(), nbytes) == 0, B.last_error()
        for _ in range(2):
            st2 = eng.scan_host_ptr(h_ptr, nbytes, opt)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            st2 = eng.scan_host_ptr(h_ptr, nbytes, opt)
        barrier()
        e2e_s = time.perf_counter() - t0
        # the ceiling of the box: a bare copy of the same pinned buffer, all ranks at once
        raw = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            assert L.sqbMemcpyH2D(d_text.data_ptr(), h_ptr, nbytes) == 0, B.last_error()
            barrier()
            raw.append(time.perf_counter() - t0)
        raw_s = min(raw)
        if cx.world > 1:
            t = torch.tensor([e2e_s, raw_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s, raw_s = float(t[0].item()), float(t[1].item())
        assert (st2.nlines, st2.nmatched, st2.nrecs) == (nlines, nmatched, nrecs), "e2e result differs"
        r["e2e"] = {"value": total_bytes / (e2e_s / steps) / 1e9, "unit": "GB/s",
                    "reads_per_s": reads * cx.world / (e2e_s / steps),
                    "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(16 * st2.nrecs + 64),
                    "h2d_raw_GBps": total_bytes / raw_s / 1e9, "devices_per_process": int(st2.devices),
                    "api": "sqbScanHost (pinned host text, chunked H2D overlapped with kernels, records D2H)"}
        L.sqbHostFree(h_ptr)
    del d_text
    return r


def config_of(wname, w, reads, rec_bytes):
    """what both arms can state identically about the workload"""
    nbytes = rec_bytes * reads
    return {"workload": wname + ": " + w["desc"], "pattern": w["pattern"], "distance": w["tau"],
            "reads_per_gpu": int(reads), "bytes_per_gpu": int(nbytes), "line_len": w["gen"]["line_len"],
            "l2_policy": "input (%.2f GB) larger than L2 (126 MB)" % (nbytes / 1e9),
            "sharding": "newline-aligned byte ranges, one rank per GPU, no data-path collective"}


def load_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_table(r, peak):
    """per kernel: ms (CUDA events inside the timed region), own algorithmic bytes, fraction of the HBM peak"""
    ob = own_bytes(r["nbytes"], r["nlines"], r["nrecs"], r["path"], r["w"]["count"])
    fused = bool(r["path"] & 2)
    times = {"k2_matcher": r["ms"]["match"], "k34_finish": r["ms"]["fin"]}
    if fused:
        times["k12_scan_pack"] = r["ms"]["k1c"]
    else:
        times["k1_scan_classify"] = r["ms"]["k1c"]
        times["k15_pack"] = r["ms"]["pack"]
    out = {}
    for k, t in times.items():
        b = ob.get(k, 0)
        out[k] = {"ms": t, "own_bytes": int(b), "achieved_GBps": (b / (t * 1e-3) / 1e9) if t > 0 else None,
                  "frac_of_peak": (b / (t * 1e-3) / 1e9 / peak) if t > 0 else None}
    return out


# --------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="metric", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (testing only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the extra BASELINE configs (one GPU only)")
    ap.add_argument("--no-bgzf", action="store_true", help="skip the BGZF (bgzip) input line (one GPU only)")
    args = ap.parse_args()
    # six warm-up steps at least: a scan slot replays its step as a CUDA graph once it has been asked for
    # the same scan three times (third time: capture), and there are two slots
    warmup_asked = args.warmup
    args.warmup = max(args.warmup, 6) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = dict(WORKLOADS[args.workload])
    reads = int(args.reads) if args.reads else w["reads"]

    from seeq_b200 import binding as B
    g = B.make_gen(**w["gen"])

    # ---------------- reference arm: rank 0 only, CPU only -------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        L = B.lib()
        w["rec_bytes"] = L.sqbGenBytes(C.byref(g), 0, 1)
        cores = host_cores()
        # a sample big enough for every core, generated once on the host
        sample_reads = min(reads, max(400_000, 600_000 * cores))
        host = B.gen_host(g, sample_reads)
        r = cpu_reference_run(w, host, cores, seconds_target=6.0, steps=args.steps, warmup=args.warmup)
        out = {"impl": "reference", "metric": "reads_scanned_GBps", "value": r["value"], "unit": "GB/s",
               "reads_per_s": r["reads_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u32", "data": "synthetic",
               "config": config_of(args.workload, w, reads, w["rec_bytes"]),
               "cpu_baseline": {"value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"]},
               "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return 0

    # ---------------- B200 arm ----------------------------------------------
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the B200 arm has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    os.environ["SEEQ_B200_DEVICE"] = str(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cx = Ctx()
    cx.B, cx.L, cx.torch, cx.dist = B, B.lib(), torch, dist
    cx.rank, cx.local_rank, cx.world = rank, local_rank, world
    # a stream of our own: the legacy default stream has handle 0, which the C-ABI reads as "use the
    # engine's stream"; both slots must run on ONE stream so that the scans follow each other on the
    # device and the CUDA events bracket them
    cx.stream = torch.cuda.Stream()
    torch.cuda.set_stream(cx.stream)

    r = device_run(cx, args.workload, args.reads, args.steps, args.warmup, not args.no_e2e, True)
    peak, peak_src = load_peak()
    nbytes, nrecs = r["nbytes"], r["nrecs"]

    # ---------------- cpu baseline + parity of the timed output (rank 0) -----
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        w2 = r["w"]
        cores = host_cores()
        sample_reads = min(r["reads"], max(400_000, 600_000 * cores))
        host = B.gen_host(r["gen"], sample_reads)
        c = cpu_reference_run(w2, host, cores, seconds_target=12.0)
        cpu = {"value": c["value"], "unit": "GB/s", "reads_per_s": c["reads_per_s"], "cores": c["cores"],
               "kind": c["kind"], "sample": c["sample"], "one_core_GBps": c["one_core_value"]}
        # the same sample through the CUDA path (rank 0's stream starts at read 0: these are the first
        # reads of the buffer the timed steps scanned) against what the reference returned for it
        sample = host[:c["sample_reads"] * r["rec_bytes"]]
        st3 = r["eng"].scan_host(sample, r["opt"])
        if w2["count"]:
            got, got_ck = int(st3.nmatched), int(st3.nmatched)
        else:
            from oracle import pyoracle
            recs = r["eng"].host_records()
            got = int(recs.size)
            got_ck = pyoracle.records_checksum(recs["line"].astype(np.uint64) + 1, recs["start"], recs["end"], recs["dist"])
        cpu["parity"] = {"checked": "count and checksum of (line, start, end, dist) over every record of the sample",
                         "reference_result": c["result"], "gpu_result": got,
                         "reference_checksum": "%016x" % c["checksum"], "gpu_checksum": "%016x" % got_ck,
                         "ok": bool(got == c["result"] and got_ck == c["checksum"])}
        assert cpu["parity"]["ok"], "CUDA output differs from the reference on the timed sample: %r" % (cpu["parity"],)

    # ---------------- the other BASELINE configs (one GPU) --------------------
    configs = None
    if world == 1 and not args.no_configs and not args.reads:
        configs = {}
        traffic_all = {}
        tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
        if os.path.exists(tp):
            try:
                traffic_all = json.load(open(tp))
            except (ValueError, OSError):
                traffic_all = {}
        for name in BASELINE_CONFIGS:
            if name == args.workload:
                continue
            try:
                x = device_run(cx, name, 0, 10, 6, not args.no_e2e, False)
            except Exception as exc:       # a config that fails is reported, the line still prints
                configs[name] = {"error": repr(exc)}
                continue
            b = x["nbytes"] + 16 * x["nrecs"] + 8
            kt = kernel_table(x, peak)
            pipes = (traffic_all.get("_pipes", {}) or {}).get(name) or {}
            configs[name] = {"desc": x["w"]["desc"], "GBps": x["value"], "reads_per_s": x["reads_per_s"],
                             "ms_per_step": x["ms_per_step"], "bytes": int(x["nbytes"]), "records": x["nrecs"],
                             "step_frac_of_peak": b / (x["ms_per_step"] * 1e-3) / 1e9 / peak,
                             "e2e_GBps": x["e2e"]["value"] if x["e2e"] else None,
                             "kernels_ms": {k: v["ms"] for k, v in kt.items()},
                             "kernels_frac_of_peak": {k: v["frac_of_peak"] for k, v in kt.items()},
                             "alu_pipe_pct_ncu": {k: v.get("alu_pipe_pct") for k, v in pipes.items()} or None,
                             "path": x["path"], "reruns": x["reruns"]}
            x["sq"].close()

    # ---------------- the same scan fed with a BGZF (bgzip) buffer (one GPU) ----
    # the compressed bytes cross the link, k0_inflate_bgzf_pair inflates them in HBM, the scan runs
    # there.  1 GiB of the workload's reads, deflated by zlib at level 6; GB/s are of TEXT; text and counts are checked.
    bgzf = None
    if world == 1 and not args.no_bgzf and not args.no_e2e and not args.reads and not w["count"]:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("bgzf_bench", os.path.join(ROOT, "tools", "bgzf_bench.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            z = mod.measure(workload=args.workload, sweep=False)
            bgzf = {"e2e": {"value": z["e2e_bgzf_GBps_of_text"], "unit": "GB/s of text",
                            "h2d_bytes_per_step": z["bgzf_bytes"], "d2h_bytes_per_step": 16 * z["nrecs"] + 64,
                            "api": z["api"]},
                    "e2e_plain_text_GBps": z["e2e_plain_GBps"], "text_bytes": z["text_bytes"], "bgzf_bytes": z["bgzf_bytes"],
                    "ratio": z["ratio"], "members": z["members"], "zlib_level": z["zlib_level"],
                    "inflate_kernel_GBps_of_text": z["inflate_kernel_GBps_of_text"],
                    "inflate_kernel_ms": min(z["inflate_kernel_ms"]),
                    "link_bound_GBps_of_text": z["link_bound_GBps_of_text"],
                    "parity": {"checked": "inflated text byte for byte against the input of the deflater; line, match and "
                                          "record counts against sqbScanHost of the plain text",
                               "ok": bool(z["same_counts"] and z["same_text"])}}
        except Exception as exc:           # reported, the line still prints
            bgzf = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline ------------------------------------------------
    b_alg = nbytes + 16 * nrecs + 8                     # N_in + 16 N_rec + 8
    kt = kernel_table(r, peak)
    dominant = max(kt, key=lambda k: kt[k]["ms"])
    step_achieved = b_alg / (r["ms_per_step"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "step", "achieved": step_achieved, "peak": peak, "unit": "GB/s",
                "frac": step_achieved / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
                "algorithmic_bytes": int(b_alg), "step_ms": r["ms_per_step"],
                "dominant_kernel": dict(name=dominant, **kt[dominant]),
                "kernels": kt,
                "step_breakdown_ms": {"k1_line_scan": r["ms"]["k1"], "k2_forward": r["ms"]["k2"], "k34_finish": r["ms"]["fin"]},
                "step_frac_of_peak": step_achieved / peak}
    traffic_path = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(traffic_path):
        try:
            tj = json.load(open(traffic_path))
            per = tj.get(args.workload) or {}
            if per:
                roofline["traffic"] = int(sum(v for v in per.values() if isinstance(v, (int, float))))
                roofline["traffic_per_kernel"] = per
                roofline["traffic_source"] = ("static: %s (ncu --set full capture of one step of this workload, "
                                              "committed; NOT measured in this run)" % (tj.get("_source", {}).get(args.workload)))
            # SURVEY 8(d), secondary: integer-pipe utilisation of the kernels from the same ncu capture
            roofline["ncu_pipes"] = tj.get("_pipes", {}).get(args.workload)
        except (ValueError, OSError):
            pass

    w = r["w"]
    out = {"metric": "reads_scanned_GBps", "value": r["value"], "unit": "GB/s", "reads_per_s": r["reads_per_s"],
           "n_gpus": world, "steps": args.steps, "warmup": warmup_asked, "ms_per_step": r["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
           "config": config_of(args.workload, w, r["reads"], r["rec_bytes"]),
           "run": {"pattern_length": len(r["sq"].keys), "lines_per_gpu": r["nlines"], "records_per_gpu": r["nrecs"],
                   "matched_lines_per_gpu": r["nmatched"], "total_lines": r["tot_lines"], "total_records": r["tot_recs"],
                   "warmup_done": args.warmup,
                   "scan": ("sqbScanDeviceLarge: newline-aligned chunks of <= 1536 MiB, records kept in HBM"
                            if r["big"] else "sqbScanDeviceIssue/Wait: one batch, two scans in flight, graph replay"),
                   "kernel_path": ("fused tokenise+pack" if r["path"] & 2 else "K1 + pack") +
                                  (", bit-sliced matcher" if r["path"] & 1 else ", word-parallel matcher")},
           "roofline": roofline, "cpu_baseline": cpu, "e2e": r["e2e"], "gpu_launches": r["launches"],
           "scan_reruns": r["reruns"], "clocks": r["clocks"], "configs": configs, "bgzf": bgzf}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
