#!/usr/bin/env python
"""bench.py -- throughput of the seeq matching path on B200 (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W [--workload cfg2] [--reads R] [--impl reference]

A *step* is one pass of the hot path (K1 line scan -> bit-plane pack -> K2 forward matcher ->
K3 reverse pass / K4 ordered compaction) over one batch of synthetic reads.
Default workload = BASELINE.json configs[1] (cfg2): `seeq -b -l -p -k -d 1
A[CG]TNNGATC` over 10 M synthetic 150-nt reads (1.51 GB per GPU; weak scaling:
every rank scans its own 10 M reads of one global read stream).  A shard of 2 GiB
or more (`--workload cfg5 --reads 39800000`: one GPU's share of BASELINE config 5)
goes through sqbScanDeviceLarge.

One JSON line is printed by rank 0:
  value        GB/s of reads scanned, whole job, input resident in HBM: CUDA events
               around K steps on the stream the scans are queued on, max over ranks;
               two scans in flight, a repeated scan replays as a CUDA graph
  e2e          same metric through the C-ABI with HOST buffers (sqbScanHost:
               pinned host text -> H2D -> kernels -> records D2H inside the timed
               region)
  roofline     the slowest single kernel of the step against the measured HBM copy
               peak: algorithmic bytes / its duration, from CUDA events the engine
               records around its kernels INSIDE the timed region (every other step);
               per-kernel times, DRAM traffic and pipe utilisation from the committed
               ncu capture next to it
  cpu_baseline the unmodified reference (oracle/_ref) on the host cores, on a
               bounded sample of the same workload
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


# BASELINE.json configs; `count` selects the count-only path (CLI -c)
WORKLOADS = {
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


def reads_of(w, override):
    return int(override) if override else w["reads"]


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
    on a bounded sample of host_text.  -> dict(value GB/s, reads/s, kind, cores, sample)."""
    from oracle import pyoracle
    pyoracle.build()
    rec_bytes = w["rec_bytes"]
    mode = 0 if w["count"] else 2
    if pyoracle.have_reference():
        ref = pyoracle.Reference()
        kind = "reference"

        def run(buf, n):
            return ref.bench(buf, w["pattern"], w["tau"], w["options"], mode, n)
    else:
        orc = pyoracle.Oracle()
        keys, _ = orc.parse(w["pattern"])
        kind = "port"
        nproc = 1

        def run(buf, n):
            t0 = time.perf_counter()
            r, nl, nm = orc.buffer_scan(buf, keys, w["tau"], w["options"])
            return time.perf_counter() - t0, (nm if w["count"] else len(r))
    # calibrate on ~8 MB, one core
    cal_reads = max(1, min(host_text.size // rec_bytes, (8 << 20) // rec_bytes))
    t_cal, _ = run(host_text[:cal_reads * rec_bytes], 1)
    rate1 = cal_reads * rec_bytes / max(t_cal, 1e-6)                 # bytes/s on one core
    want = int(rate1 * nproc * seconds_target * 0.7)                 # imperfect scaling margin
    reads = max(nproc, min(host_text.size // rec_bytes, want // rec_bytes))
    sample = host_text[:reads * rec_bytes]
    for _ in range(warmup):
        run(sample, nproc)
    times = []
    result = None
    for _ in range(max(1, steps)):
        t, result = run(sample, nproc)
        times.append(t)
    # a sample the host cores finish in a fraction of a second is repeated (about 2 s of wall time)
    while steps <= 1 and sum(times) < 2.0 and len(times) < 8:
        t, result = run(sample, nproc)
        times.append(t)
    t = float(np.mean(times))
    return dict(value=sample.size / t / 1e9, reads_per_s=reads / t, unit="GB/s", cores=nproc, kind=kind,
                one_core_value=rate1 / 1e9,
                sample="%d reads (%.1f MB) of the same workload, %d processes over newline-aligned shards, "
                       "%.2f s per pass, mean of %d passes" % (reads, sample.size / 1e6, nproc, t, len(times)),
                seconds=t, result=int(result))


# --------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (testing only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    # six warm-up steps at least: a scan slot replays its step as a CUDA graph once it has been asked for
    # the same scan three times (third time: capture), and there are two slots
    args.warmup = max(args.warmup, 6) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = dict(WORKLOADS[args.workload])
    reads = reads_of(w, args.reads)

    from seeq_b200 import binding as B
    g = B.make_gen(**w["gen"])
    L = None

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
               "config": {"workload": args.workload + ": " + w["desc"], "reads_per_gpu": reads,
                          "line_len": w["gen"]["line_len"]},
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
    L = B.lib()
    rec_bytes = L.sqbGenBytes(C.byref(g), 0, 1)
    w["rec_bytes"] = rec_bytes
    nbytes = rec_bytes * reads

    sq = B.Seeq(w["pattern"], w["tau"])
    eng = B.Engine.borrowed(sq.engine())

    # this rank's slice of the global read stream, generated on the device
    # a stream of our own: the legacy default stream has handle 0, which the C-ABI reads as
    # "use the engine's stream"; both slots must run on ONE stream so that the scans follow
    # each other on the device and the CUDA events bracket them
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    d_text = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    first_read = rank * reads
    assert L.sqbGenDevice(C.byref(g), first_read, reads, d_text.data_ptr(), stream.cuda_stream) == 0, B.last_error()
    torch.cuda.synchronize()

    opt = w["options"] | (B.SQB_COUNT_ONLY if w["count"] else 0)
    # a shard of 2 GiB or more (cfg5 at 12.5 GB per GPU: --reads 39800000) is scanned in newline-aligned
    # chunks by sqbScanDeviceLarge; its records travel to the host while the next chunk is matched
    big = nbytes >= (1 << 31)

    # One step = one scan.  Two scans may be in flight (sqbScanDeviceIssue / Wait, one slot of
    # result arrays each): step i+1 is queued on the same stream before the host waits for
    # step i, so the device never idles between steps while the host reads the counters back.
    big_stats = {}

    def issue(i, timing=False):
        o = opt | (B.SQB_TIMING if timing else 0)
        if big:      # like the one-batch scans of this arm, the records stay in HBM
            big_stats[i] = eng.scan_device_large(d_text.data_ptr(), nbytes, o | B.SQB_DEVICE_RESULTS, stream.cuda_stream)
        else:
            eng.scan_device_issue(i & 1, d_text.data_ptr(), nbytes, o, stream.cuda_stream)

    def wait(i):
        return big_stats.pop(i) if big else eng.scan_device_wait(i & 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        issue(i, timing=(i & 1) == 0)         # as in the timed region: the scans of slot 0 carry the events
        st = wait(i)
    clocks = Clocks(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k2_ms, k1_ms, fin_ms, launches, reruns = [], [], [], 0, 0
    match_ms, pack_ms, k1c_ms = [], [], []

    def account(st):
        nonlocal launches
        k1_ms.append(st.kernel_ms[0])
        k2_ms.append(st.kernel_ms[1])
        fin_ms.append(st.kernel_ms[2])
        match_ms.append(st.kernel_ms[3])
        pack_ms.append(st.kernel_ms[4])
        k1c_ms.append(st.kernel_ms[5])
        launches += st.launches

    # Per-kernel times for the roofline come from INSIDE the timed region: the steps of slot 0 (every
    # other step) carry SQB_TIMING -- the engine's CUDA events around its single kernels, as nodes of
    # the replayed graph.  Nine event records cost a step ~37 us (r1v: cfg2 1.396 -> 1.434 ms with all
    # steps instrumented), so the steps of slot 1 run bare.
    def timed(i):
        return (i & 1) == 0

    barrier()
    ev0.record(stream)
    issue(0, timing=timed(0))
    for i in range(1, args.steps):
        issue(i, timing=timed(i))
        st = wait(i - 1)
        if timed(i - 1):
            account(st)
        else:
            launches += st.launches
        reruns += st.reruns
    st = wait(args.steps - 1)
    if timed(args.steps - 1):
        account(st)
    else:
        launches += st.launches
    reruns += st.reruns
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    nlines, nmatched, nrecs = st.nlines, st.nmatched, st.nrecs

    # the tiny exchanges: global line base, totals; time = max over ranks
    from seeq_b200 import shard
    line_base, tot_lines, tot_matched, tot_recs = shard.exchange(nlines, nmatched, nrecs, dist if world > 1 else None,
                                                                  device="cuda")
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    total_bytes = nbytes * world
    value = total_bytes / (ms_per_step * 1e-3) / 1e9
    reads_per_s = reads * world / (ms_per_step * 1e-3)

    # ---------------- end to end through the C-ABI with host buffers ---------
    e2e = None
    host_ok = True
    if big:
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except (OSError, IndexError, ValueError):
            avail = 0
        host_ok = avail > 3 * nbytes * max(1, world)
    if not args.no_e2e and host_ok:
        h_ptr = L.sqbHostAlloc(nbytes + 64)
        assert h_ptr, B.last_error()
        torch.cuda.synchronize()
        # same bytes as on the device (copied back once, outside any timed region)
        assert L.sqbMemcpyD2H(h_ptr, d_text.data_ptr(), nbytes) == 0, B.last_error()
        e2e_opt = opt
        for _ in range(2):
            st2 = eng.scan_host_ptr(h_ptr, nbytes, e2e_opt)
        barrier()
        t0 = time.perf_counter()
        e2e_launches = 0
        for _ in range(args.steps):
            st2 = eng.scan_host_ptr(h_ptr, nbytes, e2e_opt)
            e2e_launches += st2.launches
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        assert (st2.nlines, st2.nmatched, st2.nrecs) == (nlines, nmatched, nrecs), "e2e result differs"
        e2e = {"value": total_bytes / (e2e_s / args.steps) / 1e9, "unit": "GB/s",
               "reads_per_s": reads * world / (e2e_s / args.steps),
               "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(16 * st2.nrecs + 64),
               "api": "sqbScanHost (pinned host text, chunked H2D overlapped with kernels, records D2H)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the slowest kernel of the step ------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    b_alg = nbytes + 16 * nrecs + 8                     # N_in + 16 N_rec + 8
    # CUDA events recorded by the engine on its own stream around the single kernels
    # (SQB_TIMING): the matcher, the bit-plane pack, the tokenizer; and around the stages
    kern = {"k2_matcher": float(np.mean(match_ms)), "k15_pack": float(np.mean(pack_ms)),
            "k1_scan_classify": float(np.mean(k1c_ms)), "k34_finish": float(np.mean(fin_ms))}
    dominant = max(kern, key=kern.get)
    kd = kern[dominant]
    achieved = b_alg / (kd * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes": int(b_alg), "kernel_ms": kd,
                "kernels_ms": kern,
                "kernels_frac_of_peak": {k: (b_alg / (v * 1e-3) / 1e9 / peak if v > 0 else None) for k, v in kern.items()},
                "step_breakdown_ms": {"k1_line_scan": float(np.mean(k1_ms)), "k2_forward": float(np.mean(k2_ms)),
                                      "k34_finish": float(np.mean(fin_ms))},
                "step_frac_of_peak": b_alg / (ms_per_step * 1e-3) / 1e9 / peak}
    traffic_path = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(traffic_path):
        try:
            tj = json.load(open(traffic_path))
            roofline["traffic"] = tj.get(args.workload, {}).get(dominant)
            # SURVEY 8(d), secondary: integer-pipe utilisation of the kernels from the same ncu capture
            roofline["ncu_pipes"] = tj.get("_pipes", {}).get(args.workload)
        except (ValueError, OSError):
            pass

    cpu = None
    if not args.no_cpu_baseline:
        cores = host_cores()
        sample_reads = min(reads, max(400_000, 600_000 * cores))
        host = B.gen_host(g, sample_reads)
        r = cpu_reference_run(w, host, cores, seconds_target=12.0)
        cpu = {"value": r["value"], "unit": "GB/s", "reads_per_s": r["reads_per_s"], "cores": r["cores"],
               "kind": r["kind"], "sample": r["sample"], "one_core_GBps": r["one_core_value"]}

    out = {"metric": "reads_scanned_GBps", "value": value, "unit": "GB/s", "reads_per_s": reads_per_s,
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
           "config": {"workload": args.workload + ": " + w["desc"], "reads_per_gpu": reads,
                      "bytes_per_gpu": int(nbytes), "line_len": w["gen"]["line_len"],
                      "lines_per_gpu": int(nlines), "records_per_gpu": int(nrecs),
                      "matched_lines_per_gpu": int(nmatched), "total_lines": int(tot_lines),
                      "total_records": int(tot_recs), "l2_policy": "input (%.2f GB) larger than L2 (126 MB)" % (nbytes / 1e9),
                      "scan": ("sqbScanDeviceLarge: newline-aligned chunks of <= 1536 MiB, records kept in HBM"
                               if big else "sqbScanDeviceIssue/Wait: one batch, two scans in flight, graph replay"),
                      "sharding": "newline-aligned byte ranges, one rank per GPU, no data-path collective"},
           "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
           "scan_reruns": int(reruns), "clocks": clk}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
