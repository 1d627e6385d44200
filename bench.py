#!/usr/bin/env python3
"""Headline benchmark: genome Gbp/s scanned + scored (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload arabidopsis|sorghum|maize|sample]

A "step" is one pass of the hot path (PAM scan both strands + ordered
compaction + Rule-Set-1 scoring of every candidate) over one synthetic genome.
  value  whole-job Gbp/s with the packed genome already resident in HBM,
         timed with CUDA events on the library's stream (max over ranks)
  e2e    the same metric through the C ABI from HOST buffers: pinned ASCII
         tokens -> H2D -> pack -> scan+score -> D2H of every candidate record
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement".
Under torchrun (N > 1) each rank owns one GPU and one contiguous shard of the
genome; the only collective is the NCCL all-gather of per-segment counts.
"""
import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
faulthandler.enable()

WORKLOADS = {
    # name: (seed, chromosome lengths in bp, GC, lower-case fraction)   SURVEY.md 8d
    "sample": (1, [230218], 0.38, 0.13),
    "arabidopsis": (2, [34_000_000, 22_000_000, 26_000_000, 21_000_000, 32_000_000], 0.36, 0.15),
    "sorghum": (3, [int(x * 1.07e6) for x in (81, 78, 74, 69, 72, 62, 65, 63, 59, 61)], 0.44, 0.60),
    "maize": (4, [int(x * 1.09e6) for x in (307, 244, 235, 247, 223, 174, 182, 181, 159, 150)], 0.47, 0.50),
}
CONFIG_NAME = {"sample": "configs[0] sample-scale synthetic", "arabidopsis": "configs[1] synthetic Arabidopsis-scale 135 Mbp x5 chr",
               "sorghum": "configs[2] synthetic Sorghum-scale 730 Mbp x10 chr", "maize": "configs[3] synthetic maize-scale 2.3 Gbp x10 chr"}
METRIC = "genome Gbp/s scanned+scored"


def workload_lengths(name, copies=1):
    return WORKLOADS[name][1] * copies


def synth_tokens(name, only=None, copies=1):
    """Synthetic genome as the reference's *formatted-path* tokens: i.i.d. bases
    at the stated GC, lower-case blocks of 1-50 kb, wrapped in the quote/paren
    decoration that str(list_of_tuples) leaves.  Built as
    uint8 arrays directly -- the text round trip is the ingest row, not this one.
    `copies` > 1 appends further independent genomes of the same shape (weak
    scaling); `only` = set of token indices to materialise (others are None)."""
    seed, lengths, gc, lower = WORKLOADS[name]
    lengths = lengths * copies
    toks = []
    lut = np.frombuffer(b"ATCG", dtype=np.uint8)
    thr = np.cumsum([(1 - gc) / 2, (1 - gc) / 2, gc / 2, gc / 2])
    for k, n in enumerate(lengths):
        if only is not None and k not in only:
            toks.append(None)
            continue
        r = np.random.default_rng(seed * 1000 + k)      # independent stream per chromosome
        u = r.random(n, dtype=np.float32)
        s = lut[np.searchsorted(thr, u, side="right").clip(0, 3)]
        i = 0
        while i < n:
            blk = int(r.integers(1000, 50000))
            if r.random() < lower:
                s[i:i + blk] |= 0x20
            i += blk
        tail = b"')," if k + 1 < len(lengths) else b"')]"
        toks.append(np.concatenate((np.frombuffer(b"'", np.uint8), s, np.frombuffer(tail, np.uint8))))
    return toks


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) == 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) == 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) == 6 for n, v in zip(names, r[2:]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_scan_score launch on this
    workload, from the committed `ncu --set full` capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[workload]["dram_bytes_per_launch"]
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_port_sample(workload, seconds_target=15.0):
    """Time the oracle's literal port of the reference (Python + np.matmul +
    csv) on a bounded prefix of the workload's first chromosome."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cropsr_oracle as oracle
    n = int(min(WORKLOADS[workload][1][0], seconds_target * 0.8e6))      # the port runs at ~0.8 Mbp/s on this workload
    tok = synth_tokens(workload, only={0})[0][1:1 + n].tobytes().decode("ascii")
    text = ">chr1\n" + tok            # clean two-line path: the token is the sequence itself
    np.random.seed(0)
    dt, n_bases, rows = oracle.timed_reference_pass(text, 20)
    threads = int(os.environ.get("OPENBLAS_NUM_THREADS", os.cpu_count() or 1))
    return {"value": n_bases / dt / 1e9, "unit": "Gbp/s", "cores": threads, "kind": "port",
            "sample": f"first {n} bp of chromosome 1 of the {workload} workload, one full "
                      f"scan+score+CSV-rows pass of oracle/cropsr_oracle.py ({rows} rows, {dt:.1f} s); "
                      "single-threaded Python, OpenBLAS threads only inside np.matmul"}, dt, n_bases


def run_reference(args, rank):
    if rank != 0:
        return
    times, nb = [], 0
    for i in range(args.warmup + args.steps):
        cb, dt, nb = cpu_port_sample(args.workload, min(10.0, 150.0 / (args.warmup + args.steps)))
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    v = nb / t / 1e9
    cb["value"] = v
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Gbp/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": CONFIG_NAME[args.workload], "guide_len": 20},
                      "cpu_baseline": cb,
                      "e2e": {"value": v, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="arabidopsis", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from cropsr_b200 import engine, shard

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cpus_before = os.sched_getaffinity(0)
    numa_node = engine.bind_host_near(local)      # before any pinned allocation: staging buffers next to the GPU
    engine.init(local)

    # ---- workload and this rank's shard.  Weak scaling: N GPUs scan N genomes of the
    # configured shape laid end to end and cut into N contiguous, tile-aligned shards.
    lengths = [n + 4 for n in workload_lengths(args.workload, world)]     # + quote/paren decoration
    n_bases_total = sum(lengths)
    plans = shard.plan(lengths, world)
    mine = plans[rank]
    toks = synth_tokens(args.workload, only={k for k, _, _ in mine}, copies=world)

    # pinned host staging of this rank's token bytes (what a host ingest would hand over)
    host_tokens = [None] * len(toks)
    pinned = []
    for k, t in enumerate(toks):
        if t is not None:
            pinned.append(engine.PinnedBuffer(len(t)))
            pinned[-1].array[:] = t
            host_tokens[k] = pinned[-1].array

    def build():
        g = engine.Genome()
        for k, a, b in mine:
            g.add_segment(k, host_tokens[k], a, b)
        return g.commit()

    genome = build()
    my_bases = sum(b - a for _, a, b in mine)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_gather_counts(res):
        """NCCL all-gather of the per-segment counts straight from the library's
        device buffer; returns the global row offsets of this rank's segments."""
        n_slots = len(lengths) + 1
        buf = torch.zeros(2 * n_slots, dtype=torch.int64, device="cuda")
        ns = len(mine)
        if ns:
            class _Raw:
                __cuda_array_interface__ = {"shape": (2 * ns,), "typestr": "<i8", "data": (res.device_counts_ptr(), False),
                                            "version": 2}
            raw = torch.as_tensor(_Raw(), device="cuda")
            buf[:ns] = raw[:ns]
            buf[n_slots:n_slots + ns] = raw[ns:]
        if world > 1:
            out = torch.empty(world * 2 * n_slots, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(out, buf)
        else:
            out = buf
        h = out.cpu().numpy().reshape(world, 2, n_slots)
        counts = [(h[r, 0, :len(plans[r])], h[r, 1, :len(plans[r])]) for r in range(world)]
        return shard.global_offsets(plans, counts)

    # ---- device-resident metric
    launches0 = None
    scan_ms, n_cand = [], 0
    sampler = None
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            barrier()
            launches0 = engine.launch_count()
            sampler = ClockSampler(local)
            sampler.start()
            t_wall0 = time.perf_counter()
        flush.zero_()
        torch.cuda.synchronize()
        res = genome.scan(20)
        offsets, total_rows = all_gather_counts(res)
        if i >= args.warmup:
            scan_ms.append(res.scan_ms())
        n_cand = res.n_plus + res.n_minus
        res.free()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3 / args.steps
    launches = engine.launch_count() - launches0
    ms = float(np.mean(scan_ms))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_bases_total / (ms * 1e-3) / 1e9

    # ---- end to end from host buffers (pinned in, pinned out), through the pipelined C-ABI call:
    # per segment H2D -> pack -> scan+score -> D2H of every candidate row, overlapped across segments
    def all_gather_host_counts(n_plus, n_minus):
        n_slots = len(lengths) + 1
        buf = torch.zeros(2 * n_slots, dtype=torch.int64)
        buf[:len(n_plus)] = torch.from_numpy(n_plus.astype(np.int64))
        buf[n_slots:n_slots + len(n_minus)] = torch.from_numpy(n_minus.astype(np.int64))
        buf = buf.cuda()
        if world > 1:
            out = torch.empty(world * 2 * n_slots, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(out, buf)
        else:
            out = buf
        h = out.cpu().numpy().reshape(world, 2, n_slots)
        counts = [(h[r, 0, :len(plans[r])], h[r, 1, :len(plans[r])]) for r in range(world)]
        return shard.global_offsets(plans, counts)

    e2e_ms, h2d, d2h = [], 0, 0
    arena = None
    segs = [(k, host_tokens[k], a, b) for k, a, b in mine]
    for i in range(args.warmup + max(3, args.steps // 2)):
        barrier()
        t0 = time.perf_counter()
        arena, n_plus, n_minus, _ = engine.scan_segments(segs, 20, arena=arena)
        all_gather_host_counts(n_plus, n_minus)
        barrier()
        dt = (time.perf_counter() - t0) * 1e3
        if i >= args.warmup:
            e2e_ms.append(dt)
        h2d = sum(min(b + 32, lengths[k]) - max(a - 32, 0) for k, a, b in mine)
        d2h = 20 * int(n_plus.sum() + n_minus.sum())
    clocks = sampler.summary()       # sampled from the first timed scan to the last end-to-end step
    e2e = float(np.mean(e2e_ms))
    g = build()                     # a warm commit (the first one of a process pays lazy module loading)
    ingest_timing = g.timing()
    g.free()
    if world > 1:
        t = torch.tensor([e2e, float(h2d), float(d2h), float(n_cand)], dtype=torch.float64, device="cuda")
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e, h2d, d2h, n_cand_total = float(mx[0]), int(t[1]), int(t[2]), int(t[3])
    else:
        n_cand_total = n_cand

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = 0.5 * my_bases + 20.0 * n_cand          # SURVEY 8d, this rank's launch
        achieved = alg_bytes / (float(np.mean(scan_ms)) * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": CONFIG_NAME[args.workload] + (f" x{world} (one per GPU)" if world > 1 else ""), "bases": n_bases_total, "candidates": n_cand_total,
                       "guide_len": 20, "sharding": f"{world} contiguous shard(s), tile-aligned, halo 32/32",
                       "l2": "flushed between steps (512 MiB memset)", "host_numa_node": numa_node},
            "e2e": {"value": n_bases_total / (e2e * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "wall_ms_per_step_incl_flush_alloc_allgather": wall_ms,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload) if world == 1 else None, "peak_source": peak_src,
                         "algorithmic_bytes": "0.5 B/base read + 20 B/candidate written (pos u32, packed u64, x f64)"},
            "ingest": ingest_timing,
        }
        if not args.no_cpu_baseline and world == 1:
            os.sched_setaffinity(0, cpus_before)          # the CPU baseline may use every core of the host
            line["cpu_baseline"] = cpu_port_sample(args.workload)[0]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
