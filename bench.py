#!/usr/bin/env python
"""bench.py -- GCUPS of the DP-relaxation hot path on the BASELINE.json workload.

Workload (BASELINE.json configs[1] / configs[4]): ecoli x sboydii, semi-global,
affine (Gotoh) gaps (same=2, diff=-1, gapInit=-2, gapExtend=-1), score-only,
~4.64 Mbp x 4.6 Mbp.  The bundled genomes are missing from the reference mount,
so seeded synthetic stand-ins are used (anyseq_b200/workloads.py) unless
sequences/ecoli.fna and sequences/sboydii.fna exist.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" = one complete score-only alignment
of the pair.  N > 1 (torchrun): the pair is split into N column strips, one per
GPU, chained by the strip-boundary column over NVLink (strong scaling).

  value     whole-job GCUPS, sequences resident in HBM, device time (CUDA events
            recorded by the library on the stream it launches on), max over ranks
  e2e       same metric through the C ABI with HOST buffers (H2D + D2H inside)
  roofline  integer/DPX issue roofline measured live on this GPU
  cpu_baseline  the oracle's restated reference CPU path on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCORING = dict(same=2, diff=-1, gap_init=-2, gap_extend=-1)
MODE = "semiglobal"
OPS_PER_CELL_AFFINE = 7        # algorithmic 32-bit integer ops per Gotoh cell
METRIC = "GCUPS (score-only ecoli x sboydii affine)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries print to fd 1 behind Python's
# back (NCCL announces its version there), so fd 1 is pointed at stderr for the
# whole run and the JSON line goes to a private duplicate of the original stdout.
_JSON_OUT = None


def _claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "250"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clk, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                clk.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v == "Active":
                    reasons.add(nm)
        if not clk:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(clk)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_median": float(np.median(pw)), "samples": len(clk)}


def cpu_baseline(q, s, sample_rows, sample_cols, threads):
    """restated reference CPU path (oracle port: 1024x1024 block wavefront, scalar
    inner loop, Gotoh as defined by the build) on a bounded sample of the workload"""
    from oracle import oracle as O
    qs, ss = q[:sample_rows], s[:sample_cols]
    t0 = time.perf_counter()
    sc = O.score_affine(MODE, qs, ss, SCORING["same"], SCORING["diff"], SCORING["gap_init"],
                        SCORING["gap_extend"], threads=threads)
    dt = time.perf_counter() - t0
    return len(qs) * len(ss) / dt / 1e9, dt, sc[0]


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host
    cores.  AnyDSL/Impala cannot be built here, so this is the oracle's
    restatement of iteration_cpu/scoring_cpu (kind "port") with all host threads."""
    if rank != 0:
        return 0
    from anyseq_b200 import workloads as W
    q, s, desc = W.whole_genome_pair(args.scale)
    threads = os.cpu_count() or 1
    side = int(args.cpu_sample) or 150000
    times = []
    for i in range(args.warmup + args.steps):
        g, dt, _ = cpu_baseline(q, s, side, side, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = side * side / (ms * 1e-3) / 1e9
    sample = f"first {side} x {side} cells of the workload per step ({side*side:.3g} cells), all {threads} host threads"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": desc + "; semiglobal affine (2,-1,-2,-1) score-only", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (development only)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="side of the CPU-baseline sample (default: 250000 for the cpu_baseline leg, "
                         "150000 per step for --impl reference)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pairs-per-launch", type=int, default=4,
                    help="N > 1, streamed steps: consecutive alignments relaxed side by side in one launch per rank")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="N > 1: barrier between steps (every alignment pays the full wavefront fill)")
    args = ap.parse_args()
    _claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import anyseq_b200 as A
    from anyseq_b200 import workloads as W
    from anyseq_b200.multigpu import StripWavefront, column_slices

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (anyseq_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    q, s, desc = W.whole_genome_pair(args.scale)
    m, n = len(q), len(s)
    cells = float(m) * float(n)
    scoring = A.affine_scoring_scheme(**SCORING)
    al = A.Aligner(local_rank)
    info = al.device_info()

    slices = column_slices(n, world)
    c0, c1 = slices[rank]
    h_q = torch.from_numpy(q).pin_memory()
    h_s = torch.from_numpy(np.ascontiguousarray(s[c0:c1])).pin_memory()
    d_q = h_q.cuda()
    d_s = h_s.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    # N > 1: two inboxes per rank boundary, so consecutive alignments stream through the ranks back to
    # back (rank 0 starts pair k+1 while the wavefront of pair k is still inside the later ranks); the
    # neighbour-to-neighbour run tokens of multigpu.RunTokens replace a barrier between steps
    ppl = 1 if (world == 1 or args.no_pipeline) else max(1, args.pairs_per_launch)
    wave = StripWavefront(al, rank, world, m, dist, depth=1 if args.no_pipeline else 2, pairs_per_launch=ppl)
    torch.cuda.synchronize()

    wave.reset()

    def step_resident(isolated=False):
        flush.zero_()
        if dist is not None and (isolated or args.no_pipeline):
            dist.barrier()          # every rank has finished the previous run
        torch.cuda.synchronize()
        part = wave.run(MODE, scoring, d_q.data_ptr(), m, d_s.data_ptr(), c0, c1, n)
        return part

    def launch_resident(npairs):
        """npairs consecutive alignments of the stream in ONE launch per rank (N > 1: a slice alone cannot fill the GPU)"""
        flush.zero_()
        torch.cuda.synchronize()
        return wave.run_multi(MODE, scoring, [d_q.data_ptr()] * npairs, m, [d_s.data_ptr()] * npairs, c0, c1, n)

    def run_steps(count):
        """-> (last partial, device ms summed over launches, kernel launches)"""
        ms, nl, last, left = 0.0, 0, None, count
        while left > 0:
            k = min(ppl, left)
            parts = launch_resident(k) if k > 1 else [step_resident()]
            ms += parts[0].kernel_ms
            nl += sum(p.kernel_launches for p in parts)
            last = parts[-1]
            left -= k
        return last, ms, nl

    # ---- value: inputs resident in HBM -----------------------------------------
    part, _, _ = run_steps(args.warmup) if args.warmup > 0 else (None, 0, 0)
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    ev0.record()
    part, dev_ms, launches = run_steps(args.steps)
    # wave.run() returns after its kernels have finished (the call synchronises its stream), so an event
    # on the idle torch stream is a device timestamp of "this rank's last step is done"
    ev1.record()
    ev1.synchronize()
    span_ms = ev0.elapsed_time(ev1)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    res = wave.combine(MODE, scoring, part)
    if world == 1:
        span_ms = dev_ms           # one rank: the sum of the per-call CUDA-event times (excludes the L2 flush)
    t = torch.tensor([span_ms, t_wall * 1e3], dtype=torch.float64, device="cuda")
    lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = t.tolist()
    ms_per_step = dev_ms_max / args.steps
    value = cells / (ms_per_step * 1e-3) / 1e9

    # latency of ONE alignment across the ranks (barrier before it, max over ranks of the call's event time)
    single_ms = None
    if world > 1:
        lat = []
        for _ in range(2):
            lat.append(step_resident(isolated=True).kernel_ms)
        tl = torch.tensor([min(lat)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        single_ms = tl.item()

    # ---- e2e: host buffers through the public API --------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 3 if world == 1 else 8))

        def launch_e2e(k):
            """k alignments whose inputs start in pinned HOST memory; returns the score of the last one"""
            if world == 1:
                # the reference-facing C ABI call with HOST pointers: H2D of both
                # sequences, kernels, D2H of the result, all inside the call
                return al.score(MODE, h_q.numpy(), h_s.numpy(), scoring).score
            dqs = [h_q.cuda(non_blocking=True) for _ in range(k)]
            dss = [h_s.cuda(non_blocking=True) for _ in range(k)]
            torch.cuda.synchronize()
            if k > 1:
                parts = wave.run_multi(MODE, scoring, [t_.data_ptr() for t_ in dqs], m, [t_.data_ptr() for t_ in dss], c0, c1, n)
            else:
                parts = [wave.run(MODE, scoring, dqs[0].data_ptr(), m, dss[0].data_ptr(), c0, c1, n)]
            return [wave.combine(MODE, scoring, p_).score for p_ in parts][-1]    # every alignment's result reaches the host

        def run_e2e(count):
            left, sc_last = count, None
            while left > 0:
                k = min(ppl, left)
                sc_last = launch_e2e(k)
                left -= k
            return sc_last

        run_e2e(min(e2e_steps, ppl))   # warm-up
        barrier()
        t0 = time.perf_counter()
        sc_e2e = run_e2e(e2e_steps)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = tt.item()
        e2e = {"value": cells * e2e_steps / dt / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": int(m + (c1 - c0)) if world > 1 else int(m + n),
               "d2h_bytes_per_step": 128, "steps": e2e_steps, "score": int(sc_e2e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline: integer/DPX issue rate measured on this GPU -------------------
    peak_alu, _ = al.measure_int_peak(0)          # VIADDMNMX/VIMNMX3 only, lane-ops/s
    peak_mix7, _ = al.measure_int_peak(2)         # the 7-op all-ALU Gotoh mix, lane-ops/s
    peak_cells_dual, _ = al.measure_int_peak(5)   # the kernel's own 4 ALU + 3 IMAD cell, cells/s
    achieved_ops = (cells / (ms_per_step * 1e-3)) * OPS_PER_CELL_AFFINE / world   # per GPU
    roofline = {
        "bound": "int_alu",
        "achieved": achieved_ops / 1e12, "peak": peak_alu / 1e12, "unit": "Tlaneop/s (int32, per GPU)",
        "frac": achieved_ops / peak_alu,
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE strip-kernel launch at this workload
        # (ncu, profiles/r01_strip_kernel_fullsize_metrics.csv): 232.7 MB + 325.1 MB; only valid for N = 1
        "traffic": 557_788_672 if (world == 1 and args.scale == 1.0) else None,
        "traffic_unit": "bytes of DRAM traffic per launch (algorithmic input: m + n = 9.24 MB; the border records live in L2)",
        "ops_per_cell": OPS_PER_CELL_AFFINE,
        "peak_source": "measured live: anyseq_measure_int_peak(kind=0), dependency-free VIADDMNMX/VIMNMX3 loop",
        "peak_gcups_alu7": peak_alu / OPS_PER_CELL_AFFINE / 1e9,
        "peak_gcups_mix7_measured": peak_mix7 / OPS_PER_CELL_AFFINE / 1e9,
        "peak_gcups_dual_pipe_mix": peak_cells_dual / 1e9,
        "frac_of_dual_pipe_mix": (value / world) / (peak_cells_dual / 1e9),
        "hbm": {"achieved_gbs": (557_788_672 / (ms_per_step * 1e-3) / 1e9) if (world == 1 and args.scale == 1.0) else None,
                "peak_gbs": 6452.8, "note": "DRAM traffic of the strip kernel / step time; far below the HBM roofline by design"},
    }

    cpu = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        side = int(args.cpu_sample) or 250000
        g, dt, _ = cpu_baseline(q, s, side, side, threads)
        # the reference's own team size is 4 threads (src/backend/backend_cpu.impala:13): reported beside it
        side4 = max(1024, side // 2)
        g4, dt4, _ = cpu_baseline(q, s, side4, side4, 4)
        cpu = {"value": g, "unit": "GCUPS", "cores": threads, "kind": "port",
               "sample": f"first {side} x {side} cells of the workload ({dt:.1f} s), restated reference CPU path "
                         f"(1024x1024 block wavefront, scalar inner loop), {threads} threads",
               "reference_team_of_4_threads": {"value": g4, "unit": "GCUPS", "cores": 4,
                                               "sample": f"first {side4} x {side4} cells ({dt4:.1f} s)"}}

    out = {
        "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": desc + "; semiglobal affine (same=2,diff=-1,gapInit=-2,gapExtend=-1) score-only",
                   "rows": m, "cols": n, "cells_per_step": cells,
                   "partition": f"{world} column strip(s), boundary column streamed over NVLink" if world > 1 else "1 GPU",
                   "l2": "L2 flushed (256 MiB memset) before every timed step",
                   "timing": ("CUDA events on the library's launch stream, summed over steps" if world == 1 else
                              "CUDA events around the K steps on every rank, max over ranks"),
                   "steps_overlap": (None if world == 1 else
                                     ("no: barrier between steps" if args.no_pipeline else
                                      f"yes: consecutive alignments stream through the ranks back to back ({ppl} per launch side by "
                                      "side, one inbox each, two inbox sets + neighbour run tokens); single_alignment_ms is one "
                                      "alignment alone, barrier before it")),
                   "pairs_per_launch": ppl,
                   "single_alignment_ms": single_ms,
                   "single_alignment_gcups": (cells / (single_ms * 1e-3) / 1e9) if single_ms else None},
        "wall_ms_per_step": wall_ms_max / args.steps,
        "score": int(res.score),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(lt.item()),
        "roofline": roofline, "cpu_baseline": cpu,
        "device": info["name"], "sm_count": info["sm_count"],
    }
    emit(out)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
