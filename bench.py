#!/usr/bin/env python
"""bench.py -- reads/s and wrap-around-DP GCUPS of the mTR hot path on B200, beside the reference on the host CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl reference] [--quick] [--inflight B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of R (default 8192) synthetic C5 reads (BASELINE.json configs[4]: 10-20 kb reads carrying one tandem repeat,
unit 2-500 bp, 5-15 % sub/ins/del noise) through the whole per-read path: directional index, candidate loop, unit
finder, wrap-around DP, chaining, formatted TSV.  Every rank runs the same step shape on its own reads (weak
scaling, no collective on the data path; torch.distributed is used only for the barrier and the max over ranks).

  value : reads/s with the batch already 2-bit packed and resident in HBM when the clock starts (mtr_pipeline_run)
  e2e   : reads/s of handle_one_file() -- the reference's own entry point (mTR.h:126) -- on a FASTA file (tmpfs) holding
          the K timed batches, stdout captured: parse, stale-state tracking, pack, H2D, every per-round H2D/D2H, chaining,
          formatting, ordered output; byte counts from mtr_file_stats
  roofline     : the dominant kernel (K3 wrap-around DP fill): algorithmic cell updates / CUDA-event kernel time
                 against the integer-ALU issue ceiling measured on this GPU by mtr_alu_probe
  roofline_di  : the directional-index kernels against measured HBM bandwidth (reported for completeness: that
                 stage is LSU/shared-memory bound, not HBM bound)
  cpu_baseline : oracle/_ref/mTR_ref_O3 (the unmodified reference, -O3) on a bounded sample of the same workload
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from mtr_b200 import synth  # noqa: E402

I_CELL_INT32 = 15.0     # integer instructions per cell update with full reference semantics
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "mTR_ref_O3")
ORACLE_BIN = os.path.join(ROOT, "oracle", "mtr_oracle")


def fasta_bytes(reads, first_id=0):
    return "".join(">%d\n%s\n" % (first_id + i, synth.to_text(r)) for i, r in enumerate(reads)).encode()


def step_reads(n, rank, step):
    return synth.long_reads(n, seed=1000 + 10007 * rank + step)[0]


def _text_chunk(args):
    n, rank, step, chunk, first_id = args
    # chunk c of a step has its own seed: the workload is defined by (reads per step, rank, step), not by the pool size
    return fasta_bytes(synth.long_reads(n, seed=1000 + 10007 * rank + step + 7919 * chunk)[0], first_id)


def make_texts(R, rank, total, chunk_reads=4096):
    """FASTA text of every step (R reads each), generated in chunks of chunk_reads reads by a process pool: chunk 0 of a
    step is exactly step_reads(chunk_reads, rank, step), so a 4096-read step is the same workload as before."""
    import multiprocessing as mp
    jobs = []
    for s in range(total):
        for c, first in enumerate(range(0, R, chunk_reads)):
            jobs.append((min(chunk_reads, R - first), rank, s, c, first))
    procs = max(1, min(len(jobs), (os.cpu_count() or 1) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))
    if procs == 1 or len(jobs) == 1:
        parts = [_text_chunk(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_text_chunk, jobs)
    per = (R + chunk_reads - 1) // chunk_reads
    return [b"".join(parts[s * per:(s + 1) * per]) for s in range(total)]


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.p = None

    def _pump(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference_cli(binary, reads, cores, flags=()):
    """Times the reference CLI on `reads`, split into `cores` contiguous chunks run concurrently (the reference
    is single-threaded and non-reentrant: process-level sharding is its only multi-core mode)."""
    with tempfile.TemporaryDirectory() as tmp:
        chunks = [c for c in np.array_split(np.arange(len(reads)), cores) if len(c)]
        paths = []
        for ci, idx in enumerate(chunks):
            p = os.path.join(tmp, "c%d.fa" % ci)
            synth.write_fasta(p, [reads[i] for i in idx], ids=[int(i) for i in idx])
            paths.append(p)
        t0 = time.perf_counter()
        procs = [subprocess.Popen([binary] + list(flags) + [p], stdout=open(p + ".out", "wb"), stderr=subprocess.DEVNULL) for p in paths]
        for pr in procs:
            pr.wait()
        dt = time.perf_counter() - t0
        out = b"".join(open(p + ".out", "rb").read() for p in paths)
        assert all(pr.returncode == 0 for pr in procs)
    return dt, out


def reference_arm(a, rank, world):
    """--impl reference: the reference's own CPU implementation on the box's host cores, same workload."""
    if rank != 0:
        return
    binary, kind = (REF_BIN, "reference") if os.path.exists(REF_BIN) else (ORACLE_BIN, "port")
    if not os.path.exists(binary):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    cores = os.cpu_count() or 1
    n = max(cores, min(a.reads, 8 * cores))              # bounded sample: ~0.3 s of CPU per read, 8 reads per core per step
    times = []
    for step in range(a.warmup + a.steps):
        reads = step_reads(n, 0, step)
        dt, _ = run_reference_cli(binary, reads, cores)
        if step >= a.warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    v = n / per_step
    line = {"impl": "reference", "metric": "reads_per_s", "value": round(v, 3), "unit": "reads/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(per_step * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "C5 synthetic long reads 10-20 kb, unit 2-500 bp, 5-15% noise", "reads_per_step": n,
                       "mode": "default (Manhattan), -m 0.6"},
            "cpu_baseline": {"value": round(v, 3), "unit": "reads/s", "cores": cores, "kind": kind,
                             "sample": "%d reads per step, split over %d processes" % (n, cores)},
            "e2e": {"value": round(v, 3), "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=int(os.environ.get("MTR_BENCH_READS", "8192")), help="reads per step per GPU")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the cpu_baseline sample (0 = 8 x cores)")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("MTR_BENCH_INFLIGHT", "1")), help="batches in flight per GPU (pipelines run concurrently)")
    ap.add_argument("--quick", action="store_true", help="tuning aid: resident loop only (no e2e, no replay, no CPU baseline); not a bench line")
    ap.add_argument("--threads", type=int, default=0, help="host worker threads per GPU (0 = cores / GPUs)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        reference_arm(a, rank, world)
        return

    # synthetic input first: the process pool forks before CUDA is initialised
    texts = make_texts(a.reads, rank, a.warmup + a.steps)

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    from mtr_b200 import capi
    threads = a.threads or max(1, (os.cpu_count() or 1) // world)
    ctx = capi.Context(local_rank)
    alu = {k: ctx.alu_probe(i) for i, k in enumerate(("viaddmnmx_s32", "lop3_iadd", "viaddmnmx_s16x2"))}
    ctx.close()

    R = a.reads
    total = a.warmup + a.steps
    keys = ("reads", "bases", "candidates", "rounds", "rounds_fast", "rounds_uf", "uf_tasks", "uf_kernel_ms", "uf_wall_ms", "jobs", "wdp_calls", "wdp_cells", "wdp_slot_cells", "wdp_dir_bytes",
            "di_position_passes", "di_bytes_in", "di_bytes_out", "h2d_bytes", "d2h_bytes", "launches", "spec_cells", "wdp_fill_ms",
            "wdp_tb_ms", "di_kernel_ms", "di_wall_ms", "rounds_wall_ms", "host_step_ms", "wdp_wall_ms")

    # Batches in flight (--inflight, default 1; handle_one_file: MTR_INFLIGHT_PER_GPU): with more than one, each batch
    # runs on its own pipeline object so that the ramp-down of one batch overlaps the ramp-up of the next.  Measured on
    # B200 + 16 host cores this loses (the long-job lanes of the two batches slow each other down), hence the default.
    inflight = max(1, a.inflight)
    n_pipes = 1 if inflight == 1 else max(inflight, min(a.steps, 2 * inflight))   # batches resident at once in the resident loop
    pipes = [capi.Pipeline(local_rank, threads=threads) for _ in range(n_pipes)]

    def run_concurrently(work):
        """work: list of callables; `inflight` threads, thread t takes items t, t + inflight, ... (ctypes drops the GIL)."""
        errs = []

        def loop(t):
            try:
                for i in range(t, len(work), inflight):
                    work[i]()
            except BaseException as e:      # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=loop, args=(t,)) for t in range(inflight)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errs:
            raise errs[0]

    # ---- warm-up: every pipeline sees at least one batch (buffers grown, kernels loaded)
    n_warm = max(a.warmup, n_pipes)
    run_concurrently([(lambda s=s: (pipes[s % n_pipes].load_fasta(texts[s % max(a.warmup, 1)]), pipes[s % n_pipes].run())) for s in range(n_warm)])

    # ---- device-resident loop: the batches are packed and in HBM before the clock starts
    acc = dict.fromkeys(keys, 0.0)
    lock = threading.Lock()
    outs = {}
    t_res = 0.0
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    timed = list(range(a.warmup, total))
    for g0 in range(0, len(timed), n_pipes):
        group = timed[g0:g0 + n_pipes]
        for i, s in enumerate(group):
            pipes[i].load_fasta(texts[s])
        torch.cuda.synchronize()

        def one(i, s):
            out = pipes[i].run()
            st = pipes[i].stats()
            with lock:
                outs[s] = out
                for k in keys:
                    acc[k] += st[k]
        t0 = time.perf_counter()
        run_concurrently([(lambda i=i, s=s: one(i, s)) for i, s in enumerate(group)])
        torch.cuda.synchronize()
        t_res += time.perf_counter() - t0
    barrier()
    t_res = max_over_ranks(t_res)
    digest = hashlib.md5()
    for s in timed:
        digest.update(outs[s])
    if a.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "reads_per_s": round(R * a.steps * world / t_res, 1), "ms_per_step": round(t_res / a.steps * 1e3, 1),
                              "fill_gcups": round(acc["wdp_cells"] / max(acc["wdp_fill_ms"], 1e-9) / 1e6, 1), "rounds_per_step": int(acc["rounds"] / a.steps),
                              "wdp_fill_ms": round(acc["wdp_fill_ms"] / a.steps, 1), "wdp_tb_ms": round(acc["wdp_tb_ms"] / a.steps, 1),
                              "md5": digest.hexdigest(), "inflight": inflight, "env": {k: v for k, v in os.environ.items() if k.startswith("MTR_")}}), flush=True)
        clocks.stop()
        for p in pipes:
            p.close()
        return

    # ---- K3 alone on exactly one step's DP jobs, replayed as a single batch (operands resident in HBM)
    for p in pipes[1:]:
        p.close()
    alone = alone_fused = None
    if rank == 0:
        pipes[0].log_jobs(True)
        pipes[0].load_fasta(texts[a.warmup])
        pipes[0].run()
        pipes[0].log_jobs(False)
        alone = pipes[0].replay_logged_jobs(iters=3, fused=False)
        alone_fused = pipes[0].replay_logged_jobs(iters=3, fused=True)
    pipes[0].close()

    # ---- end to end through the reference's own entry point: handle_one_file(path) (mTR.h:126) on a FASTA file holding
    # the K timed batches, exactly as main.c calls it -- parse, stale-state tracking, 2-bit pack, H2D, directional index,
    # every per-round H2D / D2H, chaining, formatting, ordered stdout.  The entry point owns its engines (two per GPU:
    # the next batch starts when the running one is down to its last reads); the pipelines above are closed first.
    need = 2 * world * max(4, a.steps) * max(len(t) for t in texts)          # every rank writes its own file; 2x headroom
    shm = tempfile.gettempdir()
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize >= need:
            shm = "/dev/shm"
    except OSError:
        pass
    e2e_path = os.path.join(shm, "mtr_bench_rank%d_%d.fa" % (rank, os.getpid()))
    os.environ.setdefault("MTR_DEVICE", str(local_rank))
    os.environ.setdefault("MTR_THREADS", str(threads))
    os.environ.setdefault("MTR_BATCH_READS", str(R))
    os.environ.setdefault("MTR_BATCH_MBASES", str(max(64, (R * 21000) >> 20)))      # a batch of R reads of <= 20 kb must not be cut by the base cap
    with open(e2e_path, "wb") as f:                      # warm-up file: two batches per engine (buffers grown, pinned memory mapped)
        for s in range(4):
            f.write(texts[s % max(a.warmup, 1)])
    capi.run_file(e2e_path)
    with open(e2e_path, "wb") as f:
        for s in timed:
            f.write(texts[s])
    barrier()
    t0 = time.perf_counter()
    n_file, out_file, fst = capi.run_file(e2e_path)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    os.unlink(e2e_path)
    assert n_file == R * a.steps, (n_file, R, a.steps)
    h2d, d2h = fst["h2d_bytes"], fst["d2h_bytes"]
    e2e_md5 = hashlib.md5(out_file).hexdigest()
    clk = clocks.stop()

    # algorithmic cells = what the reference executes: cells spent on speculative candidates that were pruned after all
    # (MTR_SPECULATE) are launched and timed but not counted
    acc["wdp_cells_launched"] = acc["wdp_cells"]
    acc["wdp_cells"] = acc["wdp_cells"] - acc["spec_cells"]
    reads_all = sum_over_ranks(R * a.steps)
    cells_all = sum_over_ranks(acc["wdp_cells"])
    kernel_ms = acc["wdp_fill_ms"] + acc["wdp_tb_ms"]
    gcups_rank = acc["wdp_cells"] / max(kernel_ms, 1e-9) / 1e6
    fill_gcups = acc["wdp_cells"] / max(acc["wdp_fill_ms"], 1e-9) / 1e6
    peak_gcups = alu["viaddmnmx_s32"] / I_CELL_INT32
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    di_bytes = acc["di_bytes_in"] + acc["di_bytes_out"]
    di_gbs = di_bytes / max(acc["di_kernel_ms"], 1e-9) / 1e6

    line = {
        "metric": "reads_per_s", "value": round(reads_all / t_res, 3), "unit": "reads/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": round(t_res / a.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "C5 synthetic long reads 10-20 kb, unit 2-500 bp, 5-15% noise", "reads_per_step_per_gpu": R,
                   "mode": "default (Manhattan), -m 0.6", "host_threads_per_gpu": threads, "batches_in_flight_per_gpu": inflight,
                   "l2": "working set per step (direction matrices, %d MB) exceeds the 126 MB L2" % (acc["wdp_dir_bytes"] / a.steps / 2 ** 20)},
        "e2e": {"value": round(reads_all / t_e2e, 3), "unit": "reads/s", "h2d_bytes_per_step": int(h2d / a.steps),
                "d2h_bytes_per_step": int(d2h / a.steps), "ms_per_step": round(t_e2e / a.steps * 1e3, 2),
                "how": "handle_one_file() on a FASTA file (tmpfs) with the K timed batches, stdout captured; parse, pack, every H2D/D2H inside",
                "output_md5": e2e_md5},
        "gpu_launches": int(acc["launches"]),
        "gcups": {"wrap_around_dp": round(gcups_rank, 2), "fill_only": round(fill_gcups, 2),
                  "whole_job_cells_per_s": round(cells_all / t_res / 1e9, 3), "algorithmic_cells_per_step": int(acc["wdp_cells"] / a.steps),
                  "speculative_cells_per_step": int(acc["spec_cells"] / a.steps)},
        "roofline": {"kernel": "wdp_fill_* (K3 wrap-around DP: fill with the traceback fused in), all launches of the timed steps", "bound": "int-alu", "achieved": round(fill_gcups, 2),
                     "peak": round(peak_gcups, 1), "unit": "GCUPS", "frac": round(fill_gcups / peak_gcups, 4),
                     "traffic": {"dram_bytes_per_cell_ncu": 0.289, "algorithmic_dir_bytes_per_cell": round(acc["wdp_dir_bytes"] / max(acc["wdp_cells"], 1), 3),
                                 "source": "profiles/r1_final_wdp_fill_summary.md (ncu --set full: 3.84 GB DRAM for 13.29 G cells)"},
                     "peak_how": "mtr_alu_probe: %.0f G lane-ops/s VIADDMNMX.RELU measured now / %.0f instr per cell (SURVEY.md 8(d))"
                                 % (alu["viaddmnmx_s32"], I_CELL_INT32),
                     "alu_probe_gops": {k: round(v, 1) for k, v in alu.items()},
                     "dir_bytes_per_cell": round(acc["wdp_dir_bytes"] / max(acc["wdp_cells"], 1), 3)},
        "roofline_kernel_alone": None if alone is None else {
            "what": "the same K3 kernels on all DP jobs of one step replayed as ONE batch (%d jobs), timed alone with CUDA events; "
                    "fill and traceback launched separately here so that achieved = fill alone" % alone["jobs"],
            "bound": "int-alu", "achieved": round(alone["wdp_cells"] / max(alone["wdp_fill_ms"], 1e-9) / 1e6, 2), "peak": round(peak_gcups, 1),
            "unit": "GCUPS", "frac": round(alone["wdp_cells"] / max(alone["wdp_fill_ms"], 1e-9) / 1e6 / peak_gcups, 4),
            "with_traceback_gcups": round(alone["wdp_cells"] / max(alone["wdp_fill_ms"] + alone["wdp_tb_ms"], 1e-9) / 1e6, 2),
            "fill_ms": round(alone["wdp_fill_ms"], 3), "tb_ms": round(alone["wdp_tb_ms"], 3), "cells": int(alone["wdp_cells"]),
            "slot_cells": int(alone["wdp_slot_cells"]), "dir_bytes": int(alone["wdp_dir_bytes"]),
            "fused_fill_plus_traceback_ms": round(alone_fused["wdp_fill_ms"] + alone_fused["wdp_tb_ms"], 3),
            "fused_gcups": round(alone_fused["wdp_cells"] / max(alone_fused["wdp_fill_ms"] + alone_fused["wdp_tb_ms"], 1e-9) / 1e6, 2)},
        "roofline_di": {"kernel": "di_codes + di_slide + di_merge (K1/K2)", "bound": "hbm", "achieved": round(di_gbs, 3),
                        "peak": hbm_peak, "unit": "GB/s", "frac": round(di_gbs / hbm_peak, 6), "traffic": None,
                        "note": "algorithmic bytes = packed reads in + 16 B per position out; the stage is LSU/shared-memory bound"},
        "breakdown_ms_per_step": {k: round(acc[k] / a.steps, 2) for k in ("di_wall_ms", "rounds_wall_ms", "host_step_ms", "wdp_wall_ms", "uf_wall_ms",
                                                                         "wdp_fill_ms", "wdp_tb_ms", "di_kernel_ms", "uf_kernel_ms")},
        "rounds_per_step": int(acc["rounds"] / a.steps), "fast_lane_rounds_per_step": int(acc["rounds_fast"] / a.steps), "uf_rounds_per_step": int(acc["rounds_uf"] / a.steps), "uf_tasks_per_step": int(acc["uf_tasks"] / a.steps), "dp_jobs_per_step": int(acc["jobs"] / a.steps),
        "clocks": clk, "output_md5": digest.hexdigest(),
    }

    if rank == 0:
        cores = os.cpu_count() or 1
        n = a.cpu_sample or 8 * cores
        binary, kind = (REF_BIN, "reference") if os.path.exists(REF_BIN) else (ORACLE_BIN, "port")
        sample = step_reads(n, 0, a.warmup)[:n] if n <= R else step_reads(n, 0, a.warmup)
        dt, ref_out = run_reference_cli(binary, sample, cores)
        line["cpu_baseline"] = {"value": round(len(sample) / dt, 3), "unit": "reads/s", "cores": cores, "kind": kind,
                                "sample": "%d reads of the step's workload (seed of step %d), %d processes, %.1f s" % (len(sample), a.warmup, cores, dt)}
        # parity at bench time, against the reference itself: the first reads of the timed step through the reference
        # sources (oracle/_ref/mTR_ref_det: the reference with its alignment set in insertion order; one
        # process, fresh state) and through a fresh pipeline on this GPU -- the two texts must be byte-identical
        try:
            par_n = min(32, R)
            par_reads = step_reads(par_n, 0, a.warmup)           # the generator is prefix-stable: these are the step's first reads
            det = os.path.join(ROOT, "oracle", "_ref", "mTR_ref_det")
            det = det if os.path.exists(det) else ORACLE_BIN
            with tempfile.TemporaryDirectory() as tmp:
                pth = os.path.join(tmp, "parity.fa")
                synth.write_fasta(pth, par_reads, ids=list(range(par_n)))
                want = subprocess.run([det, pth], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            pp = capi.Pipeline(local_rank, threads=threads)
            pp.load_fasta(fasta_bytes(par_reads))
            got = pp.run()
            pp.close()
            line["parity"] = {"reads": par_n, "against": os.path.relpath(det, ROOT), "identical": got == want, "records": want.count(b"\n"),
                              "md5": hashlib.md5(got).hexdigest()}
        except Exception as e:      # noqa: BLE001 -- the bench line must still be printed
            line["parity"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
