#!/usr/bin/env python
"""bench.py -- reads/s and wrap-around-DP GCUPS of the mTR hot path on B200, beside the reference on the host CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl reference] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of R (default 8192) synthetic C5 reads (BASELINE.json configs[4]: 10-20 kb reads carrying one tandem repeat,
unit 2-500 bp, 5-15 % sub/ins/del noise) through the whole per-read path: directional index, candidate loop, unit
finder, wrap-around DP, chaining, formatted TSV.  Every rank runs the same step shape on its own reads (weak
scaling, no collective on the data path; torch.distributed is used only for the barrier and the max over ranks).

  value : reads/s with the K timed batches already 2-bit packed and resident in HBM when the clock starts (one
          mtr_pipeline_run over all of them: directional index, engine waves, D2H of the accepted repeats, chaining, text)
  e2e   : reads/s of handle_one_file() -- the reference's own entry point (mTR.h:126) -- on a FASTA file (tmpfs) holding
          the K timed batches, stdout captured: parse, stale-state tracking, pack, H2D, engine, D2H, chaining,
          formatting, ordered output; byte counts from mtr_file_stats
  roofline     : the dominant kernels (K3 wrap-around DP fill + traceback): algorithmic cell updates / the UNION of the
                 K3 kernel intervals of all engine contexts, against the integer-ALU issue ceiling measured on this GPU
                 by mtr_alu_probe; frac_by_wall divides by the wall clock of the step instead
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
    ap.add_argument("--parity-reads", type=int, default=512, help="reads of the timed workload compared with the reference at bench time")
    ap.add_argument("--quick", action="store_true", help="tuning aid: resident loop only (no e2e, no CPU baseline); not a bench line")
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
    lib = capi.load_library()
    ctx = capi.Context(local_rank)
    alu = {k: ctx.alu_probe(i) for i, k in enumerate(("viaddmnmx_s32", "lop3_iadd", "viaddmnmx_s16x2"))}
    ctx.close()

    R = a.reads
    total = a.warmup + a.steps
    keys = [k for k, _ in capi.PipelineStats._fields_]
    pipe = capi.Pipeline(local_rank)

    # ---- warm-up: the W warm-up steps through the pipeline as ONE load, the shape of the timed region (every buffer of every
    # engine context reaches its final size, kernels are loaded): no allocation inside the timed region
    if a.warmup > 0:
        pipe.load_fasta(b"".join(texts[s] for s in range(a.warmup)))
        pipe.run()

    # ---- device-resident loop: ALL timed steps are parsed, 2-bit packed and uploaded before the clock starts (K x R reads
    # resident in HBM, cut into groups of MTR_GROUP_READS reads dealt round-robin to the engine contexts); the timed region
    # is one mtr_pipeline_run over them -- directional index, waves, D2H of the accepted repeats, chaining, formatting
    timed = list(range(a.warmup, total))
    clocks = ClockSampler(local_rank)
    clocks.start()
    pipe.load_fasta(b"".join(texts[s] for s in timed))
    lib.mtr_engine_dp_busy_ms(local_rank, 1)
    barrier()
    t0 = time.perf_counter()
    out_res = pipe.run()
    torch.cuda.synchronize()
    t_res = time.perf_counter() - t0
    barrier()
    dp_busy_ms = lib.mtr_engine_dp_busy_ms(local_rank, 1)
    t_res = max_over_ranks(t_res)
    acc = pipe.stats()
    res_md5 = hashlib.md5(out_res).hexdigest()
    pipe.close()
    if a.quick:
        if rank == 0:
            cells = acc["wdp_cells"] - acc["spec_cells"]
            print(json.dumps({"quick": True, "cells_run": int(cells), "dir_bytes": int(acc["wdp_dir_bytes"]), "shared_cells_frac": round(acc["shared_cells"] / max(cells + acc["shared_cells"], 1), 3), "reads_per_s": round(R * a.steps * world / t_res, 1), "ms_per_step": round(t_res / a.steps * 1e3, 1),
                              "dp_busy_ms_per_step": round(dp_busy_ms / a.steps, 1), "gcups_busy": round(cells / max(dp_busy_ms, 1e-9) / 1e6, 1),
                              "gcups_wall": round(cells / t_res / 1e9, 1), "waves_per_group": round(acc["waves"] / max(acc["groups"], 1), 1),
                              "groups": acc["groups"], "md5": res_md5, "env": {k: v for k, v in os.environ.items() if k.startswith("MTR_")}}), flush=True)
        clocks.stop()
        return

    # ---- end to end through the reference's own entry point: handle_one_file(path) (mTR.h:126) on a FASTA file holding
    # the K timed batches, exactly as main.c calls it -- parse, stale-state tracking, 2-bit pack, H2D, directional index,
    # waves, D2H, chaining, formatting, ordered stdout.  The entry point owns its own engine contexts.
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
    with open(e2e_path, "wb") as f:                      # warm-up file: the shape of the timed file
        for s in range(max(1, a.warmup)):
            f.write(texts[s])
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

    # cells: `run` = what the K3 kernels computed for candidates the reference visits too (cells of look-ahead candidates that
    # were pruned after all are launched and timed but not counted); `reference` = what the reference executes for the
    # same reads = run + the search DPs that were NOT launched because a sibling chain ran the identical DP
    cells = acc["wdp_cells"] - acc["spec_cells"]
    cells_ref = cells + acc["shared_cells"]
    reads_all = sum_over_ranks(R * a.steps)
    cells_all = sum_over_ranks(cells)
    # K3 time of the timed region: the UNION of the intervals in which fill / traceback kernels of any engine context of this
    # GPU were running (several queues and contexts share the GPU, their K3 phases overlap) -- never more than the wall clock
    gcups_busy = cells / max(dp_busy_ms, 1e-9) / 1e6
    gcups_wall = cells / t_res / 1e9
    # ceiling: integer-ALU issue rate measured now / instructions per cell: 15 for the int32 kernels, 7.5
    # for the paired int16x2 kernels (one VIADDMNMX.S16x2 serves both penalty sets); harmonic mix by the cells of each family
    peak_i32 = alu["viaddmnmx_s32"] / I_CELL_INT32
    peak_p16 = alu["viaddmnmx_s16x2"] / (I_CELL_INT32 / 2)
    f16 = acc["wdp_cells_p16"] / max(acc["wdp_cells"], 1)
    peak_gcups = 1.0 / ((1.0 - f16) / peak_i32 + f16 / peak_p16)
    traffic = traffic_detail = None
    try:                                                       # measured DRAM traffic of the fill kernels: profiles/r2_k3_traffic.json (ncu, all K3 launches of a run)
        traffic_detail = json.load(open(os.path.join(ROOT, "profiles", "r2_k3_traffic.json")))
        # per launch, like the contract asks: dram__bytes_read.sum + dram__bytes_write.sum averaged over the profiled K3 launches
        traffic = round((traffic_detail["dram_bytes_read"] + traffic_detail["dram_bytes_write"]) / max(traffic_detail["k3_launches"], 1))
    except (OSError, ValueError, KeyError, TypeError):
        pass

    line = {
        "metric": "reads_per_s", "value": round(reads_all / t_res, 3), "unit": "reads/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": round(t_res / a.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "C5 synthetic long reads 10-20 kb, unit 2-500 bp, 5-15% noise", "reads_per_step_per_gpu": R,
                   "mode": "default (Manhattan), -m 0.6", "engine_contexts_per_gpu": int(os.environ.get("MTR_GROUPS_PER_GPU", "2")),
                   "read_slots_per_context": int(os.environ.get("MTR_ENGINE_SLOTS", "16384")),
                   "groups": "the resident reads are cut into equal groups, one or more per context (<= 320 Mbases each)",
                   "l2": "working set per step (direction matrices, %d MB) exceeds the 126 MB L2" % (acc["wdp_dir_bytes"] / a.steps / 2 ** 20)},
        "e2e": {"value": round(reads_all / t_e2e, 3), "unit": "reads/s", "h2d_bytes_per_step": int(h2d / a.steps),
                "d2h_bytes_per_step": int(d2h / a.steps), "ms_per_step": round(t_e2e / a.steps * 1e3, 2),
                "how": "handle_one_file() on a FASTA file (tmpfs) with the K timed batches, stdout captured; parse, pack, every H2D/D2H inside",
                "output_md5": e2e_md5},
        "gpu_launches": int(acc["launches"]),
        "gcups": {"k3_busy": round(gcups_busy, 2), "whole_job_cells_per_s": round(cells_all / t_res / 1e9, 3),
                  "cells_run_per_step": int(cells / a.steps), "cells_reference_per_step": int(cells_ref / a.steps),
                  "cells_shared_per_step": int(acc["shared_cells"] / a.steps), "speculative_cells_per_step": int(acc["spec_cells"] / a.steps),
                  "reference_cells_per_s": round(sum_over_ranks(cells_ref) / t_res / 1e9, 3), "int16x2_share_of_cells": round(f16, 3)},
        "roofline": {"kernel": "wdp_fill_family<int32 | int16x2> (K3 wrap-around DP: fill + fused traceback), all launches of the timed region",
                     "bound": "int-alu", "achieved": round(gcups_busy, 2), "peak": round(peak_gcups, 1), "unit": "GCUPS",
                     "frac": round(gcups_busy / peak_gcups, 4),
                     "how": "cells run / union of the K3 kernel intervals of all queues and contexts (%.1f ms of %.1f ms wall per step)"
                            % (dp_busy_ms / a.steps, t_res / a.steps * 1e3),
                     "frac_by_wall": round(gcups_wall / peak_gcups, 4),
                     "traffic": traffic, "traffic_unit": "DRAM bytes per K3 launch (ncu)", "traffic_detail": traffic_detail, "algorithmic_dir_bytes_per_cell": round(acc["wdp_dir_bytes"] / max(acc["wdp_cells"], 1), 3),
                     "peak_how": "mtr_alu_probe now: %.0f G lane-ops/s VIADDMNMX.RELU / 15 instr per int32 cell = %.0f GCUPS, %.0f G VIADDMNMX.S16x2 / 7.5 per paired cell = %.0f GCUPS, "
                                 "harmonic mix by the cells of each family (%.0f %% int16x2)" % (alu["viaddmnmx_s32"], peak_i32, alu["viaddmnmx_s16x2"], peak_p16, 100 * f16),
                     "alu_probe_gops": {k: round(v, 1) for k, v in alu.items()}},
        "breakdown_ms_per_step": {"k3_busy_union": round(dp_busy_ms / a.steps, 2),
                                  **{k: round(acc[k] / a.steps, 2) for k in ("dp_ms", "di_kernel_ms", "uf_kernel_ms", "engine_wall_ms", "pack_ms", "chain_ms")}},
        "waves_per_group": round(acc["waves"] / max(acc["groups"], 1), 1), "groups_per_step": int(acc["groups"] / a.steps),
        "dp_jobs_per_step": int(acc["jobs"] / a.steps), "dp_tasks_per_step": int(acc["dp_tasks"] / a.steps),
        "count_tables_per_step": int(acc["tables"] / a.steps), "walks_per_step": int(acc["walks"] / a.steps),
        "clocks": clk, "output_md5": res_md5,
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
        # sources (oracle/_ref/mTR_ref_det: the reference with its alignment set in insertion order; chunks
        # run as separate processes, so the reads come in chunks with fresh state each) and through fresh pipelines on this
        # GPU -- the texts must be byte-identical
        try:
            par_n = min(a.parity_reads, R)
            par_reads = step_reads(par_n, 0, a.warmup)           # the generator is prefix-stable: these are the step's first reads
            det = os.path.join(ROOT, "oracle", "_ref", "mTR_ref_det")
            det = det if os.path.exists(det) else ORACLE_BIN
            chunks = [c for c in np.array_split(np.arange(par_n), cores) if len(c)]
            dt_par, want = run_reference_cli(det, par_reads, cores)
            got = b""
            for idx in chunks:                                    # same chunking: every chunk starts from a fresh process state
                pp = capi.Pipeline(local_rank)
                pp.load_fasta(fasta_bytes([par_reads[i] for i in idx], int(idx[0])))
                got += pp.run()
                pp.close()
            line["parity"] = {"reads": par_n, "against": os.path.relpath(det, ROOT), "identical": got == want, "records": want.count(b"\n"),
                              "md5": hashlib.md5(got).hexdigest(), "reference_seconds": round(dt_par, 1)}
        except Exception as e:      # noqa: BLE001 -- the bench line must still be printed
            line["parity"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
