"""Kernel-level throughput probe for K3 (not the bench): harvests the DP jobs the reference executes on a few
C5-shaped reads (via the oracle), replicates them to fill the GPU and reports GCUPS per launch."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
import wdp_cases  # noqa: E402
from mtr_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=6)
    ap.add_argument("--rep", type=int, default=16)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--pair", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--kind", default="long")
    a = ap.parse_args()
    if a.kind == "long":
        reads, _ = synth.long_reads(a.reads, seed=a.seed)
    else:
        reads, _ = synth.rand_seq_reads(100, 10, 0.016, 0.09, 0.038, 1000, 1000, a.reads, seed=a.seed)
    t0 = time.time()
    o = oracle_lib.Oracle()
    hj, tails = o.harvest_dp_jobs(reads)
    st = o.stats(); o.close()
    t_cpu = time.time() - t0
    jobs = [j for j in hj if j["kind"] == 0]
    if a.pair:      # consecutive (1,1,3),(1,3,1) calls of wrap_around_DP collapse into one paired job
        jobs = [j for j in jobs if (j["gain"], j["mis"], j["indel"]) == (1, 1, 3)]
    print("harvested %d dp jobs (%d kept) from %d reads in %.1fs CPU; oracle cells dp=%d revise=%d" %
          (len(hj), len(jobs), len(reads), t_cpu, st["dp_cells"], st["revise_cells"]), flush=True)
    jobs = jobs * a.rep
    ctx = capi.Context(0)
    packed, woff, lens = capi.pack_reads(reads, tails)
    ctx.upload_reads(packed, woff, lens)
    arr, units, aux = wdp_cases.build_job_array(jobs, pair=bool(a.pair))
    ctx.wdp_upload(arr, units, aux)
    for it in range(a.iters):
        ctx.wdp_launch()
        s = ctx.stats()
        print(json.dumps(dict(it=it, fill_ms=round(s["wdp_fill_ms"], 3), tb_ms=round(s["wdp_tb_ms"], 3), cells=s["wdp_cells"],
                              slot_cells=s["wdp_slot_cells"], dir_MB=s["wdp_dir_bytes"] >> 20, jobs=len(jobs),
                              gcups_fill=round(s["wdp_cells"] / s["wdp_fill_ms"] / 1e6, 1),
                              gcups_total=round(s["wdp_cells"] / (s["wdp_fill_ms"] + s["wdp_tb_ms"]) / 1e6, 1))), flush=True)
    ulen = np.array([len(j["unit"]) for j in jobs]); rows = np.array([j["rows"] for j in jobs])
    for lo, hi in ((1, 16), (17, 32), (33, 64), (65, 128), (129, 256), (257, 512)):
        msk = (ulen >= lo) & (ulen <= hi)
        print("ulen %3d-%3d: jobs %7d cells %.3e meanrows %.0f" % (lo, hi, msk.sum(), float((ulen[msk] * rows[msk]).sum()), rows[msk].mean() if msk.any() else 0))
    ctx.close()


if __name__ == "__main__":
    main()
