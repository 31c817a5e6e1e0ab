"""Shared job builders for the wrap-around DP parity tests (CPU oracle vs CUDA kernel)."""
from __future__ import annotations

import numpy as np

from mtr_b200 import capi, synth

PARAM_SETS = [(1, 1, 3), (1, 3, 1), (5, 1, 1)]


def random_jobs(rng, n_reads=6, n_jobs=300, max_rows=700, max_ulen=499, tail=True):
    """Random windows over reads that contain tandem repeats, units partly derived from the read."""
    reads = []
    for r in range(n_reads):
        ulen = int(rng.integers(2, 120))
        rd, _ = synth.rand_seq_reads(ulen, int(rng.integers(5, 30)), 0.03, 0.05, 0.05, 150, 150, 1, seed=int(rng.integers(1 << 30)))
        reads.append(rd[0])
    tails = [(int(rng.integers(4)), int(rng.integers(4))) if tail else (0, 0) for _ in reads]
    jobs = []
    for _ in range(n_jobs):
        r = int(rng.integers(n_reads))
        L = len(reads[r])
        rows = int(rng.integers(1, min(max_rows, L) + 1))
        first = int(rng.integers(-1, L + 1 - rows + 1))      # first + rows <= L + 1
        # ulen <= rows: the reference only clears row 0 for j <= rows (wrap_around_DP.c:250), so a unit longer
        # than the window reads stale cells of an earlier DP; the pipeline never produces such a job
        # (period <= width / 5, consensus.c:283) and the ABI defines row 0 as zero.
        ulen = int(min(max_ulen, rows, np.exp(rng.uniform(0, np.log(max_ulen + 1)))))
        ulen = max(1, ulen)
        if rng.random() < 0.6 and ulen <= L:
            s = int(rng.integers(0, L - ulen + 1))
            unit = np.array(reads[r][s:s + ulen], dtype=np.uint8)
        else:
            unit = rng.integers(0, 4, ulen).astype(np.uint8)
        g, mm, ind = PARAM_SETS[int(rng.integers(3))]
        jobs.append(dict(read=r, first=first, rows=rows, unit=unit, gain=g, mis=mm, indel=ind))
    return reads, tails, jobs


def window(reads, tails, job):
    """The read bases x_1..x_rows of a job, with the two stale tail bases appended to the read."""
    rd = np.concatenate([np.asarray(reads[job["read"]], dtype=np.int32), np.asarray(tails[job["read"]], dtype=np.int32)])
    return rd[job["first"] + 1: job["first"] + 1 + job["rows"]]


def build_job_array(jobs, mode=capi.TB_COUNTS, pair=False):
    """-> (JOB_DTYPE array, units uint8, aux_bytes).  pair=True: every job gets both (1,1,3) and (1,3,1)."""
    arr = np.zeros(len(jobs), dtype=capi.JOB_DTYPE)
    units, off, aux = [], 0, 0
    for i, j in enumerate(jobs):
        a = arr[i]
        a["read"], a["first"], a["rows"], a["unit_off"], a["ulen"] = j["read"], j["first"], j["rows"], off, len(j["unit"])
        if pair:
            a["gain"], a["mis"], a["indel"], a["n_param"] = (1, 1), (1, 3), (3, 1), 2
        else:
            a["gain"], a["mis"], a["indel"], a["n_param"] = (j["gain"], 0), (j["mis"], 0), (j["indel"], 0), 1
        a["mode"] = mode
        if mode == capi.TB_CONSENSUS:
            a["aux_off"] = aux // 4
            aux += (len(j["unit"]) + 1) * 9 * 4
        elif mode == capi.TB_PATH:
            cap = j["rows"] * 2 + 64
            a["aux_off"], a["aux_cap"] = aux, cap
            aux += cap
        units.append(j["unit"]); off += len(j["unit"])
    return arr, np.concatenate(units).astype(np.uint8) if units else np.zeros(0, np.uint8), aux


FIELDS = ("best", "max_i", "max_j", "end_i", "end_j", "n_match", "n_mismatch", "n_ins", "n_del", "n_scanned", "path_len")


def check_against_oracle(oracle, reads, tails, jobs, res, aux=None, arr=None, mode=capi.TB_COUNTS, pair=False):
    """Bit-exact comparison of every result field (and histograms / paths) with the CPU oracle."""
    bad = []
    for i, j in enumerate(jobs):
        x = window(reads, tails, j)
        sets = [(1, 1, 3), (1, 3, 1)] if pair else [(j["gain"], j["mis"], j["indel"])]
        for p, (g, mm, ind) in enumerate(sets):
            exp = oracle.wrap_dp(x, j["unit"], g, mm, ind, mode=mode)
            got = res[i, p]
            for f in FIELDS:
                if int(got[f]) != exp[f]:
                    bad.append((i, p, f, int(got[f]), exp[f], j["rows"], len(j["unit"]), (g, mm, ind)))
                    break
            if got["flags"] != 0:
                bad.append((i, p, "flags", int(got["flags"]), 0))
            if mode == capi.TB_CONSENSUS:
                U = len(j["unit"])
                blk = aux.view(np.int32)[arr[i]["aux_off"]: arr[i]["aux_off"] + (U + 1) * 9]
                if not (np.array_equal(blk[:(U + 1) * 5].reshape(U + 1, 5), exp["consensus"]) and
                        np.array_equal(blk[(U + 1) * 5:].reshape(U + 1, 4), exp["missing"])):
                    bad.append((i, p, "hist"))
            if mode == capi.TB_PATH:
                pth = aux[arr[i]["aux_off"]: arr[i]["aux_off"] + exp["path_len"]]
                if not np.array_equal(pth, exp["path"]):
                    bad.append((i, p, "path"))
    return bad
