"""K3 parity: CUDA wrap-around DP (through the C ABI) vs the CPU oracle, bit-exact on every field."""
import numpy as np
import pytest

import oracle_lib
import wdp_cases
from mtr_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    o = oracle_lib.Oracle()
    yield o
    o.close()


def _run(ctx, reads, tails, jobs, mode=capi.TB_COUNTS, pair=False):
    packed, woff, lens = capi.pack_reads(reads, tails)
    ctx.upload_reads(packed, woff, lens)
    arr, units, aux_bytes = wdp_cases.build_job_array(jobs, mode=mode, pair=pair)
    res, aux = ctx.wdp_run(arr, units, aux_bytes)
    return arr, res, aux


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_jobs_counts(gpu_ctx, oracle, seed):
    rng = np.random.default_rng(seed)
    reads, tails, jobs = wdp_cases.random_jobs(rng, n_jobs=400)
    arr, res, aux = _run(gpu_ctx, reads, tails, jobs)
    bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res)
    assert not bad, bad[:10]


def test_random_jobs_paired(gpu_ctx, oracle):
    rng = np.random.default_rng(11)
    reads, tails, jobs = wdp_cases.random_jobs(rng, n_jobs=300)
    arr, res, aux = _run(gpu_ctx, reads, tails, jobs, pair=True)
    bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res, pair=True)
    assert not bad, bad[:10]


def test_fused_and_split_traceback_agree(gpu_ctx, oracle, monkeypatch):
    """The fill kernels trace each task back themselves (default); the split mode (fill kernels + one traceback kernel,
    used to time the fill alone) must give the same bytes, in both kernel families, incl. histograms and paths."""
    import os
    rng = np.random.default_rng(77)
    reads, tails, jobs = wdp_cases.random_jobs(rng, n_jobs=400, max_rows=1200)
    try:
        for fam in ("latency", "throughput"):
            monkeypatch.setenv("MTR_WDP_MODE", fam)
            os.environ["MTR_WDP_MODE"] = fam
            for mode, pair in ((capi.TB_COUNTS, True), (capi.TB_COUNTS, False), (capi.TB_CONSENSUS, False), (capi.TB_PATH, False)):
                out = []
                for fused in (True, False):
                    gpu_ctx.wdp_set_fused_traceback(fused)
                    arr, res, aux = _run(gpu_ctx, reads, tails, jobs, mode=mode, pair=pair)
                    out.append((res.tobytes(), None if aux is None else aux.tobytes()))
                assert out[0] == out[1], (fam, mode, pair)
                if mode == capi.TB_COUNTS:
                    bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res, pair=pair)
                    assert not bad, (fam, pair, bad[:10])
    finally:
        gpu_ctx.wdp_set_fused_traceback(True)
        os.environ.pop("MTR_WDP_MODE", None)


@pytest.mark.parametrize("mode", ["throughput", "latency"])
def test_kernel_families(gpu_ctx, oracle, mode, monkeypatch):
    """Small batches default to the latency classes; force each family (throughput classes pair the two penalty sets
    of a wrap_around_DP call into one int16x2 task when the scores fit) and check both against the oracle."""
    import os
    monkeypatch.setenv("MTR_WDP_MODE", mode)
    os.environ["MTR_WDP_MODE"] = mode
    rng = np.random.default_rng(31)
    reads, tails, jobs = wdp_cases.random_jobs(rng, n_jobs=500, max_rows=1500)
    arr, res, aux = _run(gpu_ctx, reads, tails, jobs, pair=True)
    bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res, pair=True)
    assert not bad, (mode, bad[:10])
    arr, res, aux = _run(gpu_ctx, reads, tails, jobs)
    bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res)
    assert not bad, (mode, bad[:10])
    # adversarial ties in the paired kernels: homopolymer and dinucleotide reads, units AC / A / ACAC
    reads2 = [np.zeros(900, np.int8), np.tile(np.array([0, 1], np.int8), 600)]
    tails2 = [(0, 0), (1, 0)]
    jobs2 = [dict(read=r, first=f, rows=rows, unit=np.array(u, np.uint8), gain=1, mis=1, indel=3)
             for r in range(2) for u in ([0], [0, 1], [1, 0], [0, 1, 0, 1], [0, 0, 1]) for f, rows in ((-1, 900), (3, 500), (100, 37))]
    arr, res, aux = _run(gpu_ctx, reads2, tails2, jobs2, pair=True)
    bad = wdp_cases.check_against_oracle(oracle, reads2, tails2, jobs2, res, pair=True)
    assert not bad, (mode, bad[:10])
    os.environ.pop("MTR_WDP_MODE", None)


def test_paired_range_limit(gpu_ctx, oracle, monkeypatch):
    """Pairs with more than 8190 rows exceed int16 and must fall back to two int32 tasks; 8190 rows just fit."""
    import os
    os.environ["MTR_WDP_MODE"] = "throughput"
    rd, _ = synth.rand_seq_reads(40, 260, 0.005, 0.01, 0.01, 50, 50, 1, seed=19)
    reads, tails = [rd[0]], [(2, 3)]
    unit = np.array(rd[0][50:90], dtype=np.uint8)
    jobs = [dict(read=0, first=10, rows=r, unit=unit, gain=1, mis=1, indel=3) for r in (8189, 8190, 8191, 9000)]
    arr, res, aux = _run(gpu_ctx, reads, tails, jobs, pair=True)
    bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res, pair=True)
    os.environ.pop("MTR_WDP_MODE", None)
    assert not bad, bad
    assert res[1, 0]["best"] > 7000


def test_consensus_and_path(gpu_ctx, oracle):
    rng = np.random.default_rng(21)
    reads, tails, jobs = wdp_cases.random_jobs(rng, n_jobs=150, max_ulen=300)
    for mode in (capi.TB_CONSENSUS, capi.TB_PATH):
        arr, res, aux = _run(gpu_ctx, reads, tails, jobs, mode=mode)
        bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res, aux=aux, arr=arr, mode=mode)
        assert not bad, (mode, bad[:10])


def test_adversarial(gpu_ctx, oracle):
    """Ties everywhere: homopolymers, unit AC, U = 1, 2, 499, the j == 1 deletion quirk, gain 5."""
    rng = np.random.default_rng(5)
    reads = [np.zeros(600, np.int8), np.tile(np.array([0, 1], np.int8), 400), np.tile(np.array([0, 0, 1], np.int8), 300),
             rng.integers(0, 4, 2000).astype(np.int8), np.tile(rng.integers(0, 4, 499).astype(np.int8), 4)]
    tails = [(0, 0), (1, 0), (3, 3), (2, 1), (0, 3)]
    units = [np.array([0], np.uint8), np.array([0, 1], np.uint8), np.array([1, 0], np.uint8), np.array([0, 0], np.uint8),
             np.array([0, 0, 1], np.uint8), np.array([0, 1, 0, 1], np.uint8), reads[4][:499].astype(np.uint8),
             rng.integers(0, 4, 33).astype(np.uint8), rng.integers(0, 4, 17).astype(np.uint8), reads[3][100:165].astype(np.uint8)]
    jobs = []
    for r, rd in enumerate(reads):
        for u in units:
            for (g, mm, ind) in wdp_cases.PARAM_SETS:
                for first, rows in ((-1, len(rd)), (0, len(rd)), (7, min(300, len(rd) - 7)), (len(rd) - 4, 5)):
                    if len(u) > rows:
                        continue
                    jobs.append(dict(read=r, first=first, rows=rows, unit=u, gain=g, mis=mm, indel=ind))
    arr, res, aux = _run(gpu_ctx, reads, tails, jobs)
    bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res)
    assert not bad, bad[:10]


def test_harvested_pipeline_jobs(gpu_ctx, oracle):
    """Every DP the reference pipeline executes on a few synthetic reads, replayed on the GPU."""
    reads = []
    for ulen, copies, seed in ((5, 30, 1), (23, 12, 2), (100, 10, 3), (160, 8, 4)):
        rd, _ = synth.rand_seq_reads(ulen, copies, 0.016, 0.09, 0.038, 150, 200, 1, seed=seed)
        reads += rd
    o2 = oracle_lib.Oracle()
    hj, tails = o2.harvest_dp_jobs(reads)
    o2.close()
    assert len(hj) > 100
    for kind, mode in ((0, capi.TB_COUNTS), (1, capi.TB_CONSENSUS)):
        jobs = [j for j in hj if j["kind"] == kind]
        arr, res, aux = _run(gpu_ctx, reads, tails, jobs, mode=mode)
        for i, j in enumerate(jobs):           # the harvested results are the pipeline's own
            for f in ("best", "max_i", "max_j", "end_i", "n_match", "n_mismatch", "n_ins", "n_del", "n_scanned"):
                assert int(res[i, 0][f]) == j["result"][f], (kind, i, f, res[i, 0], j["result"])
        bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res, aux=aux, arr=arr, mode=mode)
        assert not bad, bad[:10]


def test_long_job_int32_scores(gpu_ctx, oracle):
    """Scores beyond the int16 range: 12 000 rows at gain 5, and a 40 000-row job."""
    rd, _ = synth.rand_seq_reads(50, 800, 0.01, 0.02, 0.02, 100, 100, 1, seed=9)
    reads, tails = [rd[0]], [(1, 2)]
    unit = np.array(rd[0][100:150], dtype=np.uint8)
    L = len(rd[0])
    jobs = [dict(read=0, first=50, rows=12000, unit=unit, gain=5, mis=1, indel=1),
            dict(read=0, first=-1, rows=L + 1, unit=unit, gain=1, mis=1, indel=3),
            dict(read=0, first=10, rows=L - 20, unit=unit, gain=1, mis=3, indel=1)]
    arr, res, aux = _run(gpu_ctx, reads, tails, jobs)
    bad = wdp_cases.check_against_oracle(oracle, reads, tails, jobs, res)
    assert not bad, bad
    assert res[0, 0]["best"] > 32767


def test_rejects_malformed_jobs(gpu_ctx):
    reads = [np.zeros(100, np.int8)]
    packed, woff, lens = capi.pack_reads(reads)
    gpu_ctx.upload_reads(packed, woff, lens)
    for bad in (dict(ulen=0), dict(ulen=500), dict(rows=200), dict(read=3), dict(first=-2)):
        j = dict(read=0, first=0, rows=50, unit=np.zeros(4, np.uint8), gain=1, mis=1, indel=3)
        arr, units, _ = wdp_cases.build_job_array([j])
        for k, v in bad.items():
            arr[0][k] = v
        with pytest.raises(capi.MtrError):
            gpu_ctx.wdp_run(arr, np.zeros(600, np.uint8))
