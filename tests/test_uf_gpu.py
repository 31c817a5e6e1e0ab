"""K4 parity: the CUDA unit finder (counts, maximum-frequency node list, greedy de Bruijn walks) vs the oracle."""
import numpy as np
import pytest

import oracle_lib
from mtr_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _cases():
    reads = []
    for ul, cp, seed in ((5, 40, 1), (23, 15, 2), (100, 12, 3), (2, 30, 4), (37, 60, 5), (320, 8, 6)):
        reads += synth.rand_seq_reads(ul, cp, 0.02, 0.06, 0.04, 200, 200, 1, seed=seed)[0]
    reads.append(np.zeros(300, np.int8))                              # homopolymer: ties everywhere
    reads.append(np.tile(np.array([0, 1], np.int8), 200))
    reads += synth.long_reads(1, seed=9)[0]
    rng = np.random.default_rng(7)
    tasks = []
    for r, rd in enumerate(reads):
        L = len(rd)
        for _ in range(40):
            qs = int(rng.integers(0, L - 20))
            qe = int(min(L - 1, qs + rng.integers(10, max(11, min(L - qs, 4000)))))
            for k in rng.choice(np.arange(2, 16), 4, replace=False):
                tasks.append((r, qs, qe, int(k)))
        tasks.append((r, 0, L - 1, 5)); tasks.append((r, L - 30, L - 1, 7)); tasks.append((r, 3, L - 1, 15))
    return reads, tasks


def test_unit_finder_matches_oracle(gpu_ctx):
    reads, tasks = _cases()
    packed, woff, lens = capi.pack_reads(reads)
    gpu_ctx.upload_reads(packed, woff, lens)
    arr = np.array(tasks, dtype=capi.UF_TASK_DTYPE)
    res, units, scores = gpu_ctx.uf_run(arr)
    o = oracle_lib.Oracle()
    n_found = 0
    for i, (r, qs, qe, k) in enumerate(tasks):
        o.load_read(reads[r])
        exp = o.unit_walks(qs, qe, k)
        got = res[i]
        assert int(got["max_freq"]) == exp["max_freq"], (i, tasks[i], got, exp)
        assert int(got["found_last"]) == exp["found_last"], (i, tasks[i], got, exp)
        for d in range(2):
            assert int(got["found"][d]) == exp["found"][d], (i, d, tasks[i], got, exp)
            if exp["found"][d]:
                n_found += 1
                p = exp["period"][d]
                assert int(got["period"][d]) == p, (i, d, tasks[i])
                off = int(got["unit_off"][d])
                assert np.array_equal(units[off:off + p], exp["units"][d]), (i, d, tasks[i])
                assert np.array_equal(scores[off:off + p], np.minimum(exp["scores"][d], 254)), (i, d, tasks[i])   # counts are kept clamped to a byte
    o.close()
    assert n_found > 200
