"""K1/K2 parity: CUDA directional index (through the C ABI) vs the CPU oracle; DI is compared bit-for-bit
as fp64 (Manhattan and Pearson), END and W as integers, including the stale-buffer hazard H3."""
import numpy as np
import pytest

import oracle_lib
from mtr_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _reads():
    reads = []
    for i, (ul, cp) in enumerate([(100, 10), (10, 20), (50, 10), (2, 10), (20, 30), (200, 10), (5, 40), (7, 100), (33, 300)]):
        r, _ = synth.rand_seq_reads(ul, cp, 0.016, 0.09, 0.038, ul * cp // 2 + 37 * i, ul * cp // 3 + 11, 2, seed=100 + i)
        reads += r
    rng = np.random.default_rng(5)
    reads = [reads[i] for i in rng.permutation(len(reads))]
    reads.append(rng.integers(0, 4, 8).astype(np.int8))          # no pass at all (w < L/2 never holds)
    reads.append(rng.integers(0, 4, 12).astype(np.int8))
    reads.append(np.zeros(400, np.int8))                           # homopolymer: sd == 0 in Pearson
    reads += synth.long_reads(2, seed=3)[0]
    reads.append(rng.integers(0, 4, 23000).astype(np.int8))       # all 20 passes
    return reads


def _oracle_expect(reads, manhattan):
    """Runs the oracle read by read (state persists, as in the reference) and records, per read, the stale
    tail the reference would find beyond the area it re-initialises, and the expected DI / END / W."""
    o = oracle_lib.Oracle(manhattan=manhattan)
    exp, stale = [], []
    for rd in reads:
        L = len(rd)
        r = 100 if L < 1000 else L // 10
        o.load_read(rd)
        M = L + r + 2 * 10240 + 64
        full = o.padded_codes(5, M)
        stale.append(full[L + 4 * r:].astype(np.uint16))
        exp.append(o.directional_index())
    o.close()
    return exp, stale


@pytest.mark.parametrize("manhattan", [True, False])
def test_directional_index_matches_oracle(gpu_ctx, manhattan):
    reads = _reads()
    exp, stale = _oracle_expect(reads, manhattan)
    packed, woff, lens = capi.pack_reads(reads)
    gpu_ctx.upload_reads(packed, woff, lens)
    stale_off = np.zeros(len(reads) + 1, dtype=np.int64)
    stale_off[1:] = np.cumsum([len(s) for s in stale])
    pos_off, di, end, w = gpu_ctx.di_run(manhattan, np.concatenate(stale), stale_off)
    for r, rd in enumerate(reads):
        sl = slice(pos_off[r], pos_off[r + 1])
        edi, eend, ew = exp[r]
        assert np.array_equal(end[sl], eend), (r, len(rd), np.flatnonzero(end[sl] != eend)[:5])
        assert np.array_equal(w[sl], ew), (r, len(rd))
        assert np.array_equal(di[sl].view(np.int64), edi.view(np.int64)), (r, len(rd), "DI differs bitwise")
    assert sum(int((e[1] > -1).sum()) for e in exp) > 100          # the case is not vacuous


def test_fresh_process_without_stale(gpu_ctx):
    """A single read in a fresh process: stale == NULL must behave as zeros."""
    reads = synth.long_reads(1, seed=11)[0]
    o = oracle_lib.Oracle()
    o.load_read(reads[0])
    edi, eend, ew = o.directional_index()
    o.close()
    packed, woff, lens = capi.pack_reads(reads)
    gpu_ctx.upload_reads(packed, woff, lens)
    pos_off, di, end, w = gpu_ctx.di_run(True)
    assert np.array_equal(end, eend) and np.array_equal(w, ew) and np.array_equal(di.view(np.int64), edi.view(np.int64))
