"""Definition of the golden cases shared by tests/golden/make_golden.py (run once, in the build container,
against the reference binary) and by the tests (which regenerate the same inputs from the same seeds)."""
from __future__ import annotations

import os
import tarfile

import numpy as np

from mtr_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
MODES = {"default": [], "a": ["-a"], "p_m07": ["-p", "-m", "0.7"]}


def synthetic_cases():
    """name -> (reads, line_width).  Seeds are fixed; the generators are numpy-only and deterministic."""
    cases = {}
    # mixed lengths in shuffled order: exercises the stale-buffer hazards H3/H4 across reads
    reads = []
    for i, (ul, cp) in enumerate([(100, 10), (10, 20), (50, 10), (2, 10), (20, 30), (200, 10), (5, 40), (7, 100)]):
        r, _ = synth.rand_seq_reads(ul, cp, 0.016, 0.09, 0.038, ul * cp // 2 + 37 * i, ul * cp // 3 + 11, 3, seed=100 + i)
        reads += r
    order = np.random.default_rng(5).permutation(len(reads))
    cases["mixed"] = ([reads[i] for i in order], 0)
    cases["long4"] = (synth.long_reads(4, seed=7)[0], 70)
    # BASELINE.json configs[0]: the 10_20_0_5_5_100_100_10 rand_seq shape
    cases["shape_10_20_0_5_5_100_100_10"] = (synth.rand_seq_reads(10, 20, 0.0, 0.05, 0.05, 100, 100, 10, seed=1)[0], 0)
    # test_single_TR/test.sh shapes: unit length i, 10 copies, flanks i*10, errors 1.6 / 9.0 / 3.8 %
    for ul in (2, 5, 10, 20, 50, 100, 200):
        cases["single_TR_%d" % ul] = (synth.rand_seq_reads(ul, 10, 0.016, 0.09, 0.038, ul * 10, ul * 10, 12, seed=200 + ul)[0], 0)
    cases["nanopore_100_100"] = (synth.standin_reads("nanopore", 3, seed=31)[0], 0)
    cases["pacbio_200_200"] = (synth.standin_reads("pacbio", 2, seed=32)[0], 0)
    return cases


def write_case(path, reads, line_width):
    synth.write_fasta(path, reads, line_width=line_width)


def extract_shipped(dst_dir):
    """Unpacks the 15 FASTA files of the reference's test_multiple_TRs/data (committed as a fixture)."""
    with tarfile.open(os.path.join(GOLDEN, "shipped_multiple_TRs.tar.gz")) as t:
        t.extractall(dst_dir)
    return sorted(f for f in os.listdir(dst_dir) if f.endswith(".fasta"))
