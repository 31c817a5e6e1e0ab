"""The full-size parity cases (BASELINE.json configs at the sizes SURVEY.md 8(d) states), shared by
tests/golden/make_golden_full.py (reference side, run once in the build container) and tests/test_full_size_gpu.py.
A case is a list of CHUNKS (FASTA texts): the reference runs every chunk as a process of its own, the tests load them
one by one -- the cross-read stale state starts fresh with each."""
from __future__ import annotations

import numpy as np

from mtr_b200 import synth

N_CHUNKS = 8
CASES = {
    # C1: test_single_TR/test.sh shapes -- unit length i, 10 copies, flanks 10 i, errors 1.6 / 9.0 / 3.8 %, 1000 reads each
    **{"C1_single_TR_%d" % ul: {"flags": [], "reads": 1000, "ul": ul} for ul in (2, 5, 10, 20, 50, 100, 200)},
    # C2: PacBio_Nanopore_read/100_100_nanopore.fasta stand-in with -a
    "C2_nanopore_100_100_a": {"flags": ["-a"], "reads": 1000},
    # C3: PacBio_Nanopore_read/200_200_pacbio_1.fasta stand-in with -p -m 0.7
    "C3_pacbio_200_200_p_m07": {"flags": ["-p", "-m", "0.7"], "reads": 500},
    # C5: the bench workload (first 2048 reads of step 0 of rank 0)
    "C5_long_reads": {"flags": [], "reads": 2048},
}


def reads_of(name):
    c = CASES[name]
    if name.startswith("C1_"):
        ul = c["ul"]
        return synth.rand_seq_reads(ul, 10, 0.016, 0.09, 0.038, ul * 10, ul * 10, c["reads"], seed=300 + ul)[0]
    if name.startswith("C2_"):
        return synth.standin_reads("nanopore", c["reads"], seed=41)[0]
    if name.startswith("C3_"):
        return synth.standin_reads("pacbio", c["reads"], seed=42)[0]
    return synth.long_reads(c["reads"], seed=1000)[0]


def chunks(name):
    reads = reads_of(name)
    out = []
    for idx in np.array_split(np.arange(len(reads)), N_CHUNKS):
        out.append("".join(">%d\n%s\n" % (int(i), synth.to_text(reads[int(i)])) for i in idx).encode())
    return out
