"""Pins the CPU oracle: its output must reproduce the reference binary's stdout, byte for byte (md5), on
the reference's own deterministic inputs (15 files x 3 modes) and on the seeded synthetic
cases whose digests tests/golden/make_golden.py took from oracle/_ref/mTR_ref_det."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import golden_cases
import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_BIN = os.path.join(ROOT, "oracle", "mtr_oracle")
DIGESTS = json.load(open(os.path.join(golden_cases.GOLDEN, "digests.json")))

# SURVEY.md App. C, column "default" -- the survey's own measurement of the stock reference binary
SURVEY_APP_C_DEFAULT = {
    "3_5.fasta": "5b17b00a36c809f28b4aeb9d4a6199b3", "10_50.fasta": "b4ed5ed2b3bf06b8f0e5296e174c5381",
    "worm_chrI.fasta": "c69cc8326939f646bf2ead406f25dd30", "2_5_10_20_set.fasta": "9bdd2886b2ab2c13234e1b8e580f6b49",
}


def run_md5(binary, flags, path):
    out = subprocess.run([binary] + flags + [path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    return hashlib.md5(out).hexdigest()


@pytest.fixture(scope="module")
def shipped_dir(tmp_path_factory, oracle_so):
    d = tmp_path_factory.mktemp("shipped")
    golden_cases.extract_shipped(str(d))
    return str(d)


def test_digests_agree_with_survey():
    for f, md5 in SURVEY_APP_C_DEFAULT.items():
        assert DIGESTS["shipped"][f]["default"]["md5"] == md5


@pytest.mark.parametrize("name", sorted(DIGESTS["shipped"]))
def test_oracle_reproduces_reference_on_shipped_files(shipped_dir, name):
    for mode, flags in golden_cases.MODES.items():
        assert run_md5(ORACLE_BIN, flags, os.path.join(shipped_dir, name)) == DIGESTS["shipped"][name][mode]["md5"], (name, mode)


@pytest.fixture(scope="module")
def synthetic_dir(tmp_path_factory, oracle_so):
    d = tmp_path_factory.mktemp("synthetic")
    for name, (reads, lw) in golden_cases.synthetic_cases().items():
        golden_cases.write_case(os.path.join(str(d), name + ".fa"), reads, lw)
    return str(d)


@pytest.mark.parametrize("name", sorted(DIGESTS["synthetic"]))
def test_oracle_reproduces_reference_on_synthetic_cases(synthetic_dir, name):
    for mode, flags in golden_cases.MODES.items():
        assert run_md5(ORACLE_BIN, flags, os.path.join(synthetic_dir, name + ".fa")) == DIGESTS["synthetic"][name][mode]["md5"], (name, mode)


def test_min_missing_table_matches_reference_fixture(oracle_so):
    table = np.load(os.path.join(golden_cases.GOLDEN, "min_missing_table.npy"))
    L = oracle_lib.lib()
    got = np.array([[[L.mtro_min_missing_raw(i, j, k) for k in range(20)] for j in range(10)] for i in range(10)])
    assert np.array_equal(got, table)
    assert L.mtro_min_missing(201, 0.3, 25) == table[0, 0, 19]
    assert L.mtro_min_missing(5, 0.01, 1) == table[9, 9, 0]
    assert L.mtro_min_missing(100, 0.2, 7) == table[3, 3, 6]       # boundaries are strict ">"


def test_reference_library_dp_matches_oracle_dp(oracle_so):
    """When oracle/_ref/libmtr_ref.so exists (built from the unmodified reference), call its own
    wrap_around_DP_sub on random windows and compare the record with the oracle's DP."""
    import ctypes as C
    so = os.path.join(ROOT, "oracle", "_ref", "libmtr_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    ref = C.CDLL(so, mode=os.RTLD_LOCAL | os.RTLD_DEEPBIND)     # it defines the same globals as the product library
    ref.malloc_global_variables()
    org = C.POINTER(C.c_int).in_dll(ref, "orgInputString")
    rr = (C.c_char * 8192)()                     # repeat_in_read, mTR.h:99-119 (6720 bytes)
    ints = C.cast(rr, C.POINTER(C.c_int))
    # field offsets in ints: ID 0 | readID 4096 bytes | inputLen.. start at int 1025
    base = 1 + 4096 // 4
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(3)
    from mtr_b200 import synth
    for trial in range(40):
        ulen = int(rng.integers(2, 60))
        rd = synth.rand_seq_reads(ulen, int(rng.integers(6, 20)), 0.03, 0.05, 0.05, 60, 60, 1, seed=int(rng.integers(1 << 30)))[0][0]
        L = len(rd)
        for i in range(L):
            org[i] = int(rd[i])
        org[L] = 0; org[L + 1] = 0
        qs = int(rng.integers(0, L // 3)); qe = int(rng.integers(2 * L // 3, L))
        unit = rd[60:60 + ulen].astype(np.int32)
        ref.clear_rr(rr)
        ints[base + 4] = ulen                                           # rep_period
        C.memmove(C.addressof(rr) + (base + 14) * 4, bytes(b"ACGT"[b] for b in unit) + b"\0", ulen + 1)   # string
        g, mm, ind = [(1, 1, 3), (1, 3, 1), (5, 1, 1)][trial % 3]
        ref.wrap_around_DP_sub(qs, qe, rr, g, mm, ind)
        x = np.concatenate([rd, [0, 0]])[qs + 1: qe + 2]
        exp = o.wrap_dp(x, unit, g, mm, ind)
        got = dict(rep_start=ints[base + 1], rep_end=ints[base + 2], repeat_len=ints[base + 3], units=ints[base + 5],
                   nm=ints[base + 6], nx=ints[base + 7], ni=ints[base + 8], nd=ints[base + 9])
        want = dict(rep_start=qs + exp["end_i"] + 1, rep_end=qs + exp["max_i"], repeat_len=exp["max_i"] - exp["end_i"],
                    units=exp["n_scanned"] // ulen, nm=exp["n_match"], nx=exp["n_mismatch"], ni=exp["n_ins"], nd=exp["n_del"])
        assert got == want, (trial, got, want)
    o.close()
