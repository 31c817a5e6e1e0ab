"""the optional cross-read clustering post-pass (mtr_cluster_records, mtr_b200/csrc/cluster.cpp) against
a literal Python transcription of the reference's k_means_clustering.c:136-355 built on the oracle's restated primitives
(mtro_freq_2mer, mtro_cmp_tr, mtro_trs_in_neighborhood -- row A7).  The reference file is dead code that does not compile
(PARITY UNPINNED, see cluster.cpp): the transcription below follows the source line by line, with the constants the source
leaves undefined set to the post-pass's defaults and Python's stable sort in place of the random-pivot quicksort."""
import ctypes as C
import functools
import os
import subprocess

import numpy as np

import oracle_lib
from mtr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MIN_MATCH_RATIO, MH, MIN_REP_LEN, MIN_NUM_repTR = 0.6, 0.3, 0, 1


def f2(unit):
    u = np.array(["ACGT".index(c) for c in unit], dtype=np.int32)
    out = np.zeros(16, dtype=np.int32)
    oracle_lib.lib().mtro_freq_2mer(u.ctypes.data_as(C.c_void_p), len(u), out.ctypes.data_as(C.c_void_p))
    return out


def cmp_TR(a, b, mode):
    ra, rb = a.get("repTR") or {}, b.get("repTR") or {}
    return oracle_lib.lib().mtro_cmp_tr(a["period"], a["f2"].ctypes.data_as(C.c_void_p), a["units"], ra.get("freq", 0), ra.get("ID", 0),
                                        b["period"], b["f2"].ctypes.data_as(C.c_void_p), b["units"], rb.get("freq", 0), rb.get("ID", 0), mode)


def near(a, rep):
    return oracle_lib.lib().mtro_trs_in_neighborhood(a["f2"].ctypes.data_as(C.c_void_p), rep["f2"].ctypes.data_as(C.c_void_p), rep["period"])


def reference_clustering(lines):
    """k_means_clustering (:251-355) over the records; returns the printed lines."""
    rr = []
    for ID, line in enumerate(lines):
        f = line.split("\t")
        rr.append({"ID": ID, "fields": f, "period": len(f[12]), "units": int(f[6]), "matches": int(f[7]), "repeat_len": int(f[4]), "string": f[12]})
    TR_list = []
    for a in rr:                                                     # :262-291
        ratio = np.float32(a["matches"]) / np.float32(a["repeat_len"])
        if MIN_REP_LEN < a["period"] * a["units"] and MIN_MATCH_RATIO < float(ratio) and 1 < a["units"]:
            TR_list.append({"ID": a["ID"], "repID": a["ID"], "repTR": None, "freq": 1, "period": a["period"], "f2": f2(a["string"]), "units": a["units"]})
    n = len(TR_list)
    TR_list.sort(key=functools.cmp_to_key(lambda x, y: cmp_TR(x, y, 1)))     # :303
    repTR_list = []                                                  # select_repTR_list, :136-167
    freq, i = 1, 0
    for i in range(n - 1):
        if cmp_TR(TR_list[i], TR_list[i + 1], 0) == 0:
            freq += 1
        elif MIN_NUM_repTR <= freq:
            rep = dict(TR_list[i]); rep["freq"] = freq
            repTR_list.append(rep)
            for k in range(i, i - freq, -1):
                TR_list[k]["repID"] = rep["ID"]; TR_list[k]["repTR"] = rep
            freq = 1
    i = n - 1
    if n > 0 and MIN_NUM_repTR <= freq:
        rep = dict(TR_list[i]); rep["freq"] = freq
        repTR_list.append(rep)
        for k in range(i, i - freq, -1):
            TR_list[k]["repID"] = rep["ID"]; TR_list[k]["repTR"] = rep
    m = len(repTR_list)
    for i in range(m):                                               # revise_repTR_list, :182-233
        a = repTR_list[i]
        lb = a["period"] - int(a["period"] * 0.1); ub = a["period"] + int(a["period"] * 0.1)
        max_freq, max_i = a["freq"], i
        for j in range(i - 1, -1, -1):
            rep = repTR_list[j]
            if lb <= rep["period"]:
                if near(a, rep) == 1 and max_freq < rep["freq"]:
                    max_freq, max_i = rep["freq"], j
            else:
                break
        for j in range(i + 1, m):
            rep = repTR_list[j]
            if rep["period"] <= ub:
                if near(a, rep) == 1 and max_freq < rep["freq"]:
                    max_freq, max_i = rep["freq"], j
            else:
                break
        repTR_list[i]["repID"] = repTR_list[max_i]["ID"]; repTR_list[i]["repTR"] = repTR_list[max_i]
    for i in range(m):
        a = repTR_list[i]
        if a["ID"] != a["repID"]:
            while a["ID"] != a["repID"]:
                a = a["repTR"]
            repTR_list[i]["repID"] = a["ID"]; repTR_list[i]["repTR"] = a
            a["freq"] += repTR_list[i]["freq"]
    for t in TR_list:                                                # update_repTRs_in_TR_list, :236-249
        a = t["repTR"]
        if a["ID"] != a["repID"]:
            while a["ID"] != a["repID"]:
                a = a["repTR"]
            t["repID"] = a["ID"]; t["freq"] = a["freq"]; t["repTR"] = a
    TR_list.sort(key=functools.cmp_to_key(lambda x, y: cmp_TR(x, y, 2)))     # :337
    out = []
    for t in TR_list:                                                # print_one_TR_with_read, :11-25 (not pretty)
        f = list(rr[t["ID"]]["fields"])
        f[5] = str(t["repTR"]["period"]); f[12] = rr[t["repTR"]["ID"]]["string"]
        out.append("%d\t%s" % (t["repID"], "\t".join(f)))
    return out


def cluster(text, params=None):
    lib = capi.load_library()
    lib.mtr_cluster_records.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.mtr_cluster_free.argtypes = [C.c_void_p]
    out, n = C.c_void_p(), C.c_int64()
    rc = lib.mtr_cluster_records(text, len(text), params, C.byref(out), C.byref(n))
    assert rc == 0
    got = C.string_at(out, n.value)
    lib.mtr_cluster_free(out)
    return got


def record(rid, unit, units, ratio, rng):
    replen = len(unit) * units + int(rng.integers(0, 3))
    m = int(round(replen * ratio))
    start = int(rng.integers(1, 500))
    return "%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%f\t%d\t%d\t%d\t%s" % (rid, 5000, start, start + replen - 1, replen, len(unit), units, m,
                                                            np.float32(m) / np.float32(replen), replen - m, 1, 2, unit)


def synthetic_records(seed, families=14):
    rng = np.random.default_rng(seed)
    lines = []
    for fam in range(families):
        L = int(rng.choice([3, 5, 7, 12, 20, 31, 48, 60, 90]))
        base = "".join("ACGT"[x] for x in rng.integers(0, 4, L))
        variants = [base, base[2:] + base[:2], base[L // 2:] + base[:L // 2]]              # rotations: the same 2-mer vector
        for _ in range(int(rng.integers(1, 5))):                                            # substitutions: near neighbours
            v = list(base)
            for p in rng.integers(0, L, int(rng.integers(1, 3))):
                v[p] = "ACGT"[int(rng.integers(0, 4))]
            variants.append("".join(v))
        if L >= 12:
            variants.append(base + base[0])                                                 # one base longer: within 10 % from L = 10 on
            variants.append(base[:-1])
        for v in variants:
            for _ in range(int(rng.integers(1, 6))):
                ratio = float(rng.choice([0.55, 0.6, 0.7, 0.85, 0.95, 1.0]))
                units = int(rng.choice([1, 2, 3, 8, 20, 40]))
                lines.append(record("read%d/%d" % (fam, len(lines)), v, units, ratio, rng))
    order = rng.permutation(len(lines))
    return [lines[i] for i in order]


def test_hand_worked_case():
    # three copies of ACG-like units (two rotations of one unit and a one-substitution neighbour), one unrelated unit of the
    # same length class and one record that does not qualify (a single unit)
    rng = np.random.default_rng(0)
    lines = [record("a", "ACGTT", 10, 0.9, rng), record("b", "GTTAC", 12, 0.9, rng), record("c", "ACGTA", 9, 0.9, rng),
             record("d", "CCCCC", 30, 0.95, rng), record("e", "ACGTT", 1, 0.9, rng)]
    got = cluster(("\n".join(lines) + "\n").encode()).decode().splitlines()
    assert got == reference_clustering(lines)
    ids = [g.split("\t")[0] for g in got]
    units = [g.split("\t")[-1] for g in got]
    # a and b share <5, 2-mer vector>: b (more units) represents them; c (AC CG GT TA AA vs AC CG GT TT TA: distance 2 > 0.3 * 5)
    # stays alone, as does d; e is not printed
    assert len(got) == 4 and ids[:2] == ["1", "1"] and units[:2] == ["GTTAC", "GTTAC"] and "e" not in [g.split("\t")[1] for g in got]


def test_synthetic_families_match_the_transcription():
    for seed in (1, 2, 3):
        lines = synthetic_records(seed)
        text = ("\n".join(lines) + "\n").encode()
        got = cluster(text).decode().splitlines()
        want = reference_clustering(lines)
        assert got == want, seed
        assert len(set(g.split("\t")[0] for g in got)) < len(got)        # something was merged


def test_other_lines_are_skipped_and_the_cli_agrees():
    lines = synthetic_records(5, families=4)
    noisy = []
    for i, l in enumerate(lines):
        noisy.append(l)
        if i % 3 == 0:
            noisy += ["", "match gain = 1, mismatch penalty = 1, indel penalty = 3", "", "ACGT-ACGT", "|||| ||||", "ACGTTACGT", ""]
    text = ("\n".join(noisy) + "\n").encode()
    want = ("\n".join(reference_clustering(lines)) + "\n").encode()
    assert cluster(text) == want
    p = subprocess.run([os.path.join(ROOT, "bin", "mTR_cluster")], input=text, stdout=subprocess.PIPE, check=True)
    assert p.stdout == want
    assert cluster(b"") == b""
