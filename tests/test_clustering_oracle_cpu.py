"""the 2-mer frequency vector (handle_one_read.c:63-72) and the distance / ordering helpers of
the reference's dead cross-read clustering (k_means_clustering.c:62-101 cmp_TR, :169-180 TRs_in_neighborhood) as
restated in the oracle.  k_means_clustering.c is not part of the reference's build (MH_distance_threshold is never
defined), so there is no binary to compare with: the expectations below are worked out by hand from the source."""
import ctypes as C

import numpy as np

import oracle_lib


def f2(unit):
    L = oracle_lib.lib()
    u = np.asarray(unit, dtype=np.int32)
    out = np.zeros(16, dtype=np.int32)
    L.mtro_freq_2mer(u.ctypes.data_as(C.c_void_p), len(u), out.ctypes.data_as(C.c_void_p))
    return out


def near(a, rep, rep_period):
    L = oracle_lib.lib()
    a = np.asarray(a, dtype=np.int32); rep = np.asarray(rep, dtype=np.int32)
    return L.mtro_trs_in_neighborhood(a.ctypes.data_as(C.c_void_p), rep.ctypes.data_as(C.c_void_p), rep_period)


def cmp_tr(a, b, mode):
    L = oracle_lib.lib()
    fa = np.asarray(a["f2"], dtype=np.int32); fb = np.asarray(b["f2"], dtype=np.int32)
    return L.mtro_cmp_tr(a["period"], fa.ctypes.data_as(C.c_void_p), a["units"], a.get("rep_freq", 0), a.get("rep_id", 0),
                         b["period"], fb.ctypes.data_as(C.c_void_p), b["units"], b.get("rep_freq", 0), b.get("rep_id", 0), mode)


def test_freq_2mer_is_circular():
    # ACGT (0 1 2 3): AC, CG, GT and the wrap-around TA
    v = f2([0, 1, 2, 3])
    want = np.zeros(16, dtype=np.int32)
    for a, b in ((0, 1), (1, 2), (2, 3), (3, 0)):
        want[a * 4 + b] += 1
    assert (v == want).all()
    # a rotation of the unit has the same circular 2-mer counts; a homopolymer of length n counts n times AA
    assert (f2([2, 3, 0, 1]) == v).all()
    assert f2([0] * 7)[0] == 7 and f2([0] * 7).sum() == 7
    assert f2([3]).tolist() == [0] * 15 + [1]


def test_trs_in_neighborhood_threshold():
    rep = f2([0, 1] * 10)                                       # period 20: AC x 10 -> AC 10, CA 10
    assert near(rep, rep, 20) == 1
    a = rep.copy(); a[1] -= 3; a[2] += 3                        # Manhattan distance 6 = 0.3 * 20: not beyond the threshold
    assert near(a, rep, 20) == 1
    a[3] += 1                                                   # 7 > 6
    assert near(a, rep, 20) == -1
    assert near(a, rep, 24) == 1                                # the threshold scales with the REPRESENTATIVE's period (7 <= 7.2)
    assert near(rep, a, 2) == -1


def test_cmp_tr_orders():
    x = {"period": 5, "f2": f2([0, 1, 2, 3, 0]), "units": 7, "rep_freq": 3, "rep_id": 11}
    y = {"period": 6, "f2": f2([0, 1, 2, 3, 0, 0]), "units": 2, "rep_freq": 9, "rep_id": 4}
    assert cmp_tr(x, y, 0) == -1 and cmp_tr(y, x, 0) == 1      # unit length first
    z = dict(x, f2=f2([0, 1, 2, 3, 1]), units=9)                # same length: first differing 2-mer count decides
    d = cmp_tr(x, z, 0)
    i = int(np.flatnonzero(x["f2"] != z["f2"])[0])
    assert d == int(x["f2"][i] - z["f2"][i]) != 0 and cmp_tr(z, x, 0) == -d
    w = dict(x, units=12)
    assert cmp_tr(x, w, 0) == 0                                 # mode 0 ignores the number of units ...
    assert cmp_tr(x, w, 1) == 7 - 12                            # ... mode 1 breaks the tie with it
    assert cmp_tr(x, y, 2) == -(3 - 9)                          # mode 2: frequency of the representative, descending ...
    assert cmp_tr(x, dict(y, rep_freq=3), 2) == 11 - 4          # ... then its identifier
    assert cmp_tr(x, y, 3) == 0
