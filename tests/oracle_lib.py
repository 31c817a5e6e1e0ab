"""ctypes binding of oracle/liboracle.so -- the checker used by tests/, smoke() and bench.py's CPU legs.
Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")


class DpResult(C.Structure):
    _fields_ = [("best", C.c_int), ("max_i", C.c_int), ("max_j", C.c_int), ("end_i", C.c_int), ("end_j", C.c_int),
                ("n_match", C.c_int), ("n_mismatch", C.c_int), ("n_ins", C.c_int), ("n_del", C.c_int),
                ("n_scanned", C.c_int), ("cells", C.c_longlong), ("path_len", C.c_int)]


class Rr(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("inputLen", "rep_start", "rep_end", "repeat_len", "rep_period", "n_units",
                                       "n_match", "n_mismatch", "n_ins", "n_del", "kmer", "gain", "mis_pen", "indel_pen")] + \
               [("unit", C.c_char * 1024), ("unit_score", C.c_int * 500)]


class WalkResult(C.Structure):
    _fields_ = [("max_freq", C.c_int), ("found_last", C.c_int), ("found", C.c_int * 2), ("period", C.c_int * 2)]


class OStats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("dp_calls", "dp_cells", "revise_calls", "revise_cells", "print_calls",
                                            "print_cells", "di_position_passes", "candidates", "searches", "reads", "bases")]


HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.c_int,
                   C.c_int, C.c_int, C.c_int, C.POINTER(DpResult))

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ODIR, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", ODIR, "oracle"])
        L = C.CDLL(so)
        L.mtro_new.restype = C.c_void_p
        L.mtro_new.argtypes = [C.c_int, C.c_float]
        L.mtro_free.argtypes = [C.c_void_p]
        L.mtro_process_file.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.mtro_process_file.restype = C.c_int
        L.mtro_load_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.mtro_process_read.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.mtro_directional_index.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mtro_padded_codes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.mtro_di_pass.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mtro_wrap_dp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.POINTER(DpResult), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mtro_find_tandem_repeat.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Rr)]
        L.mtro_unit_walks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(WalkResult), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mtro_min_missing.argtypes = [C.c_int, C.c_double, C.c_int]
        L.mtro_min_missing.restype = C.c_int
        L.mtro_min_missing_raw.argtypes = [C.c_int, C.c_int, C.c_int]
        L.mtro_min_missing_raw.restype = C.c_int
        L.mtro_get_stats.argtypes = [C.c_void_p, C.POINTER(OStats)]
        L.mtro_set_dp_hook.argtypes = [C.c_void_p, HOOK, C.c_void_p]
        L.mtro_org.argtypes = [C.c_void_p]
        L.mtro_org.restype = C.c_void_p
        L.mtro_set_output.argtypes = [C.c_void_p, C.c_void_p]
        L.mtro_trs_in_neighborhood.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.mtro_freq_2mer.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.mtro_trs_in_neighborhood.restype = C.c_int
        L.mtro_cmp_tr.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mtro_cmp_tr.restype = C.c_int
        _lib = L
    return _lib


_libc = C.CDLL(None)
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


class Oracle:
    def __init__(self, manhattan=True, min_match_ratio=0.6):
        self.L = lib()
        self.h = self.L.mtro_new(1 if manhattan else 0, min_match_ratio)
        self._hook = None
        self._fp = None

    def close(self):
        if self.h:
            self.L.mtro_free(self.h)
            self.h = None
        if self._fp:
            _libc.fclose(self._fp)
            self._fp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_output_path(self, path):
        if self._fp:
            _libc.fclose(self._fp)
        self._fp = _libc.fopen(path.encode(), b"w")
        self.L.mtro_set_output(self.h, self._fp)

    def finish_output(self):
        if self._fp:
            _libc.fclose(self._fp)
            self._fp = None

    def load_read(self, bases):
        b = np.ascontiguousarray(bases, dtype=np.int32)
        self.L.mtro_load_read(self.h, b.ctypes.data, len(b))
        self.cur_len = len(b)

    def process_read(self, read_id, bases, print_alignment=0):
        b = np.ascontiguousarray(bases, dtype=np.int32)
        self.L.mtro_process_read(self.h, str(read_id).encode(), b.ctypes.data, len(b), print_alignment)

    def process_file(self, path, print_alignment=0):
        return self.L.mtro_process_file(self.h, path.encode(), print_alignment)

    def directional_index(self):
        n = self.cur_len
        di = np.zeros(n, dtype=np.float64); end = np.zeros(n, dtype=np.int32); w = np.zeros(n, dtype=np.int32)
        self.L.mtro_directional_index(self.h, di.ctypes.data, end.ctypes.data, w.ctypes.data)
        return di, end, w

    def padded_codes(self, k, n):
        out = np.zeros(n, dtype=np.int32)
        self.L.mtro_padded_codes(self.h, k, out.ctypes.data, n)
        return out

    def di_pass(self, k, w):
        L = self.cur_len
        r = 100 if L < 1000 else L // 10
        tmp = np.zeros(L + 2 * r, dtype=np.float64)
        self.L.mtro_di_pass(self.h, k, w, tmp.ctypes.data)
        return tmp

    def wrap_dp(self, x, u, gain, mis, indel, mode=0, want_dirs=False):
        """x: read bases of rows 1..R (array length R), u: unit bases (length U).  Returns dict."""
        R, U = len(x), len(u)
        xa = np.zeros(R + 1, dtype=np.int32); xa[1:] = x
        ua = np.zeros(U + 1, dtype=np.int32); ua[1:] = u
        res = DpResult()
        cons = np.zeros((U + 1) * 5, dtype=np.int32) if mode == 1 else None
        miss = np.zeros((U + 1) * 4, dtype=np.int32) if mode == 1 else None
        path = np.zeros(R * 2 + R * U // 4 + 1024, dtype=np.uint8) if mode == 2 else None
        dirs = np.zeros((R + 1) * (U + 1), dtype=np.uint8) if want_dirs else None
        p = lambda a: None if a is None else a.ctypes.data
        self.L.mtro_wrap_dp(self.h, xa.ctypes.data, R, ua.ctypes.data, U, gain, mis, indel, mode, C.byref(res),
                            p(cons), p(miss), p(path), p(dirs))
        out = {n: getattr(res, n) for n, _ in DpResult._fields_}
        if mode == 1:
            out["consensus"] = cons.reshape(U + 1, 5); out["missing"] = miss.reshape(U + 1, 4)
        if mode == 2:
            out["path"] = path[:res.path_len].copy()
        if want_dirs:
            out["dirs"] = dirs.reshape(R + 1, U + 1)
        return out

    def unit_walks(self, qs, qe, k):
        res = WalkResult()
        uf = np.zeros(500, np.uint8); sf = np.zeros(500, np.int32); ub = np.zeros(500, np.uint8); sb = np.zeros(500, np.int32)
        self.L.mtro_unit_walks(self.h, qs, qe, k, C.byref(res), uf.ctypes.data, sf.ctypes.data, ub.ctypes.data, sb.ctypes.data)
        out = dict(max_freq=res.max_freq, found_last=res.found_last, found=(res.found[0], res.found[1]),
                   period=(res.period[0], res.period[1]))
        out["units"] = (uf[:res.period[0]].copy(), ub[:res.period[1]].copy())
        out["scores"] = (sf[:res.period[0]].copy(), sb[:res.period[1]].copy())
        return out

    def stats(self):
        s = OStats()
        self.L.mtro_get_stats(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in OStats._fields_}

    def harvest_dp_jobs(self, reads, print_alignment=0, max_jobs=None):
        """Runs the whole pipeline on `reads` and records every DP the reference would execute.
        Returns a list of dicts: read, first, rows, unit (uint8 array), gain, mis, indel, kind, result."""
        jobs = []
        org = self.L.mtro_org(self.h)
        cur = {"read": 0}

        def hook(user, kind, x, rows, u, ulen, g, mm, ind, res):
            if max_jobs is not None and len(jobs) >= max_jobs:
                return
            first = (C.addressof(x.contents) - org) // 4
            unit = np.array([u[j] for j in range(1, ulen + 1)], dtype=np.uint8)
            r = res.contents
            jobs.append(dict(read=cur["read"], first=first, rows=rows, unit=unit, gain=g, mis=mm, indel=ind, kind=kind,
                             result={n: getattr(r, n) for n, _ in DpResult._fields_}))

        self._hook = HOOK(hook)
        self.L.mtro_set_dp_hook(self.h, self._hook, None)
        devnull = _libc.fopen(b"/dev/null", b"w")
        self.L.mtro_set_output(self.h, devnull)
        tails = []
        prev = np.zeros(2 + max(len(r) for r in reads) + 2, dtype=np.int32)   # shadow of orgInputString
        for i, rd in enumerate(reads):
            cur["read"] = i
            L = len(rd)
            tails.append((int(prev[L]), int(prev[L + 1])))
            prev[:L] = rd
            self.process_read(i, rd, print_alignment)
        _libc.fclose(devnull)
        self.L.mtro_set_dp_hook(self.h, HOOK(0), None)
        return jobs, tails
