"""Helper of tests/test_bridge_*.py (run as a subprocess): drives the reference's C <-> C++ bridge (mTR.h:146-175) of ONE
library -- the reference itself (oracle/_ref/libmtr_ref.so) or the product (libmtr_b200.so / the simulated-device build) --
on a synthetic read, and leaves what the library printed on stdout.

    python bridge_driver.py <library.so> <ref|ours> <print_alignment> [nodp]      nodp: skip the section that needs the DP (a GPU)
"""
import ctypes as C
import sys

import numpy as np


def make_read():
    rng = np.random.default_rng(11)
    unit_a = rng.integers(0, 4, 7)
    unit_b = rng.integers(0, 4, 23)
    unit_c = rng.integers(0, 4, 150)

    def noisy(unit, copies, rate):
        out = []
        for b in np.tile(unit, copies):
            u = rng.random()
            if u < rate / 3:
                continue                              # deletion
            if u < 2 * rate / 3:
                out.append(int(rng.integers(0, 4)))   # insertion
            out.append(int((b + 1 + rng.integers(0, 3)) % 4) if rng.random() < rate / 3 else int(b))
        return out
    parts, marks = [], []
    pos = 0
    for unit, copies, rate in ((unit_a, 30, 0.06), (unit_b, 12, 0.1), (unit_c, 5, 0.12)):
        flank = [int(x) for x in rng.integers(0, 4, 180)]
        parts += flank
        pos += len(flank)
        rep = noisy(unit, copies, rate)
        marks.append((pos, pos + len(rep) - 1, unit))
        parts += rep
        pos += len(rep)
    parts += [int(x) for x in rng.integers(0, 4, 200)]
    return np.array(parts, dtype=np.int32), marks


def main():
    path, kind, print_alignment = sys.argv[1], sys.argv[2], int(sys.argv[3])
    nodp = len(sys.argv) > 4 and sys.argv[4] == "nodp"
    lib = C.CDLL(path)
    libc = C.CDLL(None)
    read, marks = make_read()
    n = len(read)
    if kind == "ref":
        lib.malloc_global_variables()
        org = C.cast(C.c_void_p.in_dll(lib, "orgInputString"), C.POINTER(C.c_int))
        keep = None
    else:
        keep = (C.c_int * (n + 64))()
        C.c_void_p.in_dll(lib, "orgInputString").value = C.addressof(keep)
        org = keep
    for i in range(n):
        org[i] = int(read[i])
    for i in range(n, n + 8):
        org[i] = 0
    lib.insert_an_alignment_into_set.argtypes = [C.c_char_p] + [C.c_int] * 14 + [C.c_char_p, C.POINTER(C.c_int)]
    lib.pretty_print_alignment.argtypes = [C.c_char_p] + [C.c_int] * 6
    lib.print_freq.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int]
    score = (C.c_int * 512)()

    def text(unit):
        return "".join("ACGT"[int(b)] for b in unit).encode()
    # the three repeats, a rotated and clipped rival of the first (fewer matches: loses its place in the chain), and a
    # short one inside the second; distinct scores and end points: no tie depends on the order of the set
    a, b, c = marks
    rows = [
        (b"read/1", n, a[0], a[1], a[1] - a[0] + 1, 7, 30, 190, 6, 5, 4, 5, 1, 1, 3, text(a[2])),
        (b"read/1", n, a[0] + 3, a[1] - 20, a[1] - a[0] - 22, 7, 26, 150, 9, 8, 7, 6, 1, 1, 1, text(np.roll(a[2], -3))),
        (b"read/1", n, b[0], b[1], b[1] - b[0] + 1, 23, 12, 240, 10, 9, 8, 7, 1, 1, 1, text(b[2])),
        (b"read/1", n, b[0] + 40, b[0] + 90, 51, 23, 2, 40, 4, 3, 2, 7, 1, 1, 1, text(b[2])),
        (b"read/1", n, c[0], c[1], c[1] - c[0] + 1, 150, 5, 600, 30, 20, 25, 9, 1, 1, 3, text(c[2])),
    ]
    for r in rows:
        lib.insert_an_alignment_into_set(*r, score)
    lib.chaining(print_alignment)
    libc.fflush(None)
    sys.stdout.write("-- second read: the set is empty again\n")
    sys.stdout.flush()
    lib.chaining(print_alignment)                       # empty set: prints nothing
    lib.insert_an_alignment_into_set(*rows[2], score)
    lib.chaining(print_alignment)
    libc.fflush(None)
    if not nodp:
        sys.stdout.write("-- pretty_print_alignment alone (gain 5, penalties 1 / 3)\n")
        sys.stdout.flush()
        lib.pretty_print_alignment(text(a[2]), 7, a[0], a[1], 5, 1, 3)
        libc.fflush(None)
    sys.stdout.write("-- print_freq\n")
    sys.stdout.flush()
    for k in (1, 2, 3, 5, 7, 9, 12):
        for m in marks:
            if k - 1 <= len(m[2]):
                lib.print_freq(m[0], m[1], len(m[2]), text(m[2]), n, k)
        lib.print_freq(n - 40, n - 1, 7, text(a[2]), n, k)          # window that runs into the read end (raw bases above L - k + 1)
    libc.fflush(None)


if __name__ == "__main__":
    main()
