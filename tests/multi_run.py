"""Runs many (file, mode) jobs through handle_one_file() of ONE library in ONE process (tests/multi_run.py as a script is the
child): one CUDA start-up for a whole family of digest tests instead of one per command line, and a check that consecutive
files in one process do not influence each other (the reference allocates its state anew per file, handle_one_file.c:71-136).

    outputs = run_many(library_path, [(path, ["-p", "-m", "0.7"]), ...], env={...})     # -> list of stdout bytes, one per job
"""
import ctypes as C
import json
import os
import subprocess
import sys

SEP = b"@@MTR_JOB_END@@\n"


def run_many(lib_path, jobs, env=None, timeout=1800):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.abspath(__file__), lib_path], input=json.dumps([[path, list(flags)] for path, flags in jobs]).encode(),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e, timeout=timeout)
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    parts = p.stdout.split(SEP)
    assert len(parts) == len(jobs) + 1 and parts[-1] == b"", (len(parts), len(jobs))
    return parts[:-1]


def _child(lib_path):
    jobs = json.loads(sys.stdin.read())
    lib = C.CDLL(lib_path)
    libc = C.CDLL(None)
    lib.handle_one_file.argtypes = [C.c_char_p, C.c_int]
    lib.handle_one_file.restype = C.c_int
    manhattan = C.c_int.in_dll(lib, "Manhattan_Distance")
    ratio = C.c_float.in_dll(lib, "min_match_ratio")
    for path, flags in jobs:
        manhattan.value, ratio.value, print_alignment = 1, 0.6, 0          # main.c:47-49
        i = 0
        while i < len(flags):                                             # main.c:50-83, getopt "acm:p"
            if flags[i] == "-a":
                print_alignment = 1
            elif flags[i] == "-p":
                manhattan.value = 0
            elif flags[i] == "-m":
                i += 1
                ratio.value = float(flags[i])
            else:
                raise SystemExit("multi_run: unknown flag %r" % flags[i])
            i += 1
        lib.handle_one_file(path.encode(), print_alignment)
        libc.fflush(None)
        os.write(1, SEP)


if __name__ == "__main__":
    _child(sys.argv[1])
