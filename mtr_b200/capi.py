"""ctypes binding of include/mtr_b200.h (libmtr_b200.so) -- the Python mirror of the C ABI.

This module never computes anything itself and has no CPU fallback: if the shared library is missing it
raises, and if no CUDA device is usable ``Context()`` raises with the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmtr_b200.so")

TB_COUNTS, TB_CONSENSUS, TB_PATH = 0, 1, 2

# every symbol include/mtr_b200.h declares (tests check that the library exports all of them)
ABI_FUNCTIONS = [
    "mtr_cuda_init", "mtr_cuda_shutdown", "mtr_last_error", "mtr_device_count", "mtr_set_blocking_sync", "mtr_set_priority", "mtr_reads_upload", "mtr_reads_share",
    "mtr_wdp_run", "mtr_wdp_upload", "mtr_wdp_launch", "mtr_wdp_download", "mtr_wdp_set_fused_traceback", "mtr_di_run", "mtr_di_run_range", "mtr_get_stats",
    "mtr_alu_probe", "mtr_uf_run", "mtr_pipeline_open", "mtr_pipeline_close", "mtr_pipeline_load_fasta", "mtr_pipeline_load_fasta_shard",
    "mtr_pipeline_run", "mtr_pipeline_get_stats", "mtr_pipeline_ctx", "mtr_engine_run", "mtr_engine_run_range", "mtr_engine_set_speculate", "mtr_engine_dp_busy_ms", "handle_one_file", "handle_one_read", "mtr_flush", "mtr_file_stats", "mtr_cluster_records", "mtr_cluster_free",
    "insert_an_alignment_into_set", "chaining", "pretty_print_alignment", "print_freq",      # the reference's C <-> C++ bridge (mTR.h:146-175)
]
ABI_GLOBALS = [
    "Manhattan_Distance", "min_match_ratio", "orgInputString", "time_all", "time_memory", "time_range",
    "time_period", "time_initialize_input_string", "time_wrap_around_DP", "time_count_table", "time_chaining",
    "query_counter",
]


class WdpJob(C.Structure):
    _fields_ = [
        ("read", C.c_int32), ("first", C.c_int32), ("rows", C.c_int32), ("unit_off", C.c_int32),
        ("ulen", C.c_int32), ("gain", C.c_int8 * 2), ("mis", C.c_int8 * 2), ("indel", C.c_int8 * 2),
        ("n_param", C.c_uint8), ("mode", C.c_uint8), ("aux_off", C.c_int64), ("aux_cap", C.c_int64),
    ]


class WdpResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "best", "max_i", "max_j", "end_i", "end_j", "n_match", "n_mismatch", "n_ins", "n_del", "n_scanned",
        "path_len", "flags")]


class Stats(C.Structure):
    _fields_ = [
        ("wdp_fill_ms", C.c_double), ("wdp_tb_ms", C.c_double), ("di_ms", C.c_double),
        ("wdp_cells", C.c_int64), ("wdp_slot_cells", C.c_int64), ("wdp_dir_bytes", C.c_int64),
        ("di_position_passes", C.c_int64), ("di_bytes_in", C.c_int64), ("di_bytes_out", C.c_int64),
        ("launches", C.c_int32), ("n_sm", C.c_int32),
        ("uf_ms", C.c_double), ("uf_tasks", C.c_int64), ("uf_table_bytes", C.c_int64),
    ]


class PipelineStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "reads", "bases", "groups", "candidates", "waves", "jobs", "dp_tasks", "wdp_cells", "wdp_slot_cells", "wdp_dir_bytes",
        "spec_cells", "wdp_cells_p16", "shared_cells", "tables", "table_positions", "walks", "repeats", "h2d_bytes", "d2h_bytes", "launches")] + \
        [(n, C.c_double) for n in ("dp_ms", "di_kernel_ms", "uf_kernel_ms", "engine_wall_ms", "pack_ms", "chain_ms", "wall_ms")]


class EngineStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "waves", "candidates", "dp_jobs", "dp_tasks", "dp_cells", "dp_slot_cells", "dp_dir_bytes", "spec_cells", "dp_cells_p16", "shared_cells", "tables",
        "table_positions", "walks", "repeats", "wrapdp_messages", "launches", "h2d_bytes", "d2h_bytes")] + \
        [(n, C.c_double) for n in ("di_ms", "dp_ms", "uf_ms", "wall_ms")]


REPEAT_DTYPE = np.dtype([(n, "<i4") for n in (
    "read", "seq", "rep_start", "rep_end", "repeat_len", "rep_period", "num_freq_unit", "num_matches", "num_mismatches",
    "num_insertions", "num_deletions", "kmer", "match_gain", "mismatch_penalty", "indel_penalty", "pad_")] + [("unit_off", "<i8")])


UF_TASK_DTYPE = np.dtype([("read", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("k", "<i4")])
UF_RESULT_DTYPE = np.dtype([("max_freq", "<i4"), ("found_last", "<i4"), ("found", "<i4", (2,)), ("period", "<i4", (2,)),
                            ("unit_off", "<i8", (2,))])
JOB_DTYPE = np.dtype([
    ("read", "<i4"), ("first", "<i4"), ("rows", "<i4"), ("unit_off", "<i4"), ("ulen", "<i4"),
    ("gain", "i1", (2,)), ("mis", "i1", (2,)), ("indel", "i1", (2,)), ("n_param", "u1"), ("mode", "u1"),
    ("aux_off", "<i8"), ("aux_cap", "<i8")], align=True)
RESULT_DTYPE = np.dtype([(n, "<i4") for n in (
    "best", "max_i", "max_j", "end_i", "end_j", "n_match", "n_mismatch", "n_ins", "n_del", "n_scanned",
    "path_len", "flags")])
assert JOB_DTYPE.itemsize == C.sizeof(WdpJob), (JOB_DTYPE.itemsize, C.sizeof(WdpJob))
assert RESULT_DTYPE.itemsize == C.sizeof(WdpResult)

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Loads libmtr_b200.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmtr_b200.so is missing (%s): build it with __graft_entry__.build(); "
                           "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.mtr_cuda_init.argtypes = [C.c_int, C.POINTER(vp)]
    lib.mtr_cuda_init.restype = C.c_int
    lib.mtr_cuda_shutdown.argtypes = [vp]
    lib.mtr_cuda_shutdown.restype = None
    lib.mtr_last_error.argtypes = [vp]
    lib.mtr_last_error.restype = C.c_char_p
    lib.mtr_device_count.restype = C.c_int
    lib.mtr_reads_upload.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.mtr_wdp_run.argtypes = [vp, vp, C.c_int, vp, i64, vp, vp, i64]
    lib.mtr_wdp_upload.argtypes = [vp, vp, C.c_int, vp, i64, i64]
    lib.mtr_wdp_launch.argtypes = [vp]
    lib.mtr_wdp_download.argtypes = [vp, vp, vp, i64]
    lib.mtr_wdp_set_fused_traceback.argtypes = [vp, C.c_int]
    lib.mtr_set_priority.argtypes = [vp, C.c_int]
    lib.mtr_di_run.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.mtr_di_run_range.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
    lib.mtr_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.mtr_alu_probe.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    lib.mtr_uf_run.argtypes = [vp, vp, C.c_int, vp, vp, vp, i64, C.POINTER(i64)]
    lib.mtr_uf_run.restype = C.c_int
    lib.mtr_pipeline_open.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    lib.mtr_pipeline_close.argtypes = [vp]
    lib.mtr_pipeline_close.restype = None
    lib.mtr_pipeline_load_fasta.argtypes = [vp, C.c_char_p, i64]
    lib.mtr_pipeline_load_fasta_shard.argtypes = [vp, C.c_char_p, i64, C.c_int, C.c_int]
    lib.mtr_pipeline_load_fasta_shard.restype = C.c_int
    lib.mtr_pipeline_run.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(i64)]
    lib.mtr_pipeline_get_stats.argtypes = [vp, C.POINTER(PipelineStats)]
    lib.mtr_engine_run.argtypes = [vp, C.c_int, C.c_float, vp, vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(EngineStats)]
    lib.mtr_engine_run.restype = C.c_int
    lib.mtr_engine_set_speculate.argtypes = [vp, C.c_int]
    lib.mtr_engine_dp_busy_ms.argtypes = [C.c_int, C.c_int]
    lib.mtr_engine_dp_busy_ms.restype = C.c_double
    lib.mtr_pipeline_ctx.argtypes = [vp]
    lib.mtr_pipeline_ctx.restype = vp
    lib.mtr_file_stats.argtypes = [C.POINTER(PipelineStats)]
    lib.handle_one_file.argtypes = [C.c_char_p, C.c_int]
    lib.handle_one_file.restype = C.c_int
    lib.mtr_flush.restype = None
    for f in ("mtr_reads_upload", "mtr_wdp_run", "mtr_wdp_upload", "mtr_wdp_launch", "mtr_wdp_download",
              "mtr_di_run", "mtr_get_stats", "mtr_alu_probe", "mtr_uf_run", "mtr_pipeline_open", "mtr_pipeline_load_fasta",
              "mtr_pipeline_run", "mtr_pipeline_get_stats"):
        getattr(lib, f).restype = C.c_int
    _lib = lib
    return lib


class MtrError(RuntimeError):
    pass


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_reads(reads: Sequence[np.ndarray], tails: Optional[Sequence[Sequence[int]]] = None):
    """2-bit packs reads (arrays of 0..3) into the layout mtr_reads_upload expects.

    Each read gets len + 2 bases: the two bases after the end are `tails[r]` (default 0, 0), the stale
    bases the reference would read from an earlier, longer read.  Reads start on 4-word
    (16-byte) boundaries so device code can use uint4 loads."""
    n = len(reads)
    word_off = np.zeros(n + 1, dtype=np.int64)
    lens = np.zeros(n, dtype=np.int32)
    for r, rd in enumerate(reads):
        lens[r] = len(rd)
        words = (len(rd) + 2 + 15) // 16
        word_off[r + 1] = word_off[r] + ((words + 3) // 4) * 4
    packed = np.zeros(int(word_off[n]), dtype=np.uint32)
    for r, rd in enumerate(reads):
        ext = np.zeros(((len(rd) + 2 + 15) // 16) * 16, dtype=np.uint32)
        ext[:len(rd)] = np.asarray(rd, dtype=np.uint32)
        if tails is not None:
            ext[len(rd)] = tails[r][0]
            ext[len(rd) + 1] = tails[r][1]
        w = (ext.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint64)
        packed[word_off[r]:word_off[r] + len(w)] = w.astype(np.uint32)
    return packed, word_off, lens


class Context:
    """One GPU context (mtr_ctx).  Raises MtrError when the library cannot get a CUDA device."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.mtr_cuda_init(device, C.byref(h))
        if rc != 0:
            raise MtrError("mtr_cuda_init(%d) failed (%d): %s" % (device, rc, self.lib.mtr_last_error(None).decode()))
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.mtr_cuda_shutdown(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise MtrError("%s failed (%d): %s" % (what, rc, self.lib.mtr_last_error(self.h).decode()))

    def upload_reads(self, packed: np.ndarray, word_off: np.ndarray, lens: np.ndarray):
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        word_off = np.ascontiguousarray(word_off, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        self._check(self.lib.mtr_reads_upload(self.h, _ptr(packed), _ptr(word_off), _ptr(lens), len(lens)), "mtr_reads_upload")
        self.n_reads = len(lens)
        self.lens = lens

    def wdp_run(self, jobs: np.ndarray, units: np.ndarray, aux_bytes: int = 0):
        """jobs: array of JOB_DTYPE; units: uint8 array.  Returns (results[n_jobs, 2], aux bytes or None)."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        units = np.ascontiguousarray(units, dtype=np.uint8)
        res = np.zeros((len(jobs), 2), dtype=RESULT_DTYPE)
        aux = np.zeros(max(aux_bytes, 1), dtype=np.uint8) if aux_bytes else None
        self._check(self.lib.mtr_wdp_run(self.h, _ptr(jobs), len(jobs), _ptr(units), units.size, _ptr(res),
                                         _ptr(aux), aux_bytes), "mtr_wdp_run")
        return res, aux

    def wdp_upload(self, jobs: np.ndarray, units: np.ndarray, aux_bytes: int = 0):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        units = np.ascontiguousarray(units, dtype=np.uint8)
        self._njobs = len(jobs)
        self._check(self.lib.mtr_wdp_upload(self.h, _ptr(jobs), len(jobs), _ptr(units), units.size, aux_bytes), "mtr_wdp_upload")

    def wdp_launch(self):
        self._check(self.lib.mtr_wdp_launch(self.h), "mtr_wdp_launch")

    def wdp_set_fused_traceback(self, on: bool):
        """True (default): the fill kernels run each task's traceback themselves; False: fill kernels, then one
        traceback kernel (so that the fill can be timed alone)."""
        self._check(self.lib.mtr_wdp_set_fused_traceback(self.h, 1 if on else 0), "mtr_wdp_set_fused_traceback")

    def wdp_download(self, aux_bytes: int = 0):
        res = np.zeros((self._njobs, 2), dtype=RESULT_DTYPE)
        aux = np.zeros(max(aux_bytes, 1), dtype=np.uint8) if aux_bytes else None
        self._check(self.lib.mtr_wdp_download(self.h, _ptr(res), _ptr(aux), aux_bytes), "mtr_wdp_download")
        return res, aux

    def di_run(self, manhattan: bool, stale: Optional[np.ndarray] = None, stale_off: Optional[np.ndarray] = None):
        """Directional index of every resident read.  Returns (pos_off, di, end, w)."""
        pos_off = np.zeros(self.n_reads + 1, dtype=np.int64)
        pos_off[1:] = np.cumsum(self.lens.astype(np.int64))
        total = int(pos_off[-1])
        di = np.zeros(max(total, 1), dtype=np.float64)
        end = np.zeros(max(total, 1), dtype=np.int32)
        w = np.zeros(max(total, 1), dtype=np.int32)
        if stale is not None:
            stale = np.ascontiguousarray(stale, dtype=np.uint16)
            stale_off = np.ascontiguousarray(stale_off, dtype=np.int64)
        self._check(self.lib.mtr_di_run(self.h, 1 if manhattan else 0, _ptr(stale), _ptr(stale_off), _ptr(pos_off),
                                        _ptr(di), _ptr(end), _ptr(w)), "mtr_di_run")
        return pos_off, di[:total], end[:total], w[:total]

    def stats(self) -> dict:
        s = Stats()
        self._check(self.lib.mtr_get_stats(self.h, C.byref(s)), "mtr_get_stats")
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def uf_run(self, tasks: np.ndarray):
        """tasks: array of UF_TASK_DTYPE.  Returns (results, units uint8, scores int32)."""
        tasks = np.ascontiguousarray(tasks, dtype=UF_TASK_DTYPE)
        cap = int(np.sum(2 * np.minimum(500, (tasks["qe"] - tasks["qs"]) // 5))) + 16
        res = np.zeros(len(tasks), dtype=UF_RESULT_DTYPE)
        units = np.zeros(cap, dtype=np.uint8)
        scores = np.zeros(cap, dtype=np.int32)
        used = C.c_int64()
        self._check(self.lib.mtr_uf_run(self.h, _ptr(tasks), len(tasks), _ptr(res), _ptr(units), _ptr(scores), cap,
                                        C.byref(used)), "mtr_uf_run")
        return res, units[:used.value], scores[:used.value]

    def engine_run(self, manhattan: bool = True, min_match_ratio: float = 0.6, stale: Optional[np.ndarray] = None,
                   stale_off: Optional[np.ndarray] = None):
        """handle_one_TR (handle_one_read.c:190-261) for every resident read on the device.  Returns (repeats as a
        REPEAT_DTYPE array ordered by (read, insertion order), units bytes, stats dict)."""
        reps, n, units, st = C.c_void_p(), C.c_int64(), C.c_void_p(), EngineStats()
        rc = self.lib.mtr_engine_run(self.h, 1 if manhattan else 0, min_match_ratio, _ptr(stale), _ptr(stale_off),
                                     C.byref(reps), C.byref(n), C.byref(units), C.byref(st))
        self._check(rc, "mtr_engine_run")
        k = int(n.value)
        if k == 0:
            arr, ub = np.zeros(0, REPEAT_DTYPE), np.zeros(0, np.uint8)
        else:
            arr = np.frombuffer((C.c_char * (k * REPEAT_DTYPE.itemsize)).from_address(reps.value), dtype=REPEAT_DTYPE).copy()
            total = int(arr["unit_off"][-1] + arr["rep_period"][-1])
            ub = np.frombuffer((C.c_char * total).from_address(units.value), dtype=np.uint8).copy()
        return arr, ub, {k2: getattr(st, k2) for k2, _ in EngineStats._fields_}

    def alu_probe(self, kind: int) -> float:
        """Giga lane-ops/s of the integer pipe (0: VIADDMNMX.RELU s32, 1: LOP3+IADD, 2: VIADDMNMX s16x2)."""
        g = C.c_double()
        self._check(self.lib.mtr_alu_probe(self.h, kind, C.byref(g)), "mtr_alu_probe")
        return g.value


class Pipeline:
    """Batch-level mirror of handle_one_file: FASTA text in host memory -> the text mTR prints.

    The reference reads its two options from globals (main.c:53-56); so does the library."""

    def __init__(self, device: int = 0, threads: int = 0, manhattan: bool = True, min_match_ratio: float = 0.6):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.mtr_pipeline_open(device, threads, C.byref(h))
        if rc != 0:
            raise MtrError("mtr_pipeline_open(%d) failed (%d): %s" % (device, rc, self.lib.mtr_last_error(None).decode()))
        self.h = h
        C.c_int.in_dll(self.lib, "Manhattan_Distance").value = 1 if manhattan else 0
        C.c_float.in_dll(self.lib, "min_match_ratio").value = min_match_ratio

    def close(self):
        if getattr(self, "h", None):
            self.lib.mtr_pipeline_close(self.h)
            self.h = None

    def load_fasta(self, text: bytes, first: int = 0, count: int = -1) -> int:
        """Loads the reads [first, first+count) of the FASTA text (default: all) and makes them resident."""
        n = self.lib.mtr_pipeline_load_fasta_shard(self.h, text, len(text), first, count)
        if n < 0:
            raise MtrError("mtr_pipeline_load_fasta failed (%d)" % n)
        return n

    def run(self, print_alignment: bool = False) -> bytes:
        out = C.c_char_p()
        n = C.c_int64()
        rc = self.lib.mtr_pipeline_run(self.h, 1 if print_alignment else 0, C.byref(out), C.byref(n))
        if rc != 0:
            raise MtrError("mtr_pipeline_run failed (%d)" % rc)
        return C.string_at(out, n.value)

    def stats(self) -> dict:
        s = PipelineStats()
        self.lib.mtr_pipeline_get_stats(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in PipelineStats._fields_}


def run_file(path: str, print_alignment: bool = False, manhattan: bool = True, min_match_ratio: float = 0.6):
    """handle_one_file (mTR.h:126) on a FASTA file, exactly as main.c calls it; what the library prints to stdout is
    captured through a temporary file.  Returns (number of reads, stdout bytes, counters of the call).  The process-wide
    runtime behind the entry point is configured by the MTR_* environment at its first use (MTR_DEVICE, MTR_GPUS,
    MTR_GROUPS_PER_GPU, MTR_GROUP_READS ...)."""
    import sys
    import tempfile
    lib = load_library()
    C.c_int.in_dll(lib, "Manhattan_Distance").value = 1 if manhattan else 0
    C.c_float.in_dll(lib, "min_match_ratio").value = min_match_ratio
    sys.stdout.flush()
    saved = os.dup(1)
    with tempfile.TemporaryFile() as tf:
        os.dup2(tf.fileno(), 1)
        try:
            n = lib.handle_one_file(path.encode(), 1 if print_alignment else 0)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tf.seek(0)
        out = tf.read()
    s = PipelineStats()
    lib.mtr_file_stats(C.byref(s))
    return n, out, {k: getattr(s, k) for k, _ in PipelineStats._fields_}
