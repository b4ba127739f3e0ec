"""ctypes mirror of include/pintron_cuda.h.  No CPU fallback: a missing library or device raises."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PC_RES_INTS = 8
PC_B_IN_GENOME = 1


class PC_OP:
    ALIGN, KBAND, EDIT, BORDERS, GAP, AFFIX, SUFCUT, PRECUT, LCS, SEED = range(10)


class LibraryMissing(RuntimeError):
    pass


class pc_job(C.Structure):
    _fields_ = [("op", C.c_uint32), ("flags", C.c_uint32), ("a_off", C.c_uint32), ("a_len", C.c_uint32),
                ("b_off", C.c_uint32), ("b_len", C.c_uint32), ("p0", C.c_int32), ("p1", C.c_int32), ("p2", C.c_int32),
                ("out_off", C.c_uint32), ("out_cap", C.c_uint32)]


JOB_DTYPE = np.dtype([("op", "<u4"), ("flags", "<u4"), ("a_off", "<u4"), ("a_len", "<u4"), ("b_off", "<u4"),
                      ("b_len", "<u4"), ("p0", "<i4"), ("p1", "<i4"), ("p2", "<i4"), ("out_off", "<u4"),
                      ("out_cap", "<u4")])
assert JOB_DTYPE.itemsize == C.sizeof(pc_job) == 44

EXPORTS = ["pc_last_error", "pc_device_count", "pc_ctx_create", "pc_ctx_destroy", "pc_genome_upload",
           "pc_stream_create", "pc_stream_destroy", "pc_host_alloc", "pc_host_free", "pc_submit", "pc_submit_device",
           "pc_stream_sync", "pc_compute_alignment_batch", "pc_kband_edit_distance_batch", "pc_edit_distance_batch",
           "pc_refine_borders_batch", "pc_gap_alignment_batch", "pc_longest_affix_batch", "pc_best_cut_batch",
           "pc_longest_common_factor_batch", "pc_build_vertex_set_batch", "pc_launch_count", "pc_stream_op_time",
           "pc_stream_reset_timers", "pc_stream_enable_timers", "pc_stream_cuda_stream", "pc_measure_int_peak"]


def library_path():
    return os.path.join(HERE, "libpintron_cuda.so")


def build_library():
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(HERE, "csrc")], check=True)


def load_library():
    path = library_path()
    if not os.path.exists(path):
        raise LibraryMissing(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    L.pc_last_error.restype = C.c_char_p
    L.pc_ctx_create.restype = C.c_void_p
    L.pc_ctx_create.argtypes = [C.c_int]
    L.pc_ctx_destroy.argtypes = [C.c_void_p]
    L.pc_genome_upload.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_double]
    L.pc_stream_create.restype = C.c_void_p
    L.pc_stream_create.argtypes = [C.c_void_p]
    L.pc_stream_destroy.argtypes = [C.c_void_p]
    L.pc_host_alloc.restype = C.c_void_p
    L.pc_host_alloc.argtypes = [C.c_size_t]
    L.pc_host_free.argtypes = [C.c_void_p]
    L.pc_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.pc_submit_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_size_t]
    L.pc_stream_sync.argtypes = [C.c_void_p]
    L.pc_launch_count.restype = C.c_uint64
    L.pc_stream_op_time.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.pc_stream_reset_timers.argtypes = [C.c_void_p]
    L.pc_stream_enable_timers.argtypes = [C.c_void_p, C.c_int]
    L.pc_stream_cuda_stream.restype = C.c_void_p
    L.pc_stream_cuda_stream.argtypes = [C.c_void_p]
    L.pc_measure_int_peak.restype = C.c_double
    L.pc_measure_int_peak.argtypes = [C.c_void_p]
    return L


class Batch:
    """Packs jobs (byte strings + op parameters) into the arena / job array layout of pc_submit."""

    def __init__(self):
        self.arena = bytearray()
        self.jobs = []
        self.var_bytes = 0

    def _put(self, s, pad=0):
        off = len(self.arena)
        self.arena += s
        self.arena += b"\0" * pad
        return off

    def add(self, op, a, b=b"", p0=0, p1=0, p2=0, b_in_genome=None, out_cap=0):
        a_off = self._put(a)
        if b_in_genome is not None:
            b_off, b_len, flags = b_in_genome[0], b_in_genome[1], PC_B_IN_GENOME
        else:
            b_off, b_len, flags = self._put(b, pad=1 if op == PC_OP.BORDERS else 0), len(b), 0
        out_off = 0
        if op in (PC_OP.ALIGN, PC_OP.GAP):
            out_cap = len(a) + b_len
            out_off, self.var_bytes = self.var_bytes, self.var_bytes + out_cap
        elif op == PC_OP.SEED:
            self.var_bytes = (self.var_bytes + 3) & ~3
            out_off, self.var_bytes = self.var_bytes, self.var_bytes + 12 * out_cap
        self.jobs.append((op, flags, a_off, len(a), b_off, b_len, p0, p1, p2, out_off, out_cap))
        return len(self.jobs) - 1

    def arrays(self):
        jobs = np.array(self.jobs, dtype=JOB_DTYPE) if self.jobs else np.zeros(0, dtype=JOB_DTYPE)
        arena = np.frombuffer(bytes(self.arena) or b"\0", dtype=np.uint8).copy()
        return arena, jobs


class Cuda:
    """One context (GPU) + one stream; mirrors the reference routine names for single calls and batches."""

    def __init__(self, device=0):
        self.L = load_library()
        self.ctx = self.L.pc_ctx_create(device)
        if not self.ctx:
            raise RuntimeError("pc_ctx_create failed: " + self.L.pc_last_error().decode())
        self.st = self.L.pc_stream_create(self.ctx)
        if not self.st:
            raise RuntimeError("pc_stream_create failed: " + self.L.pc_last_error().decode())

    def close(self):
        if self.st:
            self.L.pc_stream_destroy(self.st)
            self.st = None
        if self.ctx:
            self.L.pc_ctx_destroy(self.ctx)
            self.ctx = None

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.L.pc_last_error().decode()}")

    def genome_upload(self, genome, word_len=15, depth_rate=0.2):
        self._check(self.L.pc_genome_upload(self.ctx, genome, len(genome), word_len, depth_rate), "pc_genome_upload")

    def run(self, batch):
        """Submit a Batch through HOST buffers; returns (res[n,8] int32, var_out uint8)."""
        arena, jobs = batch.arrays()
        n = len(jobs)
        res = np.zeros((n, PC_RES_INTS), dtype=np.int32)
        var = np.zeros(max(batch.var_bytes, 1), dtype=np.uint8)
        self._check(self.L.pc_submit(self.st, arena.ctypes.data, len(batch.arena), jobs.ctypes.data, n, res.ctypes.data,
                                     var.ctypes.data, batch.var_bytes), "pc_submit")
        self._check(self.L.pc_stream_sync(self.st), "pc_stream_sync")
        return res, var

    def run_arrays(self, arena, jobs, var_bytes):
        """Same as run() for an arena / job array pair that was edited by hand (tests of the argument checks)."""
        n = len(jobs)
        res = np.zeros((n, PC_RES_INTS), dtype=np.int32)
        var = np.zeros(max(var_bytes, 1), dtype=np.uint8)
        self._check(self.L.pc_submit(self.st, arena.ctypes.data, len(arena), jobs.ctypes.data, n, res.ctypes.data,
                                     var.ctypes.data, var_bytes), "pc_submit")
        self._check(self.L.pc_stream_sync(self.st), "pc_stream_sync")
        return res, var

    def launch_count(self):
        return int(self.L.pc_launch_count())

    # ---- single-call conveniences named after the reference routines -----------------------------------
    def compute_alignment(self, est, gen):
        b = Batch(); b.add(PC_OP.ALIGN, est, gen)
        res, var = self.run(b)
        assert res[0, 0] == 0, res[0]
        return int(res[0, 1]), var[:res[0, 2]].tobytes()

    def K_band_edit_distance(self, s1, s2, k):
        b = Batch(); b.add(PC_OP.KBAND, s1, s2, p0=k)
        res, _ = self.run(b)
        assert res[0, 0] == 0, res[0]
        return bool(res[0, 1]), int(res[0, 2])

    def edit_distance(self, s1, s2):
        b = Batch(); b.add(PC_OP.EDIT, s1, s2)
        res, _ = self.run(b)
        assert res[0, 0] == 0, res[0]
        return int(res[0, 1])

    def general_refine_borders(self, p, t, max_errs, min_cut=0, max_cut=None):
        b = Batch(); b.add(PC_OP.BORDERS, p, t, p0=max_errs, p1=min_cut, p2=len(p) if max_cut is None else max_cut)
        res, _ = self.run(b)
        assert res[0, 0] == 0, res[0]
        return bool(res[0, 1]), [int(x) for x in res[0, 2:6]]

    def compute_gap_alignment(self, est, gen):
        b = Batch(); b.add(PC_OP.GAP, est, gen)
        res, var = self.run(b)
        assert res[0, 0] == 0, res[0]
        return var[:res[0, 1]].tobytes(), [int(x) for x in res[0, 2:7]]

    def find_longest_affix(self, est, gen):
        b = Batch(); b.add(PC_OP.AFFIX, est, gen)
        res, _ = self.run(b)
        assert res[0, 0] == 0, res[0]
        return (True, int(res[0, 2]), int(res[0, 3])) if res[0, 1] else (False, 0, 0)

    def compute_best_suffix_cut(self, s1, s2):
        b = Batch(); b.add(PC_OP.SUFCUT, s1, s2)
        res, _ = self.run(b)
        assert res[0, 0] == 0, res[0]
        return tuple(int(x) for x in res[0, 1:4])

    def compute_best_prefix_cut(self, s1, s2):
        b = Batch(); b.add(PC_OP.PRECUT, s1, s2)
        res, _ = self.run(b)
        assert res[0, 0] == 0, res[0]
        return tuple(int(x) for x in res[0, 1:4])

    def find_longest_common_factor_dp(self, s1, s2, s1_in_genome=None):
        b = Batch(); b.add(PC_OP.LCS, s2, s1, b_in_genome=s1_in_genome)
        res, _ = self.run(b)
        assert res[0, 0] == 0, res[0]
        return int(res[0, 1]), int(res[0, 2]), int(res[0, 3])

    def build_vertex_set(self, est, mfl=15, cap=1 << 14):
        b = Batch(); b.add(PC_OP.SEED, est, p0=mfl, out_cap=cap)
        res, var = self.run(b)
        assert res[0, 0] == 0, res[0]
        tri = var[:12 * res[0, 1]].view(np.int32).reshape(-1, 3)
        return [tuple(int(x) for x in r) for r in tri]
