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
ENGINE_EXPORTS = ["pc_engine_create", "pc_engine_destroy", "pc_engine_gpu_count", "pc_engine_backend", "pc_engine_open",
                  "pc_engine_resize_lane", "pc_engine_close", "pc_engine_enable_timers", "pc_engine_segment_count",
                  "pc_engine_segment_fd", "pc_engine_segment_base", "pc_submit_parts"]
PCE_MAX_LANES, PCE_MAX_SESSION_LANES = 512, 256
PCE_FREE, PCE_IDLE, PCE_POSTED, PCE_RUNNING, PCE_DONE = range(5)


class pc_part(C.Structure):
    _fields_ = [("arena", C.c_void_p), ("arena_bytes", C.c_size_t), ("jobs", C.c_void_p), ("njobs", C.c_int),
                ("res", C.c_void_p), ("var_out", C.c_void_p), ("var_out_bytes", C.c_size_t)]


class pc_session_req(C.Structure):
    _fields_ = [("gpu", C.c_int), ("genome", C.c_char_p), ("genome_len", C.c_size_t), ("word_len", C.c_int),
                ("depth_rate", C.c_double), ("nlanes", C.c_int), ("arena_cap", C.c_uint64), ("var_cap", C.c_uint64),
                ("jobs_cap", C.c_uint32)]


class pc_session_info(C.Structure):
    _fields_ = [("session", C.c_uint32), ("gpu", C.c_int), ("nlanes", C.c_int), ("lane", C.c_uint32 * PCE_MAX_SESSION_LANES)]


class pc_session_stats(C.Structure):
    _fields_ = [("batches", C.c_uint64), ("lanes_merged", C.c_uint64), ("jobs", C.c_uint64), ("launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("retries", C.c_uint64), ("busy_s", C.c_double),
                ("op_ms", C.c_double * 10)]


class pce_lane(C.Structure):          # include/pintron_engine.h, 128 bytes
    _fields_ = [("state", C.c_uint32), ("rc", C.c_int32), ("session", C.c_uint32), ("njobs", C.c_uint32),
                ("arena_len", C.c_uint64), ("var_len", C.c_uint64), ("seg", C.c_uint32), ("jobs_cap", C.c_uint32),
                ("arena_off", C.c_uint64), ("arena_cap", C.c_uint64), ("jobs_off", C.c_uint64), ("res_off", C.c_uint64),
                ("var_off", C.c_uint64), ("var_cap", C.c_uint64), ("batches", C.c_uint64), ("jobs_total", C.c_uint64),
                ("pad", C.c_uint8 * 24)]


assert C.sizeof(pce_lane) == 128
PCE_LANES_OFFSET = 128                # pce_hdr: magic, version, doorbell, sleepers, pad to 128, then the lanes


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
    L.pc_submit_parts.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(pc_part), C.c_int]
    L.pc_engine_create.restype = C.c_void_p
    L.pc_engine_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_size_t]
    L.pc_engine_destroy.argtypes = [C.c_void_p]
    L.pc_engine_backend.restype = C.c_char_p
    L.pc_engine_open.argtypes = [C.c_void_p, C.POINTER(pc_session_req), C.POINTER(pc_session_info)]
    L.pc_engine_resize_lane.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32]
    L.pc_engine_close.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(pc_session_stats)]
    L.pc_engine_enable_timers.argtypes = [C.c_void_p, C.c_int]
    L.pc_engine_segment_count.argtypes = [C.c_void_p, C.c_int]
    L.pc_engine_segment_fd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    L.pc_engine_segment_base.restype = C.c_void_p
    L.pc_engine_segment_base.argtypes = [C.c_void_p, C.c_int, C.c_int]
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

    def run_parts(self, batches):
        """Several Batches as ONE device batch (pc_submit_parts: what the engine does with the lanes it merges)."""
        keep, parts = [], (pc_part * len(batches))()
        for k, b in enumerate(batches):
            arena, jobs = b.arrays()
            res = np.zeros((len(jobs), PC_RES_INTS), dtype=np.int32)
            var = np.zeros(max(b.var_bytes, 1), dtype=np.uint8)
            keep.append((arena, jobs, res, var))
            parts[k] = pc_part(arena.ctypes.data, len(b.arena), jobs.ctypes.data, len(jobs), res.ctypes.data, var.ctypes.data, b.var_bytes)
        self._check(self.L.pc_submit_parts(self.st, None, parts, len(batches)), "pc_submit_parts")
        self._check(self.L.pc_stream_sync(self.st), "pc_stream_sync")
        return [(k[2], k[3]) for k in keep]

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


class Engine:
    """An in-process batch engine (include/pintron_engine.h) driven from Python: tests and bench.py post lanes the way
    the est-fact host does (pce_post / pce_wait are re-stated with plain stores and polling: no futex from Python)."""

    def __init__(self, devices=(0,), segment_bytes=0):
        self.L = load_library()
        arr = (C.c_int * len(devices))(*devices)
        self.e = self.L.pc_engine_create(arr, len(devices), segment_bytes)
        if not self.e:
            raise RuntimeError("pc_engine_create failed: " + self.L.pc_last_error().decode())

    def open(self, genome, nlanes, arena_cap, jobs_cap, var_cap, gpu=0, word_len=15, depth_rate=0.2):
        req = pc_session_req(gpu, genome, len(genome), word_len, depth_rate, nlanes, arena_cap, var_cap, jobs_cap)
        info = pc_session_info()
        if self.L.pc_engine_open(self.e, C.byref(req), C.byref(info)):
            raise RuntimeError("pc_engine_open failed: " + self.L.pc_last_error().decode())
        return Session(self, info)

    def close(self):
        if self.e:
            self.L.pc_engine_destroy(self.e)
            self.e = None


class Session:
    def __init__(self, eng, info):
        self.eng, self.id, self.gpu = eng, info.session, info.gpu
        self.lanes = [int(info.lane[k]) for k in range(info.nlanes)]
        self.hdr = eng.L.pc_engine_segment_base(eng.e, self.gpu, 0)

    def lane(self, k):
        return pce_lane.from_address(self.hdr + PCE_LANES_OFFSET + 128 * self.lanes[k])

    def _view(self, seg, off, nbytes, dtype=np.uint8):
        base = self.eng.L.pc_engine_segment_base(self.eng.e, self.gpu, seg)
        return np.ctypeslib.as_array(C.cast(base + off, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype)

    def post(self, k, batch):
        """Copy a Batch into lane k and post it (the doorbell is bumped; the engine polls with a 20 ms timeout at worst)."""
        l = self.lane(k)
        arena, jobs = batch.arrays()
        assert len(batch.arena) <= l.arena_cap and len(jobs) <= l.jobs_cap and batch.var_bytes <= l.var_cap, "lane too small"
        self._view(l.seg, l.arena_off, max(len(batch.arena), 1))[:len(batch.arena)] = arena[:len(batch.arena)]
        self._view(l.seg, l.jobs_off, jobs.nbytes or 44)[:jobs.nbytes] = jobs.view(np.uint8)
        l.njobs, l.arena_len, l.var_len = len(jobs), len(batch.arena), batch.var_bytes
        l.state = PCE_POSTED
        bell = C.c_uint32.from_address(self.hdr + 8)
        bell.value = bell.value + 1

    def wait(self, k, njobs, var_bytes, timeout=120.0):
        import time
        l = self.lane(k)
        t0 = time.time()
        while l.state != PCE_DONE:
            if time.time() - t0 > timeout:
                raise TimeoutError("engine did not finish the lane")
            time.sleep(0.0005)
        if l.rc:
            raise RuntimeError(f"engine batch failed ({l.rc})")
        res = self._view(l.seg, l.res_off, njobs * PC_RES_INTS * 4, np.int32).reshape(njobs, PC_RES_INTS).copy()
        var = self._view(l.seg, l.var_off, max(var_bytes, 1)).copy()
        return res, var

    def close(self):
        st = pc_session_stats()
        self.eng.L.pc_engine_close(self.eng.e, self.id, C.byref(st))
        return st
