#!/usr/bin/env python
"""bench.py — PIntron est-fact on B200: ESTs/s on synthetic data of BASELINE.json's shapes, one rank per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C4|C5]     our arm (CUDA through the C ABI)
  python bench.py --impl reference ...                                          the reference est-fact on the host cores

Workload (default C3 = BASELINE.json configs[2]: 200 kbp genomic region x 100 000 ESTs of 300-800 nt per GPU, weak
scaling; ESTs shard across ranks with no data-path collective, genome index replicated: SURVEY.md §8(e)).

  value   device-resident leg.  One step = EVERY device batch the shipped est-fact really issues for R ESTs of this rank
          — the engine's merged lane batches exactly as it formed them in a real run (PC_CAPTURE, pintron_b200/replay.py),
          all inputs already in HBM — submitted in order through pc_submit_device + pc_stream_sync over two pc_streams taken
          in turn (the engine's two submission loops), timed with CUDA events, L2 flushed between steps.  Launch overheads of the real batching are
          inside; host control flow and copies are not.
  e2e     the shipped est-fact PROGRAM on E ESTs per GPU exactly as pintron.py calls it (genomic.txt / ests.txt in the
          working directory, no arguments beyond execution knobs): process start, FASTA parsing, engine session (genome
          upload + index build), all host control flow, every H2D / D2H copy through pinned lanes, six output files;
          wall clock of the process; every step starts in a directory that holds only the two inputs.  The GPU server est-factd is resident, as deployed (started before the warm-up);
          `e2e_cold` is one run with the engine inside the process (CUDA context creation inside the timed region).
          This is the number to hold against `--impl reference`.
  kernels the same jobs as ONE merged batch with per-op CUDA-event timers: GCUPS and INT-ALU fractions per kernel
          (`roofline` = the dominant one) at full occupancy.
  parity  our five output files on the first P ESTs == the reference's (the cpu_baseline run's outputs, md5).
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "est_fact_ESTs_per_sec"
WORKLOADS = {"C3": ("C3: synthetic 200 kbp genomic region x 100000 ESTs of 300-800 nt per GPU", 100000, 20000),
             "C4": ("C4: synthetic 2 Mbp multi-gene locus x 30000 reads per GPU (80 % ESTs, 20 % mRNAs of 1-6 kbp, long introns, polyA tails)", 30000, 10000),
             "C5": ("C5: synthetic 200-exon titin-like gene x 200 full-length / partial mRNAs of 10-100 kbp per GPU", 200, 100)}
# useful integer ops per cell of the REFERENCE's recurrences.  GAP cells are plane-cells (3 per DP
# position: L, G, R), 9 ops per position = 3 per plane-cell; LCS: compare, two N tests, run update
OPS_PER_CELL = {"ALIGN": 5, "KBAND": 5, "EDIT": 5, "BORDERS": 5, "GAP": 3, "AFFIX": 5, "SUFCUT": 5, "PRECUT": 5, "LCS": 4}
KERNEL_OF = {"GAP": "k_gap_pairs (compute_gap_alignment)", "BORDERS": "k_borders_packed (general_refine_borders)",
             "EDIT": "k_myers<EDIT> (edit_distance)", "KBAND": "k_myers<KBAND> (K_band_edit_distance)",
             "ALIGN": "k_warp_per_job<ALIGN> (compute_alignment)", "AFFIX": "k_warp_per_job<AFFIX> (find_longest_affix)",
             "LCS": "k_lcs (find_longest_common_factor_dp)", "SEED": "k_seed (build_vertex_set)"}
FILES = ["raw-multifasta-out.txt", "processed-ests.txt", "megs.txt", "processed-megs.txt", "meg-edges.txt"]
OP_NAMES = ["ALIGN", "KBAND", "EDIT", "BORDERS", "GAP", "AFFIX", "SUFCUT", "PRECUT", "LCS", "SEED"]


def config_of(workload):
    """The SAME object in both arms (the driver compares it)."""
    return {"workload": WORKLOADS[workload][0], "ests_per_gpu": WORKLOADS[workload][1],
            "sharding": "ESTs dealt to ranks, genome index replicated per GPU, no data-path collective",
            "l2": "flushed between timed steps (256 MB write)"}


class ClockSampler(threading.Thread):
    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu, [], False

    def run(self):
        """NVML in-process (cheap); nvidia-smi as the fallback.  Sample rows: [sm_mhz, sm_max_mhz, hw_slowdown, hw_thermal,
        sw_thermal, sw_power_cap] with 'Active' / 'Not Active' strings for the reasons."""
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            bits = [(N.nvmlClocksThrottleReasonHwSlowdown, 2), (N.nvmlClocksThrottleReasonHwThermalSlowdown, 3),
                    (N.nvmlClocksThrottleReasonSwThermalSlowdown, 4), (N.nvmlClocksThrottleReasonSwPowerCap, 5)]
            while not self.stop_flag:
                row = [str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx), "Not Active", "Not Active", "Not Active", "Not Active"]
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for b, i in bits:
                    if r & b:
                        row[i] = "Active"
                self.samples.append(row)
                time.sleep(0.1)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, "reasons": reasons}


def _log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def fresh_run_dir(d):
    """Each whole-program step starts as pintron.py's STEP 2 does: a directory that holds genomic.txt and ests.txt only
    (leftovers of the previous step would make the program pay for truncating hundreds of MB of page cache)."""
    for f in os.listdir(d):
        if f not in ("genomic.txt", "ests.txt"):
            os.remove(os.path.join(d, f))


def md5_files(d):
    return {f: hashlib.md5(open(os.path.join(d, f), "rb").read()).hexdigest() for f in FILES}


class RefShards:
    """The UNMODIFIED reference est-fact (oracle/_ref/est-fact), one process per host core over contiguous EST shards
    (each rebuilding its suffix tree: the reference's real cost)."""

    def __init__(self, workload, per_core, cores=None):
        from pintron_b200.synth import Synth, ests_fasta_parallel
        self.exe = os.path.join(ROOT, "oracle", "_ref", "est-fact")
        self.cores = cores or (os.cpu_count() or 1)
        self.per_core, self.n = per_core, per_core * self.cores
        self.tmp = tempfile.mkdtemp(prefix="pintron_ref_")
        gtxt = Synth(workload, reads=1).genome_fasta()
        ests = ests_fasta_parallel(workload, self.n, 0, self.n, procs=max(1, self.cores - 1))
        recs = ests.split(b"\n>")
        recs = [recs[0]] + [b">" + r for r in recs[1:]]
        assert len(recs) == self.n, (len(recs), self.n)
        self.all_ests = ests
        self.dirs = []
        for c in range(self.cores):
            d = os.path.join(self.tmp, f"shard{c:03d}")
            os.makedirs(d)
            open(os.path.join(d, "genomic.txt"), "wb").write(gtxt)
            open(os.path.join(d, "ests.txt"), "wb").write(b"\n".join(recs[c * per_core:(c + 1) * per_core]) + (b"\n" if not recs[(c + 1) * per_core - 1].endswith(b"\n") else b""))
            self.dirs.append(d)
        self.genome_fasta = gtxt

    def available(self):
        return os.path.exists(self.exe)

    def step(self, timeout=None):
        """One pass over all shards; seconds, or None when `timeout` expired first (the processes are killed)."""
        for d in self.dirs:
            fresh_run_dir(d)
        t0 = time.perf_counter()
        procs = [subprocess.Popen([self.exe], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in self.dirs]
        rcs = []
        for p in procs:
            try:
                rcs.append(p.wait(timeout=None if timeout is None else max(0.1, timeout - (time.perf_counter() - t0))))
            except subprocess.TimeoutExpired:
                for q in procs:
                    if q.poll() is None:
                        q.kill()
                for q in procs:
                    q.wait()
                return None
        assert all(r == 0 for r in rcs), rcs
        return time.perf_counter() - t0

    def cells(self):
        """DP cells the reference computes on these ESTs, per routine: one pass of oracle/_ref/est-fact-cells (the same
        unmodified sources with counting interposers, oracle/ref_cells.c) over the shards; None when it is not built."""
        exe = os.path.join(ROOT, "oracle", "_ref", "est-fact-cells")
        if not os.path.exists(exe):
            return None
        procs = [subprocess.Popen([exe], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in self.dirs]
        if any(p.wait() != 0 for p in procs):
            return None
        tot = {}
        for d in self.dirs:
            for k, v in json.load(open(os.path.join(d, "cells.json"))).items():
                t = tot.setdefault(k, {"cells": 0, "calls": 0})
                t["cells"] += v["cells"]; t["calls"] += v["calls"]
        return tot

    def md5s(self):
        """md5 of the shard outputs concatenated in shard order (== the single run: ESTs are independent)."""
        out = {}
        for f in FILES:
            h = hashlib.md5()
            for d in self.dirs:
                h.update(open(os.path.join(d, f), "rb").read())
            out[f] = h.hexdigest()
        return out

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = config_of(args.workload)
    sh = RefShards(args.workload, args.ref_reads_per_core)
    if not sh.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/est-fact not built (make -C oracle ref)"}))
        return
    for _ in range(args.warmup):
        sh.step()
    times = [sh.step() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    v = sh.n / sec
    sample = (f"{sh.n} of the {cfg['ests_per_gpu']} {args.workload} ESTs per step = {sh.cores} shards x {sh.per_core}, one unmodified est-fact process "
              f"per host core, each building its own suffix tree; whole-process wall clock")
    sh.close()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "ESTs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": "ESTs/s", "cores": sh.cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "ESTs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def host_threads(world_gpus_on_box):
    """est-fact worker threads per GPU: this GPU's share of the host cores (cores // GPUs of the BOX, whatever N is, so
    that the N = 1 run of a scaling series uses what one GPU gets at N = 8), the engine's submission loop and the writers float over the same cores)."""
    cores = os.cpu_count() or 1
    share = max(1, cores // max(1, world_gpus_on_box))
    # measured on the 16-core box (tools/e2e_probe.py): 16 workers beat 14 by ~2 % although the engine's submission loop and the
    # writers then share cores with them
    return share - 1 if 6 <= share < 8 else share       # 16 cores / 1 GPU -> 16; 32 cores / 8 GPUs -> 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=None, help="ESTs per GPU per step, device legs (default 20000 C3)")
    ap.add_argument("--e2e-reads", type=int, default=None, help="ESTs per GPU per step, whole-program leg (default = the workload's count)")
    ap.add_argument("--e2e-max-steps", type=int, default=3)
    ap.add_argument("--e2e-max-warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=None, help="est-fact worker threads per GPU (default: this GPU's share of the cores)")
    ap.add_argument("--ref-reads-per-core", type=int, default=None, help="reference arm / cpu baseline: ESTs per host core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the whole-program legs (profiling runs)")
    ap.add_argument("--kernels-only", action="store_true", help="profiling runs: only the merged one-batch leg (ncu captures its big launches)")
    ap.add_argument("--no-extra", action="store_true", help="skip the short C4 / C5 whole-program legs appended to the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    name, e2e_default, dev_default = WORKLOADS[args.workload]
    args.reads = args.reads or dev_default
    args.e2e_reads = args.e2e_reads or e2e_default
    if args.ref_reads_per_core is None:
        args.ref_reads_per_core = {"C3": 1000, "C4": 150, "C5": 2}[args.workload] if args.impl == "reference" else {"C3": 200, "C4": 100, "C5": 1}[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import pintron_b200
    from pintron_b200 import replay
    from pintron_b200.binding import PC_RES_INTS
    from pintron_b200.synth import Synth, ests_fasta_parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the est-fact hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    exe = os.path.join(ROOT, "pintron_b200", "bin", "est-fact")
    daemon = os.path.join(ROOT, "pintron_b200", "bin", "est-factd")
    if not (os.path.exists(exe) and os.path.exists(daemon)):
        raise SystemExit("bench.py: pintron_b200/bin/est-fact / est-factd are not built (python __graft_entry__.py)")
    cores = os.cpu_count() or 1
    gpus_on_box = torch.cuda.device_count()
    threads = args.threads or host_threads(gpus_on_box)
    gen_procs = max(1, cores // world - 1)

    # ---- the resident server (one per box, all GPUs), as deployed ------------------------------------------------
    tag = os.environ.get("MASTER_PORT", str(os.getpid())) if world > 1 else str(os.getpid())
    srv_dir = os.path.join(tempfile.gettempdir(), f"pintron_bench_{tag}")
    sock = os.path.join(srv_dir, "efd.sock")
    srv = None
    if local == 0:
        shutil.rmtree(srv_dir, ignore_errors=True)
        os.makedirs(srv_dir)
        t0 = time.perf_counter()
        srv = subprocess.Popen([daemon, "--socket", sock, "--foreground", "--idle-timeout", "900"], stdout=open(os.path.join(srv_dir, "efd.log"), "wb"), stderr=subprocess.STDOUT)
        while not os.path.exists(sock):
            if srv.poll() is not None:
                raise SystemExit("bench.py: est-factd exited: " + open(os.path.join(srv_dir, "efd.log")).read()[-1500:])
            time.sleep(0.02)
        _log(f"est-factd up in {time.perf_counter() - t0:.2f} s ({gpus_on_box} GPU(s))")
    barrier()
    env_srv = dict(os.environ, EST_FACTD_SOCKET=sock, EST_FACT_NO_SPAWN="1")

    def est_fact(cwd, form, extra=(), env=None, timeout=None):
        """One run of the shipped program; returns (seconds, info parsed from its log)."""
        cmd = [exe, "--threads", str(threads), "--devices", str(local), "--engine", form, *extra]
        fresh_run_dir(cwd)
        t0 = time.perf_counter()
        p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env_srv if form == "daemon" else (env or os.environ), timeout=timeout)
        sec = time.perf_counter() - t0
        err = p.stderr.decode("latin1")
        assert p.returncode == 0, err[-2000:]
        info = {}
        for line in err.splitlines():
            w = line.replace(",", "").replace(";", "").split()
            if "bytes host->device" in line:
                info["h2d"], info["d2h"] = int(w[w.index("host->device:") + 1]), int(w[w.index("device->host:") + 1])
            if "scheduler:" in line and "workers" in line:
                info["workers_s"] = float(w[w.index("workers") + 1])
                info["session_open_s"] = float(w[w.index("after") + 1])
            if "timeline" in line and "ests.txt" in w:
                info["first_window_s"] = float(w[w.index("ests.txt") + 1])
            if "@Timer Total" in line:
                info["program_total_s"] = int(line.split()[-2]) / 1e6
            if "kernel launches:" in line:
                info["launches"] = int(w[w.index("launches:") + 1])
                info["jobs"] = int(w[w.index("jobs:") + 1])
                info["lane_batches"] = int(w[w.index("batches:") + 1])
            if "merged device batches" in line:
                info["device_batches"] = int(w[w.index("merged") - 1])
            if "thread-seconds" in line:
                info["per_est_code_thread_s"] = float(w[w.index("code") + 1])
                info["wait_on_device_thread_s"] = float(w[w.index("device") + 1])
        return sec, info

    def write_inputs(d, workload, total, start, count):
        open(os.path.join(d, "genomic.txt"), "wb").write(Synth(workload, reads=1).genome_fasta())
        open(os.path.join(d, "ests.txt"), "wb").write(ests_fasta_parallel(workload, total, start, count, procs=gen_procs))

    # ---- device workload: the device batches of a real run over this rank's first R ESTs --------------------------
    synth = Synth(args.workload, reads=1)
    cap_dir = tempfile.mkdtemp(prefix=f"pintron_cap_r{rank}_")
    write_inputs(cap_dir, args.workload, args.e2e_reads * world, rank * args.e2e_reads, args.reads)
    _log("inputs written; capture run of est-fact (engine in-process, PC_CAPTURE)")
    cap_file = os.path.join(cap_dir, "jobs.capture")
    est_fact(cap_dir, "inproc", ("--no-aux-outputs",), env=dict(os.environ, PC_CAPTURE=cap_file))
    batches = replay.device_batches(cap_file)
    m_arena, m_jobs, m_var, n_records = replay.merge(cap_file)
    shutil.rmtree(cap_dir, ignore_errors=True)
    cells = replay.algorithmic_cells(m_arena, synth.genome, m_jobs)
    n_jobs = len(m_jobs)
    jobs_per_op = {nm: int((m_jobs["op"] == i).sum()) for i, nm in enumerate(OP_NAMES) if (m_jobs["op"] == i).any()}
    _log(f"{n_jobs} jobs in {len(batches)} device batches ({n_records} lane batches), arena {len(m_arena) >> 20} MB")

    for k_, v_ in os.environ.items():          # kernel timing experiments: PC_*_BENCH=x reaches this process's library only, not the capture run
        if k_.startswith("PC_") and k_.endswith("_BENCH"):
            os.environ[k_[:-6]] = v_
    cu = pintron_b200.Cuda(local)
    L = cu.L
    cu.genome_upload(synth.genome, 15, 0.2)
    int_peak = L.pc_measure_int_peak(cu.ctx)
    stream = torch.cuda.ExternalStream(L.pc_stream_cuda_stream(cu.st))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    # every device batch resident in HBM: one arena / jobs / res / var tensor, a slice per batch
    tot_a = sum(len(b[0]) + 32 for b in batches)
    tot_j = sum(len(b[1]) for b in batches)
    tot_v = sum(b[2] + 32 for b in batches)
    d_arena = torch.zeros(tot_a + 64, dtype=torch.uint8, device="cuda")
    d_jobs = torch.zeros(max(tot_j, 1) * 44, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(max(tot_j, 1) * PC_RES_INTS, dtype=torch.int32, device="cuda")
    d_var = torch.zeros(tot_v + 64, dtype=torch.uint8, device="cuda")
    plan, oa, oj, ov = [], 0, 0, 0
    for arena, jobs, var_bytes, _ in batches:
        d_arena[oa:oa + len(arena)].copy_(torch.from_numpy(arena))
        d_jobs[oj * 44:(oj + len(jobs)) * 44].copy_(torch.from_numpy(jobs.view(np.uint8)))
        # small batches are ordered by the submitting thread from the host copy of the jobs, as the engine does (pc_submit_parts)
        plan.append((d_arena.data_ptr() + oa, len(arena), d_jobs.data_ptr() + oj * 44, len(jobs), d_res.data_ptr() + oj * PC_RES_INTS * 4,
                     d_var.data_ptr() + ov, var_bytes, jobs.ctypes.data if len(jobs) < (1 << 14) else None, jobs))
        oa += (len(arena) + 32 + 15) & ~15; oj += len(jobs); ov += (var_bytes + 32 + 15) & ~15
    torch.cuda.synchronize()

    # two pc_streams taken in turn, as the engine's two submission loops do (pc_engine.cu): batch k+1 is formed and enqueued
    # while batch k runs; a stream is synchronised before it takes its next batch and at the end of the step
    st2 = L.pc_stream_create(cu.ctx)
    assert st2, L.pc_last_error()
    sts = [cu.st, st2]

    def step_batches():
        busy = [False, False]
        for k, (a, ab, j, n, r, v, vb, hj, _keep) in enumerate(plan):
            q = k & 1
            if busy[q]:
                assert L.pc_stream_sync(sts[q]) == 0, L.pc_last_error()
            rc = L.pc_submit_device(sts[q], a, ab, j, hj, n, r, v, vb)
            assert rc == 0, L.pc_last_error()
            busy[q] = True
        for q in (0, 1):
            if busy[q]:
                assert L.pc_stream_sync(sts[q]) == 0, L.pc_last_error()

    # the same jobs as ONE batch (kernel-level leg)
    md_arena = torch.zeros(len(m_arena) + 16, dtype=torch.uint8, device="cuda")
    md_arena[:len(m_arena)].copy_(torch.from_numpy(m_arena))
    md_jobs = torch.from_numpy(m_jobs.view(np.uint8).copy()).cuda()
    md_res = torch.zeros(n_jobs * PC_RES_INTS, dtype=torch.int32, device="cuda")
    md_var = torch.zeros(max(m_var, 1) + 16, dtype=torch.uint8, device="cuda")

    def step_merged():
        rc = L.pc_submit_device(cu.st, md_arena.data_ptr(), len(m_arena), md_jobs.data_ptr(), None, n_jobs, md_res.data_ptr(), md_var.data_ptr(), m_var)
        assert rc == 0, L.pc_last_error()
        assert L.pc_stream_sync(cu.st) == 0, L.pc_last_error()

    def timed(step_fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream (pc_stream_sync waits for the side streams, so the
        closing event is after all of them), L2 flushed between steps."""
        ms = []
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step_fn()
                e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        return ms

    _log("device warm-up")
    if args.kernels_only:
        args.no_e2e = args.no_cpu_baseline = True
        plan[:] = plan[:1]
    timed(step_batches, args.warmup)
    st0 = d_res.view(-1, PC_RES_INTS)[:tot_j, 0]
    PC_E_OUTCAP = -2      # a SEED job whose triples did not fit: est-fact re-issues it with the reported capacity (both are in the stream)
    experiment = any(k_.endswith("_BENCH") for k_ in os.environ)
    assert experiment or int(((st0 != 0) & (st0 != PC_E_OUTCAP)).sum().item()) == 0, "a job failed on the device"
    timed(step_merged, 2)
    st1 = md_res.view(-1, PC_RES_INTS)[:, 0]
    assert experiment or int(((st1 != 0) & (st1 != PC_E_OUTCAP)).sum().item()) == 0, "a job failed on the device (merged batch)"

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = cu.launch_count()
    ms_dev = timed(step_batches, args.steps)
    launches_dev = cu.launch_count() - launches0
    barrier()
    _log("device leg timed; kernel-level leg")
    import ctypes as C
    L.pc_stream_enable_timers(cu.st, 1)
    L.pc_stream_reset_timers(cu.st)
    ksteps = min(args.steps, 5)
    ms_merged = timed(step_merged, ksteps)
    op_ms, op_launch = {}, {}
    for i, nm in enumerate(OP_NAMES):
        m, k = C.c_double(), C.c_uint64()
        L.pc_stream_op_time(cu.st, i, C.byref(m), C.byref(k))
        if k.value:
            op_ms[nm], op_launch[nm] = m.value / ksteps, k.value // ksteps
    L.pc_stream_enable_timers(cu.st, 0)
    barrier()

    # ---- whole-program legs ------------------------------------------------------------------------------------------
    work = tempfile.mkdtemp(prefix=f"pintron_e2e_r{rank}_")
    e2e_info, cold = {}, None
    n_out = None
    if args.no_e2e:
        ms_e2e = [float("nan")]
    else:
        write_inputs(work, args.workload, args.e2e_reads * world, rank * args.e2e_reads, args.e2e_reads)
        e2e_warmup, e2e_steps = min(args.warmup, args.e2e_max_warmup), min(args.steps, args.e2e_max_steps)
        _log("whole-program leg (client of the resident est-factd)")
        for _ in range(e2e_warmup):
            est_fact(work, "daemon")
        barrier()
        ms_e2e = []
        for _ in range(e2e_steps):
            sec, e2e_info = est_fact(work, "daemon")
            ms_e2e.append(sec * 1e3)
        barrier()
        n_out = sum(1 for _ in open(os.path.join(work, "processed-ests.txt"), "rb")) // 2
        if rank == 0:
            _log("whole-program, cold: engine inside the process")
            sec, ci = est_fact(work, "inproc")
            cold = {"value": args.e2e_reads / sec, "unit": "ESTs/s", "ms_per_step": sec * 1e3, "gpus": 1,
                    "what": "one run with --engine inproc: CUDA context creation, kernel loading and lane pinning inside the timed region",
                    "session_open_s": ci.get("session_open_s"), "workers_s": ci.get("workers_s")}
        barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t_dev = torch.tensor([sum(ms_dev) / len(ms_dev), sum(ms_e2e) / len(ms_e2e)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_step, ms_step_e2e = t_dev.tolist()
    value = args.reads * world / (ms_step * 1e-3)
    e2e = args.e2e_reads * world / (ms_step_e2e * 1e-3)

    # ---- roofline of the dominant kernel (kernel-level leg) ----------------------------------------------------------
    peaks, ncu = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_kernels.json"))).get(args.workload, {})
    except Exception:
        pass
    dom = max(op_ms, key=op_ms.get)

    def kernel_entry(k):
        e = {"ms": op_ms[k], "launches": op_launch.get(k)}
        if k in cells and k in OPS_PER_CELL:
            e["cells"] = cells[k]
            e["gcups"] = cells[k] / (op_ms[k] * 1e-3) / 1e9
            e["int_alu_frac"] = cells[k] * OPS_PER_CELL[k] / (op_ms[k] * 1e-3) / int_peak if int_peak else None
            e["ops_per_cell"] = OPS_PER_CELL[k]
        n = ncu.get(k)
        if n:
            e["ncu"] = n
        return e
    per_kernel = {k: kernel_entry(k) for k in op_ms}
    if dom in OPS_PER_CELL:
        ach = cells[dom] * OPS_PER_CELL[dom] / (op_ms[dom] * 1e-3) / 1e12
        n = ncu.get(dom, {})
        roof = {"bound": "int_alu", "kernel": KERNEL_OF.get(dom, dom), "achieved": ach, "peak": int_peak / 1e12, "unit": "Tlane-op/s",
                "frac": ach / (int_peak / 1e12) if int_peak else None,
                "traffic": n.get("dram_bytes_per_launch") if n.get("reads") == args.reads else None,
                "ncu_alu_pipe_pct": n.get("alu_pipe_pct"), "ncu_issue_active_pct": n.get("issue_active_pct"),
                "ncu_source": "profiles/r2_ncu_kernels.json (ncu --set full of this command; null when captured at another batch size)",
                "peak_source": "pc_measure_int_peak: VIADDMNMX chains measured live on this GPU = 148 SMs x 64 lanes/clk x SM clock "
                               "(MEASURED_PEAKS.json has no integer figure)",
                "gcups": cells[dom] / (op_ms[dom] * 1e-3) / 1e9, "ops_per_cell": OPS_PER_CELL[dom],
                "cells_note": "GAP cells are plane-cells: 3 per DP position, 9 reference ops per position" if dom == "GAP" else None}
    else:
        seed_sel = m_jobs["op"] == 9
        emitted = int(md_res.view(n_jobs, PC_RES_INTS)[torch.from_numpy(np.nonzero(seed_sel)[0]).cuda(), 1].clamp(min=0).sum().item())
        alg_bytes = int(m_jobs["a_len"][seed_sel].sum()) + 12 * emitted
        ach = alg_bytes / (op_ms[dom] * 1e-3) / 1e9
        pk = peaks.get("hbm_gbs", 6650.0)
        roof = {"bound": "hbm", "kernel": KERNEL_OF.get(dom, dom), "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "note": "genome and index are L2-resident at these sizes: the kernel is latency bound, the HBM fraction is reported as the contract asks"}

    if rank == 0:
        # ---- parity + CPU baseline: the reference on a bounded sample with every host core, its bytes against ours -------
        cpu = parity = cells_basis = None
        if not args.no_cpu_baseline:
            _log("reference on the host cores (cpu_baseline) and byte parity on the same ESTs")
            sh = RefShards(args.workload, args.ref_reads_per_core)
            if sh.available():
                sec = sh.step()
                if world == 1:
                    cpu = {"value": sh.n / sec, "unit": "ESTs/s", "cores": sh.cores, "kind": "reference",
                           "sample": f"the first {sh.n} {args.workload} ESTs, {sh.cores} unmodified est-fact processes (one per core, {sh.per_core} ESTs each, "
                                     f"each building its suffix tree), wall {sec:.2f} s"}
                pd = tempfile.mkdtemp(prefix="pintron_parity_")
                open(os.path.join(pd, "genomic.txt"), "wb").write(sh.genome_fasta)
                open(os.path.join(pd, "ests.txt"), "wb").write(sh.all_ests)
                est_fact(pd, "daemon")
                ours, ref = md5_files(pd), sh.md5s()
                bad = [f for f in FILES if ours[f] != ref[f]]
                parity = {"status": "identical" if not bad else "MISMATCH", "ests": sh.n, "files": FILES, "differing": bad,
                          "how": "md5 of our five output files (client of est-factd) == md5 of the reference's shard outputs concatenated"}
                ref_cells = sh.cells()
                if ref_cells:
                    # the same ESTs through our host: the jobs it issues, counted with the same formulas
                    cap = os.path.join(pd, "jobs.capture")
                    est_fact(pd, "inproc", ("--no-aux-outputs",), env=dict(os.environ, PC_CAPTURE=cap))
                    pa, pj, _, _ = replay.merge(cap)
                    oc = replay.algorithmic_cells(pa, synth.genome, pj)
                    oj = {nm: int((pj["op"] == i).sum()) for i, nm in enumerate(OP_NAMES)}
                    oc["AFFIX"] = sum(oc.get(k, 0) for k in ("AFFIX", "SUFCUT", "PRECUT"))
                    oj["AFFIX"] = sum(oj.get(k, 0) for k in ("AFFIX", "SUFCUT", "PRECUT"))
                    cells_basis = {"ests": sh.n, "per_routine": {}, "how": "reference: oracle/_ref/est-fact-cells = the unmodified sources with counting "
                                   "interposers on compute_alignment, K_band_edit_distance, edit_distance, compute_edit_distance, edit_distance_matrix, "
                                   "compute_gap_alignment (oracle/ref_cells.c); ours: the captured job stream of est-fact on the same ESTs; AFFIX = "
                                   "edit_distance_matrix callers (longest affix + prefix / suffix cuts); LCS is a static function in the reference "
                                   "and is not counted"}
                    for k, rv in ref_cells.items():
                        ratio = oc.get(k, 0) / rv["cells"] if rv["cells"] else None
                        cells_basis["per_routine"][k] = {"reference_cells": rv["cells"], "reference_calls": rv["calls"], "our_cells": oc.get(k, 0),
                                                          "our_jobs": oj.get(k, 0), "ours_over_reference": ratio}
                        for kk in ((k,) if k != "AFFIX" else ("AFFIX", "SUFCUT", "PRECUT")):
                            if ratio and kk in per_kernel and "gcups" in per_kernel[kk]:
                                per_kernel[kk]["gcups_reference_basis"] = per_kernel[kk]["gcups"] / ratio
                    if roof.get("gcups") and cells_basis["per_routine"].get(dom, {}).get("ours_over_reference"):
                        roof["gcups_reference_basis"] = roof["gcups"] / cells_basis["per_routine"][dom]["ours_over_reference"]
                shutil.rmtree(pd, ignore_errors=True)
            sh.close()
        extra = {}
        if not (args.no_e2e or args.no_extra) and args.workload == "C3" and world == 1:
            extra = extra_workloads(est_fact, write_inputs, exe)
        line = {
            "metric": METRIC, "value": value, "unit": "ESTs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": config_of(args.workload),
            "value_what": f"device-resident leg: the {len(batches)} device batches est-fact's engine really formed for {args.reads} ESTs per GPU "
                          f"({n_jobs} jobs, {n_records} lane batches merged by the engine), inputs in HBM, pc_submit_device + pc_stream_sync per batch over two pc_streams taken in turn (the engine's two submission loops)",
            "device_leg": {"ests_per_gpu_per_step": args.reads, "jobs_per_step": n_jobs, "device_batches_per_step": len(batches),
                           "lane_batches_per_step": n_records, "launches_per_step": launches_dev // args.steps, "jobs_per_op": jobs_per_op},
            "kernels": {"what": "the same jobs as ONE merged batch, per-op CUDA-event timers (single stream)", "ms_per_step": sum(ms_merged) / len(ms_merged),
                        "ests_per_sec": args.reads / (sum(ms_merged) / len(ms_merged) * 1e-3), "per_kernel": per_kernel,
                        "cells_basis": cells_basis,
                        "gcups_basis": "gcups: cells of the jobs OUR host issues, counted with the reference's formulas (SURVEY.md §8(d)); the host issues some "
                                       "DP calls speculatively (all four splice-shift variants), so EDIT counts more cells than the reference would compute; gcups_reference_basis: the same time, only the cells the "
                                       "reference itself computes on these ESTs (cells_basis, measured on the parity sample)"},
            "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "e2e": {"value": e2e, "unit": "ESTs/s", "ms_per_step": ms_step_e2e,
                    "h2d_bytes_per_step": e2e_info.get("h2d"), "d2h_bytes_per_step": e2e_info.get("d2h"),
                    "what": "one run of the shipped est-fact program per step as pintron.py calls it (process start, FASTA parsing, engine session with "
                            "genome upload + index build, host control flow, every H2D/D2H copy through pinned lanes, six output files), client of the "
                            "resident est-factd; wall clock of the process, max over ranks",
                    "ests_per_gpu_per_step": args.e2e_reads, "steps": min(args.steps, args.e2e_max_steps),
                    "warmup": min(args.warmup, args.e2e_max_warmup), "ests_aligned_rank0": n_out, "host_threads_per_gpu": threads,
                    "host_cores": cores, "gpus_on_box": gpus_on_box,
                    "device_jobs_per_step": e2e_info.get("jobs"), "gpu_launches_per_step": e2e_info.get("launches"),
                    "lane_batches_per_step": e2e_info.get("lane_batches"), "device_batches_per_step": e2e_info.get("device_batches"),
                    "last_step_breakdown_s": {k: e2e_info.get(k) for k in ("first_window_s", "session_open_s", "workers_s", "program_total_s",
                                                                           "per_est_code_thread_s", "wait_on_device_thread_s")}},
            "e2e_cold": cold,
            "gpu_launches": int(launches_dev) + (0 if args.no_e2e else min(args.steps, args.e2e_max_steps)) * int(e2e_info.get("launches") or 0),
            "clocks": sampler.summary(), "int_alu_peak_tlaneops": int_peak / 1e12,
        }
        line.update(extra)
        print(json.dumps(line))
    shutil.rmtree(work, ignore_errors=True)
    if os.environ.get("PC_PROFILE"):
        L.pc_debug_dump()
    barrier()
    if srv is not None:
        subprocess.run([daemon, "--socket", sock, "--stop"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        try:
            srv.wait(timeout=20)
        except subprocess.TimeoutExpired:
            srv.kill()
        shutil.rmtree(srv_dir, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()
    L.pc_stream_destroy(st2)
    cu.close()


def extra_workloads(est_fact, write_inputs, exe):
    """Short whole-program runs of the other BASELINE.json shapes appended to the default (C3) line, so that the driver's
    record carries them: C4 (configs[3]) and C5 (configs[4]) subsamples, with the reference on a bounded sample of the
    same reads beside them.  Every leg has a time limit: an extra leg never takes the main line down."""
    out = {}
    for wl, reads, ref_per_core, limit in (("C4", 10000, 8, 45.0), ("C5", 8, 1, 30.0)):
        ent = {"workload": WORKLOADS[wl][0], "subsample": f"{reads} reads per GPU (the full shape is the --workload {wl} run)"}
        try:
            d = tempfile.mkdtemp(prefix=f"pintron_{wl}_")
            write_inputs(d, wl, reads, 0, reads)
            try:
                sec, info = est_fact(d, "daemon", timeout=limit)           # one run, no warm-up: the server is warm from the main legs
                ent["e2e"] = {"value": reads / sec, "unit": "reads/s", "ms_per_step": sec * 1e3, "reads": reads, "workers_s": info.get("workers_s"),
                              "device_jobs": info.get("jobs"), "gpu_launches": info.get("launches"), "device_batches": info.get("device_batches")}
            except subprocess.TimeoutExpired:
                ent["e2e"] = {"value": None, "note": f"{reads} reads did not finish within {limit:.0f} s"}
            shutil.rmtree(d, ignore_errors=True)
            sh = RefShards(wl, ref_per_core)
            if sh.available():
                rs = sh.step(timeout=limit)
                ent["reference"] = ({"value": sh.n / rs, "unit": "reads/s", "cores": sh.cores, "sample": f"the first {sh.n} reads, one process per core, wall {rs:.1f} s"}
                                    if rs is not None else {"value": None, "cores": sh.cores, "note": f"the first {sh.n} reads (one per core x {ref_per_core}) did not finish within {limit:.0f} s"})
            sh.close()
        except Exception as e:      # noqa: BLE001
            ent["error"] = repr(e)[:300]
        out[wl.lower()] = ent
    return out


if __name__ == "__main__":
    main()
