#!/usr/bin/env python
"""bench.py — est-fact hot path on synthetic C3 (200 kbp genomic region x ESTs of 300-800 nt), one rank per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R]          our arm (CUDA through the C ABI)
  python bench.py --impl reference ...                                     the reference est-fact on the host cores

Two legs per run, both on synthetic C3 data, ESTs/s as the metric:
  value  one step = one pass of the DEVICE hot path over a batch of R ESTs per GPU that is already resident in HBM:
         maximal-pairing discovery (SEED) for every EST plus the DP jobs its simulated exon structure implies
         (compute_alignment on the first and last exon, K_band_edit_distance per exon, compute_gap_alignment per
         intron, the genome LCS scan for a fifth of the ESTs); timed with CUDA events on the launching stream.
  e2e    one step = one run of the shipped est-fact PROGRAM (pintron_b200/bin/est-fact: C host + libpintron_cuda.so
         through the C ABI) on E ESTs per GPU, exactly as pintron.py calls it: genomic.txt / ests.txt in the working
         directory, process start, CUDA context, index build, every H2D / D2H copy, the host control flow and the six
         output files are all inside the timed region (wall clock of the process).  This is the number to hold
         against the reference arm (`--impl reference`: the unmodified est-fact, one process per host core).
ESTs shard across ranks with no data-path collective: weak scaling, genome index replicated.
"""
import argparse
import json
import math
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
WORKLOADS = {"C3": "C3: synthetic 200 kbp genomic region x ESTs of 300-800 nt",
             "C4": "C4: synthetic 2 Mbp multi-gene locus x ESTs (80 %) and mRNAs of 1-6 kbp (20 %), long introns, polyA tails"}
OPS_PER_CELL = {"ALIGN": 5, "KBAND": 5, "GAP": 9}      # useful int ops per DP cell, SURVEY.md §8(d)


def kband_k(n):
    rate = 0.04 if n <= 50 else (0.035 if n <= 100 else 0.03)
    return max(1, math.ceil(n * rate))


def build_jobs(synth, start, count):
    """Jobs + algorithmic cell counts (the REFERENCE's recurrences: n*m, (2k+1)*m, 3*n*m)."""
    from pintron_b200 import Batch, PC_OP
    g = synth.genome
    b = Batch()
    cells = {"ALIGN": 0, "KBAND": 0, "GAP": 0, "LCS": 0}
    seed_bytes = 0
    n_reads = 0
    for _, _, pieces, fwd in synth.reads(start, count):
        n_reads += 1
        b.add(PC_OP.SEED, fwd, p0=15, out_cap=96)
        seed_bytes += len(fwd)
        offs = np.cumsum([0] + [q1 - q0 for q0, q1 in pieces])
        ex = [(fwd[offs[i]:offs[i + 1]], g[q0:q1]) for i, (q0, q1) in enumerate(pieces)]
        ex = [(e, t) for e, t in ex if e and t]
        if not ex:
            continue
        for e, t in (ex[0], ex[-1]):
            b.add(PC_OP.ALIGN, e, t)
            if e != t:
                cells["ALIGN"] += len(e) * len(t)
        for e, t in ex:
            k = kband_k(len(t))
            b.add(PC_OP.KBAND, t, e, p0=k)
            n, m = max(len(e), len(t)), min(len(e), len(t))
            if e != t and n - m <= k:
                cells["KBAND"] += (2 * k + 1) * m if 2 * k + 1 < n else n * m
        for i in range(len(pieces) - 1):
            d1, a0 = pieces[i][1], pieces[i + 1][0]
            est = fwd[max(0, offs[i + 1] - 30):offs[i + 1] + 30]
            gen = g[max(pieces[i][0], d1 - 30):d1] + g[d1:d1 + 70] + g[a0 - 70:a0] + g[a0:min(pieces[i + 1][1], a0 + 30)]
            if est and gen:
                b.add(PC_OP.GAP, est, gen)
                cells["GAP"] += 3 * len(est) * len(gen)
        if n_reads % 5 == 0 and pieces[0][0] > 0:
            b.add(PC_OP.LCS, fwd[:40], b_in_genome=(0, pieces[0][0]))
            cells["LCS"] += pieces[0][0] * 40
    return b, cells, seed_bytes, n_reads


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


def run_reference(args, rank, world):
    """The UNMODIFIED reference est-fact (oracle/_ref/est-fact), one process per host core over EST shards."""
    if rank != 0:
        return
    from pintron_b200.synth import Synth
    exe = os.path.join(ROOT, "oracle", "_ref", "est-fact")
    cores = os.cpu_count() or 1
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/est-fact not built (make -C oracle ref)"}))
        return
    per_core = args.ref_reads_per_core
    synth = Synth(args.workload, reads=cores * per_core)
    gtxt = synth.genome_fasta()
    tmp = tempfile.mkdtemp(prefix="pintron_ref_")
    dirs = []
    for c in range(cores):
        d = os.path.join(tmp, f"shard{c}")
        os.makedirs(d)
        open(os.path.join(d, "genomic.txt"), "wb").write(gtxt)
        open(os.path.join(d, "ests.txt"), "wb").write(synth.ests_fasta(c * per_core, per_core))
        dirs.append(d)

    def one_step():
        t0 = time.perf_counter()
        procs = [subprocess.Popen([exe], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in dirs]
        rcs = [p.wait() for p in procs]
        assert all(r == 0 for r in rcs), rcs
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    shutil.rmtree(tmp, ignore_errors=True)
    n = cores * per_core
    sec = sum(times) / len(times)
    v = n / sec
    sample = f"{n} {args.workload} ESTs per step = {cores} shards x {per_core}, one est-fact process per core, each rebuilding its suffix tree"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "ESTs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload] + " (bounded sample)", "reads_per_step": n},
        "cpu_baseline": {"value": v, "unit": "ESTs/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "ESTs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline_sample(workload="C3", seconds_budget=20.0):
    """Reference est-fact on a bounded sample with every host core (kind=reference), else the oracle port."""
    from pintron_b200.synth import Synth
    exe = os.path.join(ROOT, "oracle", "_ref", "est-fact")
    cores = os.cpu_count() or 1
    if not os.path.exists(exe):
        return None
    per_core = 150 if workload == "C3" else 12
    synth = Synth(workload, reads=cores * per_core)
    tmp = tempfile.mkdtemp(prefix="pintron_cpu_")
    gtxt = synth.genome_fasta()
    dirs = []
    for c in range(cores):
        d = os.path.join(tmp, f"s{c}")
        os.makedirs(d)
        open(os.path.join(d, "genomic.txt"), "wb").write(gtxt)
        open(os.path.join(d, "ests.txt"), "wb").write(synth.ests_fasta(c * per_core, per_core))
        dirs.append(d)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([exe], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in dirs]
    rcs = [p.wait() for p in procs]
    sec = time.perf_counter() - t0
    shutil.rmtree(tmp, ignore_errors=True)
    if any(rcs):
        return None
    n = cores * per_core
    return {"value": n / sec, "unit": "ESTs/s", "cores": cores, "kind": "reference",
            "sample": f"{n} {workload} ESTs, {cores} est-fact processes (one per core, {per_core} ESTs each), wall {sec:.2f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=20000, help="ESTs per GPU per step, device leg")
    ap.add_argument("--e2e-reads", type=int, default=100000, help="ESTs per GPU per step, whole-program leg")
    ap.add_argument("--ref-reads-per-core", type=int, default=200)
    ap.add_argument("--workload", default="C3", choices=["C3", "C4"], help="synthetic shape (pintron_b200/synth.py); C3 is the default bench line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the whole-program leg (profiling runs: ncu would follow the child)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import pintron_b200
    from pintron_b200.binding import PC_RES_INTS
    from pintron_b200.synth import Synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the est-fact hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    synth = Synth(args.workload, reads=args.reads * world)
    cu = pintron_b200.Cuda(local)
    L = cu.L
    cu.genome_upload(synth.genome, 15, 0.2)
    int_peak = L.pc_measure_int_peak(cu.ctx)
    batch, cells, seed_bytes, n_reads = build_jobs(synth, rank * args.reads, args.reads)
    arena, jobs = batch.arrays()
    n = len(jobs)

    # pinned host buffers (e2e leg) and device-resident copies (value leg)
    h_arena = torch.from_numpy(arena).pin_memory()
    h_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).pin_memory()
    h_res = torch.zeros(n * PC_RES_INTS, dtype=torch.int32).pin_memory()
    h_var = torch.zeros(max(batch.var_bytes, 1), dtype=torch.uint8).pin_memory()
    d_arena, d_jobs = h_arena.cuda(), h_jobs.cuda()
    d_res = torch.zeros(n * PC_RES_INTS, dtype=torch.int32, device="cuda")
    d_var = torch.zeros(max(batch.var_bytes, 1) + 16, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    stream = torch.cuda.ExternalStream(L.pc_stream_cuda_stream(cu.st))
    jobs_host_ptr = jobs.ctypes.data

    def step_device():
        rc = L.pc_submit_device(cu.st, d_arena.data_ptr(), len(batch.arena), d_jobs.data_ptr(), jobs_host_ptr, n,
                                d_res.data_ptr(), d_var.data_ptr(), batch.var_bytes)
        assert rc == 0, L.pc_last_error()

    def step_host():
        rc = L.pc_submit(cu.st, h_arena.data_ptr(), len(batch.arena), h_jobs.data_ptr(), n, h_res.data_ptr(),
                         h_var.data_ptr(), batch.var_bytes)
        assert rc == 0, L.pc_last_error()

    def timed(step_fn, steps, with_timers=False):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps."""
        ms = []
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step_fn()
                e1.record(stream)
            assert L.pc_stream_sync(cu.st) == 0, L.pc_last_error()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        return ms

    timed(step_device, args.warmup)
    assert int((d_res.view(n, PC_RES_INTS)[:, 0] != 0).sum().item()) == 0, "a job failed on the device"
    timed(step_host, 1)

    # ---- whole-program leg: the shipped est-fact on E ESTs of this rank ------------------------------------------
    exe = os.path.join(ROOT, "pintron_b200", "bin", "est-fact")
    if not os.path.exists(exe):
        raise SystemExit("bench.py: pintron_b200/bin/est-fact is not built (python __graft_entry__.py)")
    work = tempfile.mkdtemp(prefix=f"pintron_e2e_r{rank}_")
    if not args.no_e2e:
        e2e_synth = Synth(args.workload, reads=args.e2e_reads * world)
        open(os.path.join(work, "genomic.txt"), "wb").write(e2e_synth.genome_fasta())
        open(os.path.join(work, "ests.txt"), "wb").write(e2e_synth.ests_fasta(rank * args.e2e_reads, args.e2e_reads))
    cores = os.cpu_count() or 1
    threads = max(1, (cores // world) * 3 // 4)
    e2e_info = {}

    def step_program():
        t0 = time.perf_counter()
        p = subprocess.run([exe, "--devices", str(local), "--threads", str(threads)], cwd=work, stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE)
        sec = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr.decode("latin1")[-2000:]
        for line in p.stderr.decode("latin1").splitlines():
            if "bytes host->device" in line:
                w = line.replace(",", "").split()
                e2e_info["h2d"], e2e_info["d2h"] = int(w[w.index("host->device:") + 1]), int(w[w.index("device->host:") + 1])
            if "scheduler:" in line and "workers" in line:
                w = line.replace(",", "").split()
                e2e_info["context_index_s"] = float(w[w.index("index") + 1])
                e2e_info["workers_s"] = float(w[w.index("workers") + 1])
            if "@Timer Total" in line:
                e2e_info["program_total_s"] = int(line.split()[-2]) / 1e6
            if "kernel launches:" in line:
                w = line.replace(",", "").split()
                e2e_info["launches"] = int(w[w.index("launches:") + 1])
                e2e_info["jobs"] = int(w[w.index("jobs:") + 1])
        return sec

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    L.pc_stream_enable_timers(cu.st, 1)
    L.pc_stream_reset_timers(cu.st)
    launches0 = cu.launch_count()
    ms_dev = timed(step_device, args.steps)
    launches = cu.launch_count() - launches0
    import ctypes as C
    op_names = ["ALIGN", "KBAND", "EDIT", "BORDERS", "GAP", "AFFIX", "SUFCUT", "PRECUT", "LCS", "SEED"]
    op_ms, op_launch = {}, {}
    for i, nm in enumerate(op_names):
        m, k = C.c_double(), C.c_uint64()
        L.pc_stream_op_time(cu.st, i, C.byref(m), C.byref(k))
        if k.value:
            op_ms[nm], op_launch[nm] = m.value / args.steps, k.value // args.steps
    L.pc_stream_enable_timers(cu.st, 0)
    barrier()
    ms_host = timed(step_host, args.steps)                  # the same device batch, submitted from pinned host buffers
    barrier()
    n_out = None
    if args.no_e2e:
        ms_e2e = [float("nan")]
    else:
        for _ in range(args.warmup):
            step_program()
        barrier()
        ms_e2e = [step_program() * 1e3 for _ in range(args.steps)]
        barrier()
        n_out = sum(1 for _ in open(os.path.join(work, "processed-ests.txt"), "rb")) // 2
    shutil.rmtree(work, ignore_errors=True)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t_dev = torch.tensor([sum(ms_dev) / len(ms_dev), sum(ms_e2e) / len(ms_e2e), sum(ms_host) / len(ms_host)], device="cuda",
                         dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_step, ms_step_e2e, ms_step_host = t_dev.tolist()
    total_reads = n_reads * world
    value = total_reads / (ms_step * 1e-3)
    e2e = args.e2e_reads * world / (ms_step_e2e * 1e-3)

    dp_cells = cells["ALIGN"] + cells["KBAND"] + cells["GAP"]
    dom = max(op_ms, key=op_ms.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        key = "k_gap_pairs<8> (compute_gap_alignment)"
        if dom == "GAP" and args.reads == 20000 and key in tr:
            traffic = tr[key]["dram_bytes_per_launch"]
    except Exception:
        pass
    if dom in OPS_PER_CELL:
        ach = cells[dom] * OPS_PER_CELL[dom] / (op_ms[dom] * 1e-3) / 1e12
        roof = {"bound": "int_alu", "kernel": "k_gap_pairs<8> (compute_gap_alignment)" if dom == "GAP" else f"k_warp_per_job<{dom}>", "achieved": ach, "peak": int_peak / 1e12,
                "unit": "Tlane-op/s", "frac": ach / (int_peak / 1e12) if int_peak else None, "traffic": traffic,
                "traffic_note": "DRAM bytes per launch from the committed ncu capture (profiles/r1_traffic.json); this kernel is INT-ALU bound, "
                                "the traffic is its direction-byte scratch",
                "peak_source": "pc_measure_int_peak (VIADDMNMX chains, measured live on this GPU)",
                "gcups": cells[dom] / (op_ms[dom] * 1e-3) / 1e9, "ops_per_cell": OPS_PER_CELL[dom]}
    else:
        alg_bytes = seed_bytes + 12 * 40 * n_reads if dom == "SEED" else cells["LCS"] / 40
        ach = alg_bytes / (op_ms[dom] * 1e-3) / 1e9
        pk = peaks.get("hbm_gbs", 6650.0)
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk,
                "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}

    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_baseline_sample(args.workload)
        line = {
            "metric": METRIC, "value": value, "unit": "ESTs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload] + ", device hot path "
                                   "(SEED + ALIGN/KBAND/GAP/LCS jobs from the simulated exon structure) for `value`; the whole est-fact program for `e2e`",
                       "reads_per_gpu_per_step": n_reads, "jobs_per_gpu_per_step": n, "l2": "flushed between steps (256 MB write)",
                       "sharding": f"ESTs dealt to {world} rank(s), genome index replicated, no collective"},
            "dp_gcups": dp_cells * world / (sum(op_ms.get(k, 0) for k in ("ALIGN", "KBAND", "GAP")) * 1e-3) / 1e9,
            "dp_cells_per_step": dp_cells * world,
            "kernel_ms_per_step": op_ms, "kernel_launches_per_step": op_launch,
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "ESTs/s", "ms_per_step": ms_step_e2e,
                    "h2d_bytes_per_step": e2e_info.get("h2d"), "d2h_bytes_per_step": e2e_info.get("d2h"),
                    "what": "one run of the shipped est-fact program per step (process start, CUDA context, index build, "
                            "host control flow, every H2D/D2H copy, six output files): wall clock of the process",
                    "ests_per_gpu_per_step": args.e2e_reads, "ests_aligned_rank0": n_out, "host_threads_per_gpu": threads,
                    "device_jobs_per_step": e2e_info.get("jobs"), "gpu_launches_per_step": e2e_info.get("launches"),
                    "last_step_breakdown_s": {"cuda_context_and_genome_index": e2e_info.get("context_index_s"),
                                              "all_ests_through_workers": e2e_info.get("workers_s"),
                                              "program_total": e2e_info.get("program_total_s")}},
            "host_buffers_device_path": {"value": total_reads / (ms_step_host * 1e-3), "unit": "ESTs/s", "ms_per_step": ms_step_host,
                                         "h2d_bytes_per_step": int(len(batch.arena) + jobs.nbytes),
                                         "d2h_bytes_per_step": int(n * PC_RES_INTS * 4 + batch.var_bytes),
                                         "what": "the `value` batch submitted from pinned host buffers through pc_submit"},
            "gpu_launches": int(launches) + args.steps * int(e2e_info.get("launches") or 0), "clocks": sampler.summary(),
            "int_alu_peak_tlaneops": int_peak / 1e12,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    cu.close()


if __name__ == "__main__":
    main()
