#!/usr/bin/env python
"""bench.py — est-fact hot path on synthetic data of BASELINE.json's shapes (default C3: 200 kbp genomic region x ESTs of
300-800 nt, 100 000 per GPU; --workload C4: 2 Mbp multi-gene locus x ESTs and mRNAs), one rank per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--workload C3|C4]     our arm (CUDA through the C ABI)
  python bench.py --impl reference ...                                                   the reference est-fact on the host cores

Two legs per run, ESTs/s as the metric:
  value  one step = one pass of the DEVICE hot path over a batch that is already resident in HBM: EVERY device job the
         shipped est-fact program issues for R ESTs of this rank (maximal-pairing discovery, compute_alignment,
         K_band_edit_distance, edit_distance, refine_borders, compute_gap_alignment, find_longest_affix, the genome LCS
         scan ...), recorded from a real run (PC_CAPTURE, pintron_b200/replay.py) and merged into ONE batch; timed with
         CUDA events on the launching stream, L2 flushed between steps.
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
OPS_PER_CELL = {"ALIGN": 5, "KBAND": 5, "EDIT": 5, "BORDERS": 5, "GAP": 9, "AFFIX": 5, "SUFCUT": 5, "PRECUT": 5}      # useful int ops per DP cell, SURVEY.md §8(d)
KERNEL_OF = {"GAP": "k_gap_pairs<8> (compute_gap_alignment)", "BORDERS": "k_warp_per_job<BORDERS> (general_refine_borders)",
             "EDIT": "k_warp_per_job<EDIT> (edit_distance)", "KBAND": "k_warp_per_job<KBAND> (K_band_edit_distance)",
             "ALIGN": "k_warp_per_job<ALIGN> (compute_alignment)", "AFFIX": "k_warp_per_job<AFFIX> (find_longest_affix)",
             "LCS": "k_lcs (find_longest_common_factor_dp)", "SEED": "k_seed (build_vertex_set)"}


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
    per_core = 150 if workload == "C3" else 100
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


def _log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C3", "C4"],
                    help="synthetic shape (pintron_b200/synth.py).  C3 = BASELINE.json configs[2] (200 kbp x 100 000 ESTs per GPU), the default: "
                         "it finishes in two minutes.  C4 = configs[3] (2 Mbp multi-gene locus, ESTs + mRNAs): about one read per thousand is an "
                         "mRNA with a thousand candidate embeddings and end exons of several kbp (a minute each for the reference), so a run "
                         "is longer and noisier")
    ap.add_argument("--reads", type=int, default=None, help="ESTs per GPU per step, device leg (default 20000 for C3, 10000 for C4)")
    ap.add_argument("--e2e-reads", type=int, default=None, help="ESTs per GPU per step, whole-program leg (default 100000 C3, 30000 C4)")
    ap.add_argument("--e2e-max-steps", type=int, default=2)
    ap.add_argument("--e2e-max-warmup", type=int, default=1)
    ap.add_argument("--ref-reads-per-core", type=int, default=None, help="reference arm: ESTs per host core per step (default 200 C3, 150 C4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the whole-program leg (profiling runs: ncu would follow the child)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    c4 = args.workload == "C4"
    args.reads = args.reads or (10000 if c4 else 20000)
    args.e2e_reads = args.e2e_reads or (30000 if c4 else 100000)
    args.ref_reads_per_core = args.ref_reads_per_core or (150 if c4 else 200)

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

    exe = os.path.join(ROOT, "pintron_b200", "bin", "est-fact")
    if not os.path.exists(exe):
        raise SystemExit("bench.py: pintron_b200/bin/est-fact is not built (python __graft_entry__.py)")
    cores = os.cpu_count() or 1
    per_gpu = max(1, cores // world)
    threads = per_gpu if per_gpu <= 4 else per_gpu * 3 // 4      # est-fact worker threads per GPU (few cores per GPU: use them all)

    # ---- the device workload: the job stream of a real est-fact run over this rank's R ESTs, merged into one batch ----
    from pintron_b200 import replay
    from pintron_b200.synth import ests_fasta_parallel
    synth = Synth(args.workload, reads=args.reads * world)
    cap_dir = tempfile.mkdtemp(prefix=f"pintron_cap_r{rank}_")
    open(os.path.join(cap_dir, "genomic.txt"), "wb").write(synth.genome_fasta())
    open(os.path.join(cap_dir, "ests.txt"), "wb").write(ests_fasta_parallel(args.workload, args.reads * world, rank * args.reads, args.reads,
                                                                             procs=max(1, cores // world - 1)))
    _log("inputs written; capture run of est-fact")
    cap = replay.capture(exe, cap_dir, threads=threads, device=local)
    _log("capture done; merging")
    for l in getattr(replay.capture, "last_log", []):
        _log("  est-fact: " + l.strip()[:400])
    arena, jobs, var_bytes, n_batches = replay.merge(cap)
    shutil.rmtree(cap_dir, ignore_errors=True)
    # the genome bytes the device holds are est-fact's: N tails stripped (io-multifasta.c:830); for synthetic ACGT genomes = as is
    cells = replay.algorithmic_cells(arena, synth.genome, jobs)
    n = len(jobs)
    n_reads = args.reads
    seed_sel = jobs["op"] == 9
    seed_bytes = int(jobs["a_len"][seed_sel].sum())
    jobs_per_op = {nm: int((jobs["op"] == i).sum()) for i, nm in enumerate(replay.OP_NAMES) if (jobs["op"] == i).any()}

    _log(f"merged {len(jobs)} jobs from {n_batches} batches, arena {len(arena) >> 20} MB")
    cu = pintron_b200.Cuda(local)
    L = cu.L
    cu.genome_upload(synth.genome, 15, 0.2)
    int_peak = L.pc_measure_int_peak(cu.ctx)

    # pinned host buffers (e2e leg) and device-resident copies (value leg)
    h_arena = torch.from_numpy(arena).pin_memory()
    h_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).pin_memory()
    h_res = torch.zeros(n * PC_RES_INTS, dtype=torch.int32).pin_memory()
    h_var = torch.zeros(max(var_bytes, 1), dtype=torch.uint8).pin_memory()
    d_arena = torch.zeros(len(arena) + 16, dtype=torch.uint8, device="cuda")      # pc_submit_device: readable 16 bytes past the arena
    d_arena[:len(arena)].copy_(h_arena)
    d_jobs = h_jobs.cuda()
    d_res = torch.zeros(n * PC_RES_INTS, dtype=torch.int32, device="cuda")
    d_var = torch.zeros(max(var_bytes, 1) + 16, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    stream = torch.cuda.ExternalStream(L.pc_stream_cuda_stream(cu.st))
    jobs_host_ptr = jobs.ctypes.data

    def step_device():
        rc = L.pc_submit_device(cu.st, d_arena.data_ptr(), len(arena), d_jobs.data_ptr(), jobs_host_ptr, n,
                                d_res.data_ptr(), d_var.data_ptr(), var_bytes)
        assert rc == 0, L.pc_last_error()

    def step_host():
        rc = L.pc_submit(cu.st, h_arena.data_ptr(), len(arena), h_jobs.data_ptr(), n, h_res.data_ptr(),
                         h_var.data_ptr(), var_bytes)
        assert rc == 0, L.pc_last_error()

    def timed(step_fn, steps, with_timers=False):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps.  The closing event
        is recorded after pc_stream_sync, so re-runs of jobs whose scratch slot was too small are inside the interval."""
        ms = []
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step_fn()
                assert L.pc_stream_sync(cu.st) == 0, L.pc_last_error()
                e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        return ms

    _log("device warm-up")
    timed(step_device, args.warmup)
    _log("device warm-up done")
    st0 = d_res.view(n, PC_RES_INTS)[:, 0]
    PC_E_OUTCAP = -2      # a SEED job whose triples did not fit: est-fact re-issues it with the reported capacity (both are in the stream)
    assert int(((st0 != 0) & (st0 != PC_E_OUTCAP)).sum().item()) == 0, "a job failed on the device"
    timed(step_host, 1)

    # ---- whole-program leg: the shipped est-fact on E ESTs of this rank ------------------------------------------
    work = tempfile.mkdtemp(prefix=f"pintron_e2e_r{rank}_")
    if not args.no_e2e:
        open(os.path.join(work, "genomic.txt"), "wb").write(Synth(args.workload, reads=1).genome_fasta())
        open(os.path.join(work, "ests.txt"), "wb").write(ests_fasta_parallel(args.workload, args.e2e_reads * world, rank * args.e2e_reads,
                                                                             args.e2e_reads, procs=max(1, cores // world - 1)))
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
    _log("device leg timed")
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
        # one run takes seconds to tens of seconds: W and K are capped for this leg (stated in the JSON line)
        e2e_warmup, e2e_steps = min(args.warmup, args.e2e_max_warmup), min(args.steps, args.e2e_max_steps)
        _log("whole-program leg")
        for _ in range(e2e_warmup):
            step_program()
        _log("whole-program warm-up done")
        barrier()
        ms_e2e = [step_program() * 1e3 for _ in range(e2e_steps)]
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

    dp_ops = [k for k in OPS_PER_CELL if k in cells]
    dp_cells = sum(cells[k] for k in dp_ops)
    dom = max(op_ms, key=op_ms.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        ent = tr.get(args.workload, {}).get(KERNEL_OF.get(dom, dom))
        if ent and ent.get("reads") == args.reads:
            traffic = ent["dram_bytes_per_launch"]
    except Exception:
        pass
    if dom in OPS_PER_CELL:
        ach = cells[dom] * OPS_PER_CELL[dom] / (op_ms[dom] * 1e-3) / 1e12
        roof = {"bound": "int_alu", "kernel": KERNEL_OF.get(dom, dom), "achieved": ach, "peak": int_peak / 1e12,
                "unit": "Tlane-op/s", "frac": ach / (int_peak / 1e12) if int_peak else None, "traffic": traffic,
                "traffic_note": "DRAM bytes per launch from the committed ncu capture (profiles/r1_traffic.json), null when that capture "
                                "was taken on another batch size",
                "peak_source": "pc_measure_int_peak (VIADDMNMX chains, measured live on this GPU)",
                "gcups": cells[dom] / (op_ms[dom] * 1e-3) / 1e9, "ops_per_cell": OPS_PER_CELL[dom]}
    else:
        # SEED: EST bytes + 12 B per emitted pairing; LCS: one byte of genome prefix per job and scanned position
        if dom == "SEED":
            emitted = int(d_res.view(n, PC_RES_INTS)[torch.from_numpy(np.nonzero(seed_sel)[0]).cuda(), 1].clamp(min=0).sum().item())
            alg_bytes = seed_bytes + 12 * emitted
        else:
            alg_bytes = int(jobs["b_len"][jobs["op"] == 8].sum())
        ach = alg_bytes / (op_ms[dom] * 1e-3) / 1e9
        pk = peaks.get("hbm_gbs", 6650.0)
        roof = {"bound": "hbm", "kernel": KERNEL_OF.get(dom, dom), "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk,
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "note": "genome and index are L2-resident at these sizes: the kernel is latency / ALU bound, the HBM fraction is reported as the contract asks"}
    # every DP kernel against the INT-ALU peak (the dominant one is `roofline`)
    per_kernel = {k: {"ms": op_ms[k], "gcups": cells[k] / (op_ms[k] * 1e-3) / 1e9,
                      "int_alu_frac": cells[k] * OPS_PER_CELL[k] / (op_ms[k] * 1e-3) / int_peak if int_peak else None}
                  for k in dp_ops if k in op_ms and op_ms[k] > 0}

    if rank == 0:
        _log("cpu baseline sample")
        cpu = None if args.no_cpu_baseline else cpu_baseline_sample(args.workload)
        line = {
            "metric": METRIC, "value": value, "unit": "ESTs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload] + "; `value`: every device job the shipped est-fact issues for these ESTs "
                                   "(recorded from a real run, merged into one HBM-resident batch); `e2e`: the whole est-fact program",
                       "reads_per_gpu_per_step": n_reads, "jobs_per_gpu_per_step": n, "jobs_per_op": jobs_per_op,
                       "batches_merged": n_batches, "l2": "flushed between steps (256 MB write)",
                       "sharding": f"ESTs dealt to {world} rank(s), genome index replicated, no collective"},
            "dp_gcups": dp_cells * world / (sum(op_ms.get(k, 0) for k in dp_ops) * 1e-3) / 1e9,
            "dp_cells_per_step": dp_cells * world, "dp_kernels": per_kernel,
            "kernel_ms_per_step": op_ms, "kernel_launches_per_step": op_launch,
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "ESTs/s", "ms_per_step": ms_step_e2e,
                    "h2d_bytes_per_step": e2e_info.get("h2d"), "d2h_bytes_per_step": e2e_info.get("d2h"),
                    "what": "one run of the shipped est-fact program per step (process start, CUDA context, index build, "
                            "host control flow, every H2D/D2H copy, six output files): wall clock of the process",
                    "ests_per_gpu_per_step": args.e2e_reads, "steps": min(args.steps, args.e2e_max_steps),
                    "warmup": min(args.warmup, args.e2e_max_warmup), "ests_aligned_rank0": n_out, "host_threads_per_gpu": threads,
                    "device_jobs_per_step": e2e_info.get("jobs"), "gpu_launches_per_step": e2e_info.get("launches"),
                    "last_step_breakdown_s": {"cuda_context_and_genome_index": e2e_info.get("context_index_s"),
                                              "all_ests_through_workers": e2e_info.get("workers_s"),
                                              "program_total": e2e_info.get("program_total_s")}},
            "host_buffers_device_path": {"value": total_reads / (ms_step_host * 1e-3), "unit": "ESTs/s", "ms_per_step": ms_step_host,
                                         "h2d_bytes_per_step": int(len(arena) + jobs.nbytes),
                                         "d2h_bytes_per_step": int(n * PC_RES_INTS * 4 + var_bytes),
                                         "what": "the `value` batch submitted from pinned host buffers through pc_submit"},
            "gpu_launches": int(launches) + (0 if args.no_e2e else min(args.steps, args.e2e_max_steps)) * int(e2e_info.get("launches") or 0), "clocks": sampler.summary(),
            "int_alu_peak_tlaneops": int_peak / 1e12,
        }
        print(json.dumps(line))
    if os.environ.get("PC_PROFILE"):
        L.pc_debug_dump()
    if world > 1:
        dist.destroy_process_group()
    cu.close()


if __name__ == "__main__":
    main()
