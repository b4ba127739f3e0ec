"""CPU, world_size 2 over gloo: the multi-GPU contract of est-fact — ranks take contiguous shards of
ests.txt, run independently with the genome replicated, and the host-side gather of their outputs in rank order is the
single-run output, byte for byte.  Ranks run the host program against the CPU oracle backend (tests/cpu_backend/); on
a GPU box bench.py does the same with the CUDA library and --devices LOCAL_RANK."""
import hashlib
import os
import shutil
import tempfile

import torch.distributed as dist
import torch.multiprocessing as mp

import estfact_util as U

CASE = "test-mattia3"


def _rank_main(rank, world, port, cpu_bin, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tmp = tempfile.mkdtemp(prefix=f"shard{rank}_")
    try:
        exp = U.unpack(CASE, tmp)
        recs = open(os.path.join(tmp, "ests.txt"), "rb").read().split(b">")[1:]
        lo, hi = rank * len(recs) // world, (rank + 1) * len(recs) // world
        open(os.path.join(tmp, "ests.txt"), "wb").write(b"".join(b">" + r for r in recs[lo:hi]))
        U.run(cpu_bin, tmp, "--quiet", "--threads", "2")
        mine = {f: open(os.path.join(tmp, f), "rb").read() for f in U.FILES}
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        if rank == 0:
            got = {f: b"".join(g[f] for g in gathered) for f in U.FILES}
            ok = all(hashlib.md5(got[f]).hexdigest() == exp[f]["md5"] and len(got[f]) == exp[f]["bytes"] for f in U.FILES)
            out_q.put(ok)
        dist.barrier()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        dist.destroy_process_group()


def test_two_ranks_concatenate_to_the_single_run():
    cpu_bin = U.build_cpu_binary()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, cpu_bin, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
