"""Data-parallel sharding: contiguous job blocks per rank, one all-gather of final latents.
world_size-2 gloo on CPU covers the N>1 host path; results must not depend on the partition."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clap2diffusion_b200.pipeline import gather_latents, jobs_for_rank


def test_jobs_partition_covers_everything():
    for n, w in ((64, 1), (64, 2), (64, 8), (7, 2), (5, 8), (0, 4)):
        parts = [jobs_for_rank(n, r, w) for r in range(w)]
        flat = [j for p in parts for j in p]
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, n_jobs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = jobs_for_rank(n_jobs, rank, world)
    # "latents" are a pure function of the job id, as in the real pipeline (per-image host noise)
    local = torch.stack([torch.full((4, 8, 8), float(j)) for j in mine]) if mine else torch.zeros(0, 4, 8, 8)
    counts = [len(jobs_for_rank(n_jobs, r, world)) for r in range(world)]
    full = gather_latents(local, counts)
    if rank == 0:
        q.put(full[:, 0, 0, 0].tolist())
    dist.destroy_process_group()


def test_gather_latents_world2_ragged():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got == [float(j) for j in range(7)]
