"""Data-parallel stage-3 step on CPU, world size 2 over gloo: each rank differentiates half of the batch, the per-level
all-reduce buckets make both ranks hold the full-batch gradient (loss pre-scaled by 1 / world), and both take the same
optimiser step.  libc2d ops are the torch doubles (host logic only; the kernels are covered by tests/test_gpu_train.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch_ops
    from oracle import pipeline as PL
    from test_train_host_logic import make_batch
    from clap2diffusion_b200.models.hierarchical_audio_v4 import ImprovedHierarchicalAudioEncoder
    from clap2diffusion_b200.train import LEVELS, Stage3Trainer
    torch.set_num_threads(2)
    W = PL.build_weights(seed=0, with_vae=False)
    full = make_batch(B=2, h=8, w=8, seed=5)
    mine = {k: v[rank:rank + 1] for k, v in full.items()}

    def trainer(**kw):
        hier = ImprovedHierarchicalAudioEncoder().eval()
        hier.load_state_dict(W["hier"])
        return Stage3Trainer(W["unet"], hier, {l: W[f"proc_{l}"] for l in LEVELS}, device="cpu", dtype=torch.float32,
                             learning_rate=1e-3, num_steps=10, **kw)
    with torch_ops.installed(), torch.no_grad():
        tr = trainer()
        assert tr.world == world
        loss = tr.forward_backward(mine["audio_embedding"], mine["image_latents"], mine["text_embedding"], mine["noise"], mine["timesteps"])
        g = tr.grad.clone()
        tr.optimizer_step()
        out = {"grad": g.numpy(), "flat": tr.flat.clone().numpy(), "loss": float(loss)}
        if rank == 0:          # single-process reference on the full batch
            dist_world = tr.world
            ref = trainer(data_parallel=False)         # no collectives: rank 1 does not build one
            assert ref.world == 1
            ref_loss = ref.forward_backward(full["audio_embedding"], full["image_latents"], full["text_embedding"], full["noise"], full["timesteps"])
            out["ref_grad"] = ref.grad.clone().numpy()
            ref.optimizer_step()
            out["ref_flat"], out["ref_loss"], out["world"] = ref.flat.clone().numpy(), float(ref_loss), dist_world
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_stage3_data_parallel_world2_gloo():
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = res[0], res[1]
    assert np.array_equal(r0["grad"], r1["grad"]) and np.array_equal(r0["flat"], r1["flat"])      # replicas stay identical
    # mean over ranks of the per-rank mse == mse over the full batch (equal shard sizes)
    assert abs(r0["loss"] + r1["loss"] - r0["ref_loss"]) < 1e-5 * abs(r0["ref_loss"])
    e = np.linalg.norm(r0["grad"] - r0["ref_grad"]) / np.linalg.norm(r0["ref_grad"])
    assert e < 1e-4, e
    step = np.linalg.norm(r0["ref_flat"] - r0["flat"]) / np.linalg.norm(r0["ref_flat"])
    assert step < 1e-4
