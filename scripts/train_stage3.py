#!/usr/bin/env python
"""Stage-3 fine-tuning on the B200 path -- the counterpart of the reference's ``scripts/train_stage3.py``.

The reference's script (``Stage3Trainer`` :22-272, ``main`` :274-330) describes the step -- AdamW(lr 1e-5, weight decay
0.01), CosineAnnealingLR(eta_min 1e-6), ``clip_grad_norm_(0.5)``, ``mse(predicted_noise, noise) * 2.0`` on
``alphas * latents + (1 - alphas) * noise`` -- around a placeholder noise predictor, and its ``main`` loop only simulates
steps.  Here the same configuration keys drive the real step of ``clap2diffusion_b200.train.Stage3Trainer``: the frozen
SD-1.5 UNet is differentiated by hand on libc2d kernels, the audio attention processors are updated, and the whole step
replays from one CUDA graph.  Checkpoint names follow the reference (``checkpoint_dir`` / ``hierarchical_v4_final.pth``,
``unet_adapter_final.pth``; ``unet.pth`` for the SD-1.5 UNet state dict -- without one the UNet is random-init, a synthetic
run); the result is written as ``unet_adapter_final.pth`` in the layout ``scripts/inference.py`` loads.

    python scripts/train_stage3.py --num-steps 1000 --batch-size 4
    python -m torch.distributed.run --nproc-per-node 8 scripts/train_stage3.py --batch-size 4        # global batch 32

Batches: ``--data DIR`` holds ``*.pt`` files, each a dict with the reference's batch keys (``audio_embedding`` [B,512],
``image_latents`` [B,4,h,w], ``text_embedding`` [B,77,768], :134-136); without it the run uses synthetic batches
(there are no datasets on the box).  Noise and timesteps are drawn per step from a seeded generator (:157-158).
"""
from __future__ import annotations

import argparse
import glob
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from clap2diffusion_b200 import synthetic  # noqa: E402
from clap2diffusion_b200 import unet as unet_mod  # noqa: E402
from clap2diffusion_b200.models.hierarchical_audio_v4 import HierarchicalAudioV4, ImprovedHierarchicalAudioEncoder  # noqa: E402
from clap2diffusion_b200.train import LEVELS, Stage3Trainer  # noqa: E402

# the reference's defaults (main :275-286); batch_size is per GPU here
DEFAULTS = {"learning_rate": 1e-5, "weight_decay": 0.01, "num_steps": 1000, "batch_size": 2, "gradient_clipping": 0.5,
            "checkpoint_dir": "../checkpoints", "save_interval": 500, "log_interval": 50}


def load_models(checkpoint_dir: Path, device, seed: int):
    """(unet state dict, hierarchical model, processor state dicts or None) from the reference's file names."""
    def load(name):
        path = checkpoint_dir / name
        if not path.exists():
            return None
        print(f"Loading {name} from {path}")
        return torch.load(path, map_location="cpu", weights_only=True)

    unet_sd = load("unet.pth")
    if unet_sd is None:
        print("No SD-1.5 UNet checkpoint (unet.pth): random-init UNet (synthetic run)")
        unet_sd = synthetic.random_state_dict(unet_mod.param_shapes(), seed, device)
    adapter = load("unet_adapter_final.pth") or {}
    procs = {lvl: adapter[f"processor_{lvl}"] for lvl in LEVELS} if all(f"processor_{lvl}" in adapter for lvl in LEVELS) else None
    torch.manual_seed(seed)                      # random-init parts are identical on every rank
    if "hierarchical_state_dict" in adapter:
        hier = ImprovedHierarchicalAudioEncoder().to(device).eval()
        hier.load_state_dict({k: v.to(device) for k, v in adapter["hierarchical_state_dict"].items()})
    else:
        legacy = load("hierarchical_v4_final.pth")
        if legacy is not None:                   # the reference's own stage-1/2 checkpoint: rigid 5-3-2 decomposition
            hier = HierarchicalAudioV4().to(device).eval()
            hier.load_state_dict(legacy, strict=False)
        else:
            hier = ImprovedHierarchicalAudioEncoder().to(device).eval()
    return unet_sd, hier, procs


class Batches:
    """Reference-format batches from ``*.pt`` files, or synthetic ones; noise / timesteps drawn per step."""

    def __init__(self, data_dir, batch_size, latent, rank, seed):
        self.files = sorted(glob.glob(os.path.join(data_dir, "*.pt"))) if data_dir else []
        self.b, self.latent, self.rank = batch_size, latent, rank
        self.gen = torch.Generator().manual_seed(seed * 1000 + rank)

    def __call__(self, step: int):
        if self.files:
            d = torch.load(self.files[(step * 131 + self.rank) % len(self.files)], map_location="cpu", weights_only=True)
            batch = {k: d[k][:self.b] for k in ("audio_embedding", "image_latents", "text_embedding")}
        else:
            ids = [(self.rank * 100003 + step) * self.b + j for j in range(self.b)]
            batch = {"audio_embedding": torch.from_numpy(np.stack([synthetic.clap_embedding(k) for k in ids])),
                     "image_latents": torch.randn(self.b, 4, self.latent, self.latent, generator=self.gen),
                     "text_embedding": torch.from_numpy(np.stack([synthetic.text_states(f"prompt {k % 8}") for k in ids]))}
        n = batch["image_latents"].shape[0]
        batch["noise"] = torch.randn(batch["image_latents"].shape, generator=self.gen)
        batch["timesteps"] = torch.randint(0, 1000, (n,), generator=self.gen)
        return batch


def save_checkpoint(trainer: Stage3Trainer, hier, config, save_dir: Path):
    """``unet_adapter_final.pth`` (processors; what inference.py reads) and the reference's stage-3 file name with its keys."""
    save_dir.mkdir(parents=True, exist_ok=True)
    sd = trainer.state_dict()
    torch.save({k: v for k, v in sd.items() if k != "optimizer_state_dict"}, save_dir / "unet_adapter_final.pth")
    torch.save({"step": sd["step"], "hierarchical_state_dict": {k: v.detach().cpu() for k, v in hier.state_dict().items()},
                "optimizer_state_dict": sd["optimizer_state_dict"], "config": dict(config)},
               save_dir / "audio_projector_stage3_finetuned.pth")
    print(f"Saved Stage 3 checkpoint to {save_dir}")


def main(argv=None):
    ap = argparse.ArgumentParser(description="CLAP2Diffusion stage-3 fine-tuning (audio attention processors; frozen SD-1.5 UNet)")
    for k, v in DEFAULTS.items():
        ap.add_argument("--" + k.replace("_", "-"), type=type(v), default=v)
    ap.add_argument("--data", default=None, help="directory of *.pt batches in the reference's format (default: synthetic)")
    ap.add_argument("--latent", type=int, default=64, help="latent height / width of synthetic batches")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--output-dir", default=None, help="where the fine-tuned checkpoints go (default: checkpoint_dir)")
    args = ap.parse_args(argv)
    config = {k: getattr(args, k) for k in DEFAULTS}
    if not torch.cuda.is_available():
        raise SystemExit("train_stage3.py: no CUDA device; the compute library has no CPU path")
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    log = print if rank == 0 else (lambda *a, **k: None)
    log("=" * 60 + "\nStage 3 Training: Final Fine-tuning\n" + "=" * 60)
    for k, v in config.items():
        log(f"  {k}: {v}")
    unet_sd, hier, procs = load_models(Path(args.checkpoint_dir), device, args.seed)
    trainer = Stage3Trainer(unet_sd, hier, procs, device=device, dtype=torch.bfloat16 if args.dtype == "bf16" else torch.float32,
                            learning_rate=args.learning_rate, weight_decay=args.weight_decay, num_steps=args.num_steps,
                            gradient_clipping=args.gradient_clipping)
    log(f"  trainable parameters: {trainer.num_params:,} (audio attention processors, {len(LEVELS)} levels); "
        f"{world} GPU(s) x {args.batch_size} samples")
    batches = Batches(args.data, args.batch_size, args.latent, rank, args.seed)
    if not args.no_graph:
        trainer.capture({k: v.to(device) for k, v in batches(0).items()})
    out_dir = Path(args.output_dir) if args.output_dir else Path(args.checkpoint_dir)
    t0, window = time.time(), []
    for step in range(args.num_steps):
        out = trainer.train_step({k: v.to(device, non_blocking=True) for k, v in batches(step).items()})
        window.append(out)
        last = step + 1 == args.num_steps
        if (step + 1) % args.log_interval == 0 or last:
            loss = float(torch.stack([o["diffusion"] for o in window]).mean()) * world     # one host read per log interval
            log(f"step {step + 1:6d}  diffusion loss {loss:.5f}  grad norm {float(window[-1]['grad_norm']):.3e}  "
                f"lr {trainer.lr(step):.2e}  {len(window) * args.batch_size * world / (time.time() - t0):.1f} samples/s")
            t0, window = time.time(), []
        if rank == 0 and ((step + 1) % args.save_interval == 0 or last):
            save_checkpoint(trainer, hier, config, out_dir)
    log("\nStage 3 fine-tuning complete!")
    if world > 1:
        trainer.release_graph()              # a live graph holding NCCL kernels keeps the communicator's teardown waiting
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return trainer


if __name__ == "__main__":
    main()
