"""User-facing pipeline: host buffers in, decoded images out, one process per GPU.

``AudioToImagePipeline`` assembles the UNet / VAE engines and the drop-in audio modules from state dicts
(checkpoints or random init) and exposes ``generate`` -- the end-to-end call that bench.py's ``e2e`` number
times: host (pinned) inputs -> H2D -> conditioning -> 50 x (UNet + CFG + scheduler) -> VAE decode -> D2H.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops, synthetic
from . import unet as unet_mod
from . import vae as vae_mod
from .models.audio_adapter_v4 import AudioAdapter
from .models.audio_attention_processor import AudioProcessorManager
from .models.hierarchical_audio_v4 import ImprovedHierarchicalAudioEncoder
from .sampler import Sampler


class AudioToImagePipeline:
    def __init__(self, unet_sd: Dict[str, torch.Tensor], vae_sd: Optional[Dict[str, torch.Tensor]] = None,
                 hier_sd: Optional[Dict[str, torch.Tensor]] = None, proc_sd: Optional[Dict[str, Dict]] = None,
                 adapter_sd: Optional[Dict[str, torch.Tensor]] = None, device="cuda", dtype=torch.bfloat16,
                 mode: str = "add", use_graph: bool = True, impl: int = ops.IMPL_AUTO):
        self.device, self.dtype = torch.device(device), dtype
        self.unet = unet_mod.SD15UNet(unet_sd, device=device, dtype=dtype, impl=impl)
        self.vae = vae_mod.VAEDecoder(vae_sd, device=device, dtype=dtype, impl=impl) if vae_sd is not None else None
        self.hier = ImprovedHierarchicalAudioEncoder().to(self.device).eval()
        if hier_sd is not None:
            self.hier.load_state_dict({k: v.to(self.device) for k, v in hier_sd.items()})
        self.adapter = None
        if adapter_sd is not None:
            self.adapter = AudioAdapter().to(self.device).eval()
            self.adapter.load_state_dict({k: v.to(self.device) for k, v in adapter_sd.items()})
        self.manager = AudioProcessorManager(self.unet)
        self.manager.setup_processors(mode=mode)
        if proc_sd is not None:
            for lvl, names in self.manager.level_mapping.items():
                if names and lvl in proc_sd:
                    proc = self.unet.sites[names[0][:-len(".processor")]].processor
                    proc.load_state_dict({k: v.to(self.device) for k, v in proc_sd[lvl].items()})
        self.sampler = Sampler(self.unet, self.hier, self.vae, use_graph=use_graph)

    @classmethod
    def random_init(cls, seed: int = 0, device="cuda", dtype=torch.bfloat16, with_vae: bool = True, **kw):
        """Random-init weights of the named architecture (no checkpoints exist on the box)."""
        torch.manual_seed(seed)
        usd = synthetic.random_state_dict(unet_mod.param_shapes(), seed, device)
        vsd = synthetic.random_state_dict(vae_mod.param_shapes(), seed, device) if with_vae else None
        pipe = cls(usd, vsd, device=device, dtype=dtype, **kw)
        return pipe

    # ------------------------------------------------------------------ end-to-end call
    @torch.no_grad()
    def generate(self, clap: np.ndarray, ctx_cond: np.ndarray, ctx_uncond: np.ndarray, noise: np.ndarray,
                 steps: int = 50, guidance: float = 7.5, scheduler: str = "ddim", decode: bool = True,
                 trace: bool = False, max_steps: Optional[int] = None) -> Dict[str, object]:
        """HOST arrays in (clap [B,512], ctx [B,77,768], noise [B,4,h,w], fp32), HOST arrays out.
        Inputs go through pinned staging buffers; the result read-back is the only synchronisation."""
        dev = self.device

        def h2d(a, dtype):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory().to(dev, non_blocking=True)
            return t if dtype == torch.float32 else ops.cast(t, dtype)

        out = self.sampler.sample(h2d(clap, torch.float32), h2d(ctx_cond, self.dtype), h2d(ctx_uncond, self.dtype),
                                  h2d(noise, torch.float32), steps=steps, guidance=guidance, scheduler=scheduler,
                                  decode=decode, trace=trace, max_steps=max_steps)
        res: Dict[str, object] = {"latents": out["latents"].cpu().numpy()}
        if "image" in out:
            res["image"] = out["image"].cpu().numpy()
        if trace:
            res["trace"] = [t.cpu().numpy() for t in out["trace"]]
        return res

    @torch.no_grad()
    def generate_from_waves(self, clap_encoder, waves: np.ndarray, ctx_cond: np.ndarray, ctx_uncond: np.ndarray,
                            noise: np.ndarray, **kw) -> Dict[str, object]:
        """The whole path of the north star from HOST waveforms: waves [B, 480000] fp32 (10 s @ 48 kHz) -> pinned H2D ->
        GPU log-mel + HTSAT tower (models.audio_encoder.CLAPAudioEncoder) -> generate().  The CLAP embedding never
        leaves the device."""
        dev = self.device
        w = torch.from_numpy(np.ascontiguousarray(waves, dtype=np.float32)).pin_memory().to(dev, non_blocking=True)
        clap = clap_encoder.encode_audio(w)

        def h2d(a, dtype):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory().to(dev, non_blocking=True)
            return t if dtype == torch.float32 else ops.cast(t, dtype)

        out = self.sampler.sample(clap.float().contiguous(), h2d(ctx_cond, self.dtype), h2d(ctx_uncond, self.dtype),
                                  h2d(noise, torch.float32), **kw)
        res: Dict[str, object] = {"latents": out["latents"].cpu().numpy()}
        if "image" in out:
            res["image"] = out["image"].cpu().numpy()
        return res

    @staticmethod
    def io_bytes(B: int, h: int = 64, w: int = 64, decode: bool = True):
        """(h2d, d2h) bytes per generate() call, counted from the tensors copied."""
        h2d = 4 * (B * 512 + 2 * B * 77 * 768 + B * 4 * h * w)
        d2h = 4 * B * 4 * h * w + (4 * B * 3 * 8 * h * 8 * w if decode else 0)
        return h2d, d2h


def jobs_for_rank(n_jobs: int, rank: int, world: int) -> List[int]:
    """Contiguous block partition of job indices (prompt x seed) over ranks."""
    base, rem = divmod(n_jobs, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def gather_latents(local: torch.Tensor, counts: Sequence[int], group=None) -> Optional[torch.Tensor]:
    """All-gather final latents [n_local,4,h,w] from every rank (the path's only collective; NCCL on GPU,
    gloo in the CPU tests).  Ragged counts are padded to the max and trimmed.  Returns the concatenation
    in rank order on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nmax = max(counts)
    pad = torch.zeros(nmax, *local.shape[1:], device=local.device, dtype=local.dtype)
    pad[: local.shape[0]].copy_(local)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
