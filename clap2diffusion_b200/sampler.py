"""The denoising loop: audio conditioning -> (UNet + CFG + scheduler) x steps -> VAE decode.

Wiring: clap [B,512] -> ImprovedHierarchicalAudioEncoder.encode -> routed
{early,mid,late} -> AudioAttnProcessor on the 16 attn2 sites; CFG pairs are batched as cat[uncond, cond]
on ONE GPU with the same audio on both halves.  Everything step-invariant (audio side, attn2 K/V, the
time-embedding table, scheduler coefficients) is computed before the loop; one step (UNet forward +
fused CFG/scheduler update writing the next UNet input) is captured in a CUDA graph and replayed, the
per-step scalars being fed through two tiny device-to-device copies.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib, ops
from .schedulers import make_plan
from .unet import SD15UNet
from .vae import VAEDecoder


class Sampler:
    def __init__(self, unet: SD15UNet, hier, vae: Optional[VAEDecoder] = None, use_graph: bool = True):
        self.unet, self.hier, self.vae = unet, hier, vae
        self.use_graph = use_graph
        self._graphs: Dict[tuple, dict] = {}
        self._plans: Dict[tuple, tuple] = {}          # (scheduler, steps) -> (plan, time-embedding table, coefficients)
        self._vae_graphs: Dict[tuple, dict] = {}      # latent shape -> captured VAE decode
        self.replayed_launches = 0      # kernels launched through graph replays (not seen by c2d_launch_count)

    # ------------------------------------------------------------------ conditioning (once per batch)
    @torch.no_grad()
    def condition(self, clap: torch.Tensor, ctx_cond: torch.Tensor, ctx_uncond: torch.Tensor,
                  use_audio: bool = True) -> Dict[str, torch.Tensor]:
        """Returns the per-site cached K/V for the CFG-doubled batch [uncond ; cond]."""
        ctx2 = torch.cat([ctx_uncond, ctx_cond], dim=0).contiguous()
        kwargs = None
        if use_audio and self.hier is not None:
            # the audio side runs in the UNet's compute dtype: bf16 -> its projector / decomposer GEMMs go to the
            # tcgen05 kernel; the fp32 parity mode keeps everything on the FFMA kernels
            a = clap.contiguous() if clap.dtype == self.unet.dtype else ops.cast(clap.contiguous(), self.unet.dtype)
            enc = self.hier.encode(a, with_tokens77=False)
            routed2 = {k: torch.cat([v, v], dim=0).contiguous() for k, v in enc["routed"].items()}
            kwargs = {"audio": routed2}
        return self.unet.prepare_conditioning(ctx2, kwargs)

    # ------------------------------------------------------------------ one step (graph body)
    def _step(self, st: dict) -> None:
        eps2 = self.unet.forward_nhwc(st["xin2"], st["temb_row"], st["kv"])
        ops.cfg_sched_step(eps2, st["x"], st["xin2"], st["guidance"], st["coef"])

    def _state(self, B: int, H: int, W: int, kv, guidance: float) -> dict:
        dev, dt = self.unet.device, self.unet.dtype
        return dict(x=torch.empty(B, 4, H, W, device=dev, dtype=torch.float32),
                    # bf16: 4 real + 4 zero channels so that conv_in is a tcgen05 implicit GEMM (16-byte TMA rows)
                    xin2=torch.zeros(2 * B, H, W, 8 if self.unet.conv_in_tc else 4, device=dev, dtype=dt),
                    temb_row=torch.empty(self.unet._temb_total, device=dev, dtype=torch.float32),
                    coef=torch.empty(3, device=dev, dtype=torch.float32), kv=kv, guidance=float(guidance), graph=None)

    # ------------------------------------------------------------------ the loop
    @torch.no_grad()
    def sample(self, clap: torch.Tensor, ctx_cond: torch.Tensor, ctx_uncond: torch.Tensor, noise: torch.Tensor,
               steps: int = 50, guidance: float = 7.5, scheduler: str = "ddim", use_audio: bool = True,
               decode: bool = True, trace: bool = False, max_steps: Optional[int] = None) -> Dict[str, object]:
        """clap [B,512], ctx_* [B,77,768], noise fp32 [B,4,H,W] -- all on the UNet's device.
        Returns dict(latents [B,4,H,W] fp32, image [B,3,8H,8W] fp32 if decode, trace [list] if trace)."""
        unet = self.unet
        B, _, H, W = noise.shape
        # the schedule's constants (scheduler coefficients, the frozen time-embedding rows of its timesteps) depend on
        # (scheduler, steps) only: built once, not once per image
        pk = (scheduler, int(steps))
        if pk not in self._plans:
            plan = make_plan(scheduler, steps)
            self._plans[pk] = (plan, unet.time_table(plan.timesteps), torch.from_numpy(plan.coef).to(unet.device))
        plan, table, coefs = self._plans[pk]
        n_run = steps if max_steps is None else min(steps, max_steps)
        kv = self.condition(clap, ctx_cond, ctx_uncond, use_audio)
        # static state (and the captured graph) is cached per problem shape; fresh K/V are copied into the
        # buffers the graph was captured on
        key = (B, H, W, float(guidance), tuple(sorted((n, tuple(t.shape)) for n, t in kv.items() if t is not None)))
        st = self._graphs.get(key)
        if st is None:
            st = self._state(B, H, W, kv, guidance)
            self._graphs[key] = st
        else:
            for n, t in kv.items():
                if t is not None:
                    st["kv"][n].copy_(t)
        # x0 = noise * init_scale ; first UNet input = cat[x0, x0] * first_in_scale : reuse the fused kernel with a
        # zero "eps" (ca = init_scale, cb = 0) so no extra elementwise kernels are needed
        st["x"].copy_(noise)
        zero_eps = torch.zeros(2 * B, H, W, 4, device=unet.device, dtype=unet.dtype)
        st["coef"].copy_(torch.tensor([plan.init_scale, 0.0, plan.first_in_scale], dtype=torch.float32))
        ops.cfg_sched_step(zero_eps, st["x"], st["xin2"], 0.0, st["coef"])
        traces: List[torch.Tensor] = []
        for i in range(n_run):
            st["temb_row"].copy_(table[i])
            st["coef"].copy_(coefs[i])
            if self.use_graph:
                if st["graph"] is None:
                    st["graph"] = self._capture(st)
                st["graph"].replay()
                self.replayed_launches += st["launches_per_replay"]
            else:
                self._step(st)
            if trace:
                traces.append(st["x"].clone())
        out: Dict[str, object] = {"latents": st["x"].clone()}
        if trace:
            out["trace"] = traces
        if decode and self.vae is not None:
            out["image"] = self._decode(out["latents"])
        return out

    def _decode(self, latents: torch.Tensor) -> torch.Tensor:
        """VAE decode; replayed from a CUDA graph per latent shape (~150 eager launches cost ~3 ms of host time per image
        batch, which is all exposed in the single-image latency)."""
        if not self.use_graph:
            return self.vae.decode(latents)
        key = tuple(latents.shape)
        st = self._vae_graphs.get(key)
        if st is None:
            z = latents.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.vae.decode(z)                       # lazy attribute / workspace setup must not happen during capture
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g):
                img = self.vae.decode(z)
            st = self._vae_graphs[key] = {"z": z, "graph": g, "img": img, "launches": _lib.launch_count() - n0}
        st["z"].copy_(latents)
        st["graph"].replay()
        self.replayed_launches += st["launches"]
        return st["img"].clone()

    def _capture(self, st: dict):
        """Warm up once on a side stream (lazy attribute / workspace setup must not happen during capture),
        restore the state it advanced, then capture one step."""
        x0, xin0 = st["x"].clone(), st["xin2"].clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step(st)
        torch.cuda.current_stream().wait_stream(side)
        st["x"].copy_(x0)
        st["xin2"].copy_(xin0)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(g):
            self._step(st)
        st["launches_per_replay"] = _lib.launch_count() - n0     # kernels one replay launches
        st["x"].copy_(x0)          # capture does not execute; keep the state untouched anyway
        st["xin2"].copy_(xin0)
        return g
