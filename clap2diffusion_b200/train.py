"""Stage-3 fine-tune step on B200 (BASELINE.json configs[4]): the audio attention processors are
trainable, the SD-1.5 UNet is frozen, data-parallel over the GPUs of one box.

What the reference has (scripts/train_stage3.py:132-191): a ``train_step`` whose diffusion term is
``mse(predict_noise_simple(...), noise) * 2.0`` around a PLACEHOLDER predictor ("in real implementation, would use
actual UNet", :196-207), AdamW(lr 1e-5, weight decay 0.01) with CosineAnnealingLR(eta_min 1e-6) (:33-46) and
``clip_grad_norm_(0.5)`` (:182-186).  As written it cannot run (``HierarchicalAudioV4`` is constructed with keyword
arguments the class does not take) and nothing trainable receives a gradient (the tokens are
re-scaled under ``no_grad``, :193-200).  This module is the step the reference describes with the placeholder filled
in by the real frozen UNet:

    noisy = a * latents + (1 - a) * noise,  a = 1 - t / 1000                       (the reference's noising, :203-204)
    eps   = UNet(noisy, t, text states with the audio injected by the AudioAttnProcessors)
    loss  = 2.0 * mse(eps, noise)                                                   (:166)
    grads of the processors' audio_proj.0 / audio_proj.3 / alpha (3 levels, 297,411 parameters)
    all-reduce (one bucket per level, launched as soon as the level's last site has been differentiated, so the
    collective overlaps the rest of the backward pass) -> global-norm clip 0.5 -> AdamW -> cosine learning rate.

The other terms of the reference's loss dict (consistency :168-171, alignment :173-175, and the stage-2 regularisers of
``compute_losses``, models/hierarchical_audio_v4.py:661-711) are functions of the frozen projector / decomposer outputs
only: they carry no gradient to the trainable set and are not part of the step.

Compute: with the UNet frozen the backward pass is activation gradients only.  Every dense layer's adjoint is the
forward kernel again (tcgen05 in bf16, FFMA in the fp32 parity mode) on a transposed / flipped copy of the frozen
weight; GroupNorm, LayerNorm, GEGLU, attention, the loss and the optimiser have their own kernels (csrc/train.cu).
PyTorch provides memory, streams and ``torch.distributed`` -- there is no autograd anywhere in this file.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch

from . import ops
from .models.audio_attention_processor import AudioAttnProcessor, AudioProcessorManager
from .unet import BLOCK_OUT, GROUPS, SD15UNet

LEVELS = ("early", "mid", "late")


class _Tape:
    """Reverse-mode bookkeeping: adjoint closures in forward order, gradients keyed by the activation's storage."""

    def __init__(self):
        self.ops: List[Callable[[], None]] = []
        self.grads: Dict[int, torch.Tensor] = {}
        self.stop = set()

    def add_grad(self, t: torch.Tensor, g: torch.Tensor) -> None:
        k = t.data_ptr()
        if k in self.stop:
            return
        cur = self.grads.get(k)
        self.grads[k] = g if cur is None else ops.add(cur, g.reshape(cur.shape))

    def pop(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        return self.grads.pop(t.data_ptr(), None)

    def no_grad(self, t: torch.Tensor) -> None:
        self.stop.add(t.data_ptr())

    def run(self) -> None:
        for fn in reversed(self.ops):
            fn()
        self.ops.clear()
        self.grads.clear()


class FrozenUNetGrad:
    """SD-1.5 UNet forward that keeps what the adjoint needs, and the adjoint pass down to the K / V of every attn2 site."""

    def __init__(self, unet: SD15UNet):
        if unet.fused_gn or unet.fold_ln or unet.fuse_geglu or unet.fused_xattn:
            raise ValueError("FrozenUNetGrad needs SD15UNet(..., fused=False): one kernel per layer, plain weights")
        self.u, self.w = unet, unet.w
        self._bw: Dict[str, torch.Tensor] = {}
        self.kv_grad_hook: Optional[Callable[[str, torch.Tensor], None]] = None

    # ---------------------------------------------------------------- adjoint weights (built once, frozen)
    def _wt(self, key: str, w: Optional[torch.Tensor] = None) -> torch.Tensor:
        t = self._bw.get(key)
        if t is None:
            t = (self.w[key] if w is None else w).t().contiguous()
            self._bw[key] = t
        return t

    def _wconv(self, key: str) -> torch.Tensor:
        """[Cout][3][3][Cin] -> [Cin][3][3][Cout] with both taps flipped: the data-gradient of a pad-1 3x3 convolution is
        the same convolution with this kernel."""
        t = self._bw.get(key)
        if t is None:
            t = self.w[key].flip(1, 2).permute(3, 1, 2, 0).contiguous()
            self._bw[key] = t
        return t

    # ---------------------------------------------------------------- layers (forward + recorded adjoint)
    def _gn(self, tp: _Tape, x, name, eps, silu):
        g, b = self.w[f"{name}.weight"], self.w[f"{name}.bias"]
        C = x.shape[-1]
        stats = None
        if x.dtype == torch.bfloat16 and C % 8 == 0 and C <= 4096:
            # one statistics pass serves the forward (one-pass apply, the inference kernels) AND the adjoint
            stats = ops.channel_stats(x, torch.zeros(x.shape[0] * C * 2, device=x.device, dtype=torch.int64))
            y = ops.group_norm_apply(x, stats, g, b, GROUPS, eps, silu)
        else:
            y = ops.group_norm(x, g, b, GROUPS, eps, silu)

        def bwd():
            gy = tp.pop(y)
            if gy is not None and x.data_ptr() not in tp.stop:
                tp.add_grad(x, ops.group_norm_bwd(x, gy.reshape(x.shape).contiguous(), g, b, GROUPS, eps, silu, stats=stats))
        tp.ops.append(bwd)
        return y

    def _ln(self, tp: _Tape, x, name):
        g, b = self.w[f"{name}.weight"], self.w[f"{name}.bias"]
        y = ops.layer_norm(x, g, b)

        def bwd():
            gy = tp.pop(y)
            if gy is not None:
                tp.add_grad(x, ops.layer_norm_bwd(x, gy.reshape(x.shape).contiguous(), g))
        tp.ops.append(bwd)
        return y

    def _conv(self, tp: _Tape, x, name, *, rowvec=None, bias=True, residual=None, stride=1, upsample=False):
        w = self.w[f"{name}.weight"]
        xin = ops.upsample2x(x) if upsample else x
        y = ops.conv3x3(xin, w, self.w[f"{name}.bias"] if bias else None, rowvec=rowvec, residual=residual, stride=stride,
                        impl=self.u.impl)

        def bwd():
            gy = tp.pop(y)
            if gy is None:
                return
            gy = gy.reshape(y.shape)
            if residual is not None:
                tp.add_grad(residual, gy)
            if x.data_ptr() in tp.stop:
                return
            gin = ops.zero_insert2x(gy) if stride == 2 else gy
            gx = ops.conv3x3(gin, self._wconv(f"{name}.weight"), None, impl=self.u.impl)
            tp.add_grad(x, ops.sumpool2x2(gx) if upsample else gx)
        tp.ops.append(bwd)
        return y

    def _lin(self, tp: _Tape, x, wkey, bkey=None, *, residual=None, w=None):
        wt = self.w[wkey] if w is None else w
        y = ops.linear(x, wt, self.w[bkey] if bkey else None, residual=residual, impl=self.u.impl)

        def bwd():
            gy = tp.pop(y)
            if gy is None:
                return
            gy = gy.reshape(y.shape)
            if residual is not None:
                tp.add_grad(residual, gy)
            if x.data_ptr() not in tp.stop:
                tp.add_grad(x, ops.linear(gy, self._wt(wkey, w), impl=self.u.impl))
        tp.ops.append(bwd)
        return y

    def _resnet(self, tp: _Tape, name, x, temb_rows, skip=None):
        w = self.w
        cout = w[f"{name}.conv1.weight"].shape[0]
        off = self.u._temb_offsets[name]
        if skip is not None:
            c1, c2 = x.shape[-1], skip.shape[-1]
            a, s = x, skip
            x = ops.concat(a, s)

            def bwd_cat(x=x, a=a, s=s, c1=c1, c2=c2):
                g = tp.pop(x)
                if g is not None:
                    tp.add_grad(a, ops.slice_channels(g, 0, c1))
                    tp.add_grad(s, ops.slice_channels(g, c1, c2))
            tp.ops.append(bwd_cat)
        h = self._gn(tp, x, f"{name}.norm1", 1e-5, True)
        # conv1.bias is folded into the time-embedding row (SD15UNet._pack_resnet); one row per sample
        h = self._conv(tp, h, f"{name}.conv1", rowvec=temb_rows[:, off:off + cout].contiguous(), bias=False)
        h = self._gn(tp, h, f"{name}.norm2", 1e-5, True)
        sc = x
        if f"{name}.conv_shortcut.weight" in w:
            sc = self._lin(tp, x, f"{name}.conv_shortcut.weight", f"{name}.conv_shortcut.bias")
        return self._conv(tp, h, f"{name}.conv2", residual=sc)

    def _self_attention(self, tp: _Tape, site, x):
        C, heads = x.shape[-1], site.heads
        qkv = self._lin(tp, x, f"{site.name}.wqkv", w=site.wqkv)
        # the long-sequence forward kernel also leaves the rows' log-sum-exp for the adjoint (one sweep over the keys less)
        lse = torch.empty(x.shape[0], heads, x.shape[1], device=x.device, dtype=torch.float32)
        o, have_lse = ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, scale=site.scale, lse=lse)

        def bwd():
            go = tp.pop(o)
            if go is None:
                return
            dqkv = torch.empty_like(qkv)
            ops.attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], o, go, heads,
                              dqkv[..., :C], dqkv[..., C:2 * C], dqkv[..., 2 * C:], scale=site.scale,
                              lse=lse if have_lse else None)
            tp.add_grad(qkv, dqkv)
        tp.ops.append(bwd)
        return o

    def _cross_attention(self, tp: _Tape, site, x, kv):
        C, heads = x.shape[-1], site.heads
        q = self._lin(tp, x, f"{site.name}.to_q", w=site.to_q.weight)
        o = ops.attention(q, kv[..., :C], kv[..., C:], heads, scale=site.scale)

        def bwd():
            go = tp.pop(o)
            if go is None:
                return
            dq, dkv = torch.empty_like(q), torch.empty_like(kv)
            ops.attention_bwd(q, kv[..., :C], kv[..., C:], o, go, heads, dq, dkv[..., :C], dkv[..., C:], scale=site.scale)
            tp.add_grad(q, dq)
            if self.kv_grad_hook is not None:
                self.kv_grad_hook(site.name, dkv)
        tp.ops.append(bwd)
        return o

    def _transformer(self, tp: _Tape, name, x, kv):
        w = self.w
        B, H, W, C = x.shape
        tb = f"{name}.transformer_blocks.0"
        s1, s2 = self.u.sites[f"{tb}.attn1"], self.u.sites[f"{tb}.attn2"]
        res = x.view(B, H * W, C)
        h = self._gn(tp, res, f"{name}.norm", 1e-6, False)
        h = self._lin(tp, h, f"{name}.proj_in.weight", f"{name}.proj_in.bias")
        a = self._self_attention(tp, s1, self._ln(tp, h, f"{tb}.norm1"))
        h = self._lin_site_out(tp, a, s1, h)
        a = self._cross_attention(tp, s2, self._ln(tp, h, f"{tb}.norm2"), kv[s2.name])
        h = self._lin_site_out(tp, a, s2, h)
        ag = self._lin(tp, self._ln(tp, h, f"{tb}.norm3"), f"{tb}.ff.net.0.proj.weight", f"{tb}.ff.net.0.proj.bias")
        f = ops.geglu(ag)

        def bwd_geglu():
            g = tp.pop(f)
            if g is not None:
                tp.add_grad(ag, ops.geglu_bwd(ag, g))
        tp.ops.append(bwd_geglu)
        h = self._lin(tp, f, f"{tb}.ff.net.2.weight", f"{tb}.ff.net.2.bias", residual=h)
        out = self._lin(tp, h, f"{name}.proj_out.weight", f"{name}.proj_out.bias", residual=res)
        return out.view(B, H, W, C)

    def _lin_site_out(self, tp: _Tape, a, site, residual):
        """to_out[0] of an attention site (+ residual)."""
        wo, bo = site.to_out[0].weight, site.to_out[0].bias
        y = ops.linear(a, wo, bo, residual=residual, impl=self.u.impl)

        def bwd():
            gy = tp.pop(y)
            if gy is None:
                return
            tp.add_grad(residual, gy)
            tp.add_grad(a, ops.linear(gy, self._wt(f"{site.name}.to_out", wo), impl=self.u.impl))
        tp.ops.append(bwd)
        return y

    # ---------------------------------------------------------------- forward
    def forward(self, tp: _Tape, x_nhwc: torch.Tensor, temb_rows: torch.Tensor, kv: Dict[str, torch.Tensor]) -> torch.Tensor:
        """x_nhwc [B,H,W,4] engine dtype; temb_rows fp32 [B, sum(Cout)] (SD15UNet.time_table, one row per sample);
        kv: attn2 site name -> [B,T,2C].  Returns eps [B,H,W,4]."""
        u, w = self.u, self.w
        h = ops.conv3x3(x_nhwc, w["conv_in.weight"], w["conv_in.bias"], impl=u.impl)
        tp.no_grad(h)                                   # nothing trainable upstream of the first cross-attention
        skips = [h]
        for i, blk in enumerate(u.down):
            for j in range(2):
                h = self._resnet(tp, f"down_blocks.{i}.resnets.{j}", h, temb_rows)
                if i == 0 and j == 0:
                    tp.no_grad(h)
                if blk["attn"]:
                    h = self._transformer(tp, f"down_blocks.{i}.attentions.{j}", h, kv)
                skips.append(h)
            if blk["sample"]:
                h = self._conv(tp, h, f"down_blocks.{i}.downsamplers.0.conv", stride=2)
                skips.append(h)
        h = self._resnet(tp, "mid_block.resnets.0", h, temb_rows)
        h = self._transformer(tp, "mid_block.attentions.0", h, kv)
        h = self._resnet(tp, "mid_block.resnets.1", h, temb_rows)
        for i, blk in enumerate(u.up):
            for j in range(3):
                h = self._resnet(tp, f"up_blocks.{i}.resnets.{j}", h, temb_rows, skip=skips.pop())
                if blk["attn"]:
                    h = self._transformer(tp, f"up_blocks.{i}.attentions.{j}", h, kv)
            if blk["sample"]:
                h = self._conv(tp, h, f"up_blocks.{i}.upsamplers.0.conv", upsample=True)
        B, H, W, C = h.shape
        h = self._gn(tp, h.view(B, H * W, C), "conv_norm_out", 1e-5, True)
        return self._conv(tp, h.view(B, H, W, C), "conv_out")


class Stage3Trainer:
    """``Stage3Trainer(unet_sd, hier, proc_sd).train_step(batch)``: one optimiser step of the audio attention processors
    through the frozen UNet.  One instance per GPU; ``group`` (optional) is the data-parallel process group (default: the
    world group when ``torch.distributed`` is initialised; ``data_parallel=False`` opts out)."""

    def __init__(self, unet_sd: Dict[str, torch.Tensor], hier, proc_sd: Optional[Dict[str, Dict[str, torch.Tensor]]] = None,
                 device="cuda", dtype=torch.bfloat16, mode: str = "add", learning_rate: float = 1e-5,
                 weight_decay: float = 0.01, num_steps: int = 3000, eta_min: float = 1e-6, gradient_clipping: float = 0.5,
                 betas=(0.9, 0.999), adam_eps: float = 1e-8, group=None, diffusion_weight: float = 2.0,
                 data_parallel: bool = True):
        if mode != "add":
            raise ValueError("the training step differentiates the 'add' injection (the reference's default mode)")
        self.device, self.dtype = torch.device(device), dtype
        self.unet = SD15UNet(unet_sd, device=device, dtype=dtype, fused=False)
        self.net = FrozenUNetGrad(self.unet)
        self.hier = hier
        self.manager = AudioProcessorManager(self.unet)
        self.manager.setup_processors(mode=mode, dropout=0.0)        # deterministic step: the injection MLP's dropout is off
        self.procs: Dict[str, AudioAttnProcessor] = {}
        self.level_of: Dict[str, str] = {}
        for lvl, names in self.manager.level_mapping.items():
            for n in names:
                site = n[:-len(".processor")]
                self.level_of[site] = lvl
                self.procs[lvl] = self.unet.sites[site].processor
        if proc_sd is not None:
            for lvl in LEVELS:
                self.procs[lvl].load_state_dict({k: v.to(self.device) for k, v in proc_sd[lvl].items()})
        # flat fp32 master parameters / gradients / Adam moments; the processors' nn.Parameters are views into `flat`
        # (every tensor starts on a 32-byte boundary: the vectorised kernels read 8 floats at a time; the padding elements
        # stay zero in parameters, gradients and moments)
        self.slots: Dict[str, Dict[str, slice]] = {}
        n = 0
        self.num_params = 0
        for lvl in LEVELS:
            self.slots[lvl] = {}
            for k, p in self.procs[lvl].named_parameters():
                n = (n + 7) & ~7
                self.slots[lvl][k] = slice(n, n + p.numel())
                n += p.numel()
                self.num_params += p.numel()
        n = (n + 7) & ~7
        self.flat = torch.zeros(n, device=self.device, dtype=torch.float32)
        for lvl in LEVELS:
            for k, p in self.procs[lvl].named_parameters():
                sl = self.slots[lvl][k]
                self.flat[sl].copy_(p.detach().reshape(-1))
                p.data = self.flat[sl].view(p.shape)
                p.requires_grad_(False)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.level_range = {lvl: slice(min(s.start for s in self.slots[lvl].values()), max(s.stop for s in self.slots[lvl].values()))
                            for lvl in LEVELS}
        self.lr0, self.eta_min, self.num_steps = learning_rate, eta_min, num_steps
        self.wd, self.clip, self.betas, self.adam_eps = weight_decay, gradient_clipping, betas, adam_eps
        self.diffusion_weight = diffusion_weight
        self.group = group
        self.world = 1
        # data_parallel=False: a stand-alone replica inside an initialised process group (no broadcast, no all-reduce)
        if data_parallel and (group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(group)
        if self.world > 1:
            # DDP semantics: every replica starts from rank 0's parameters (the processors' tensors are views of `flat`)
            torch.distributed.broadcast(self.flat, src=torch.distributed.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._time_table: Optional[torch.Tensor] = None      # [1000, sum(Cout)] rows of SD15UNet.time_table, built on first use
        self.step_count = 0
        # the optimiser's schedule lives on the device (row t: learning rate and the two bias corrections of step t) next to a
        # device step counter, so that a training step has no host-side scalar and can be replayed from ONE CUDA graph
        tt = range(num_steps + 1)
        self._sched = torch.tensor([[self.lr(t), 1.0 - betas[0] ** (t + 1), 1.0 - betas[1] ** (t + 1)] for t in tt],
                                   dtype=torch.float64).to(torch.float32).to(self.device).contiguous()
        self._step_dev = torch.zeros(1, device=self.device, dtype=torch.int32)
        self._graph = None
        self._static: Dict[str, torch.Tensor] = {}
        self._scalars = torch.zeros(4, device=self.device, dtype=torch.float32)       # [clip scale, grad norm, -, -]
        self._acc = torch.zeros(2, device=self.device, dtype=torch.float64)           # [loss, sum of squares]

    # ---------------------------------------------------------------- pieces
    def lr(self, step: Optional[int] = None) -> float:
        """CosineAnnealingLR(T_max=num_steps, eta_min) of the reference (:41-46), value used AT optimiser step `step`."""
        t = self.step_count if step is None else step
        return self.eta_min + (self.lr0 - self.eta_min) * (1.0 + math.cos(math.pi * t / self.num_steps)) / 2.0

    def _processor_backward(self, lvl: str, G: torch.Tensor, tokens: torch.Tensor) -> None:
        """Adjoint of  ehs' = ehs + sigmoid(alpha) * mean_k(W2 gelu(W1 a_k + b1) + b2)  (reference :88-97) given
        G = d loss / d ehs' [B,T,768] summed over the level's sites; writes the level's slice of self.grad."""
        p = self.procs[lvl]
        B, K, Da = tokens.shape
        W1, b1, W2, b2 = p.audio_proj[0].weight, p.audio_proj[0].bias, p.audio_proj[3].weight, p.audio_proj[3].bias
        a = tokens.reshape(B * K, Da)
        a = (a if a.dtype == torch.float32 else ops.cast(a.contiguous(), torch.float32)).contiguous()
        z = ops.linear(a, W1, b1)                                              # [BK, 64]
        h = ops.unary(z, ops.ACT_GELU)
        hbar = ops.token_mean(h.view(B, K, -1))                                # [B, 64]
        af = ops.linear(hbar, W2, b2)                                          # [B, 768]
        s = ops.colsum(G.contiguous())                                         # [B, 768] fp32: sum over the text positions
        sl = self.slots[lvl]
        g = self.grad
        daf = ops.gate_bwd(s, af, p.alpha, g[sl["alpha"]])
        ops.colsum(daf.view(1, B, -1), out=g[sl["audio_proj.3.bias"]].view(1, -1))
        dafT = ops.transpose(daf.view(1, B, -1)).view(-1, B)                   # [768, B]
        hbarT = ops.transpose(hbar.view(1, B, -1)).view(-1, B)                 # [64, B]
        ops.linear(dafT, hbarT, out=g[sl["audio_proj.3.weight"]].view(W2.shape))          # dW2 = daf^T hbar
        dhbar = ops.linear(daf, W2.t().contiguous())                           # [B, 64]
        dz = ops.gelu_bwd_bcast(z, dhbar, K)                                   # [BK, 64]
        ops.colsum(dz.view(1, B * K, -1), out=g[sl["audio_proj.0.bias"]].view(1, -1))
        dzT = ops.transpose(dz.view(1, B * K, -1)).view(-1, B * K)             # [64, BK]
        aT = ops.transpose(a.view(1, B * K, -1)).view(-1, B * K)               # [768, BK]
        ops.linear(dzT, aT, out=g[sl["audio_proj.0.weight"]].view(W1.shape))              # dW1 = dz^T a

    # ---------------------------------------------------------------- the step
    def forward_backward(self, audio_emb: torch.Tensor, latents: torch.Tensor, text_states: torch.Tensor,
                         noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """Fills self.grad (all-reduced over the group) and returns the device scalar loss (float64 [1]).
        audio_emb [B,512] (CLAP), latents / noise fp32 [B,4,H,W], text_states [B,77,768], timesteps [B] (integers)."""
        dev, dt = self.device, self.dtype
        B = latents.shape[0]
        self.grad.zero_()
        self._acc.zero_()
        # conditioning (frozen projector / decomposer) and the trainable injection
        clap = audio_emb.to(dev)
        clap = clap if clap.dtype == dt else ops.cast(clap.float().contiguous(), dt)
        routed = self.hier.encode(clap.contiguous(), with_tokens77=False)["routed"]
        ehs = text_states.to(dev)
        ehs = (ehs if ehs.dtype == dt else ops.cast(ehs.float().contiguous(), dt)).contiguous()
        kv: Dict[str, torch.Tensor] = {}
        ctx = {lvl: self.procs[lvl].context(ehs, routed[lvl]) for lvl in LEVELS}
        for name, lvl in self.level_of.items():
            kv[name] = ops.linear(ctx[lvl], self.unet.sites[name].wkv)
        # the reference's noising (train_stage3.py:203-204)
        t = timesteps.to(dev).float()
        a = (1.0 - t / 1000.0).view(-1, 1, 1, 1)
        noise = noise.to(dev).float().contiguous()
        noisy = (a * latents.to(dev).float() + (1.0 - a) * noise).contiguous()
        x = ops.nchw_to_nhwc(noisy, dt)
        # time-embedding rows: the time MLP and the 22 time_emb_proj layers are frozen, so the whole table of the 1000
        # training timesteps is evaluated once (82 MB fp32) and a step only gathers its B rows
        if self._time_table is None:
            self._time_table = torch.cat([self.unet.time_table([float(v) for v in range(i, min(i + 250, 1000))])
                                          for i in range(0, 1000, 250)], 0)
        temb_rows = self._time_table.index_select(0, timesteps.to(dev).long().clamp(0, 999))
        # forward with the tape, loss, adjoint
        tp = _Tape()
        pending = {lvl: sum(1 for v in self.level_of.values() if v == lvl) for lvl in LEVELS}
        G: Dict[str, Optional[torch.Tensor]] = {lvl: None for lvl in LEVELS}
        handles = []

        def on_kv_grad(site: str, dkv: torch.Tensor) -> None:
            lvl = self.level_of[site]
            wkv_t = self.net._wt(f"{site}.wkv", self.unet.sites[site].wkv)          # [768, 2C]
            G[lvl] = ops.linear(dkv, wkv_t, residual=G[lvl])                         # d ehs' of this site, summed per level
            pending[lvl] -= 1
            if pending[lvl] == 0:
                self._processor_backward(lvl, G[lvl], routed[lvl])
                if self.world > 1:
                    # one bucket per level, in flight while the rest of the backward pass runs
                    handles.append(torch.distributed.all_reduce(self.grad[self.level_range[lvl]], group=self.group, async_op=True))

        self.net.kv_grad_hook = on_kv_grad
        eps = self.net.forward(tp, x, temb_rows, kv)
        # loss scaled by 1 / world: the all-reduce SUM then yields the data-parallel mean (DDP semantics)
        g_eps = ops.mse_loss_grad(eps, noise, self.diffusion_weight / self.world, self._acc[0:1])
        tp.add_grad(eps, g_eps)
        tp.run()
        self.net.kv_grad_hook = None
        for h in handles:
            h.wait()
        return self._acc[0:1]

    def optimizer_step(self) -> None:
        """clip_grad_norm_(0.5) + AdamW + cosine schedule on the flat buffers (train_stage3.py:182-189)."""
        self.step_count += 1
        ops.sumsq(self.grad, self._acc[1:2])
        ops.clip_scale(self._acc[1:2], self.clip, self._scalars[0:1], self._scalars[1:2])
        ops.adamw_step_sched(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self._sched, self._step_dev, self.betas[0],
                             self.betas[1], self.adam_eps, self.wd, self._scalars[0:1])
        self._step_dev.add_(1)
        for p in self.procs.values():          # parameters changed under the processors' cast caches
            p._cache._c.clear()
            p._kv_cache.clear()

    def train_step(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """batch: audio_embedding [B,512], image_latents [B,4,H,W], text_embedding [B,77,768] (reference :134-136) plus
        the step's randomness noise [B,4,H,W] and timesteps [B] (drawn by the caller so that runs are reproducible).
        Returns device tensors {'diffusion': loss (this rank's share of the mean), 'grad_norm': pre-clip global norm}."""
        if self._graph is not None:
            return self._replay(batch)
        loss = self.forward_backward(batch["audio_embedding"], batch["image_latents"], batch["text_embedding"],
                                     batch["noise"], batch["timesteps"])
        self.optimizer_step()
        return {"diffusion": loss.clone(), "grad_norm": self._scalars[1:2].clone()}

    # ---------------------------------------------------------------- the step as one CUDA graph
    _KEYS = ("audio_embedding", "image_latents", "text_embedding", "noise", "timesteps")

    def capture(self, batch: Dict[str, torch.Tensor]) -> None:
        """Capture forward + reverse pass + all-reduce buckets + optimiser update for batches shaped like `batch` into one
        CUDA graph; `train_step` replays it from then on.  Eagerly the step is ~1100 launches at ~27 us of Python / ctypes
        each -- host-bound at 29 ms where the device needs less; a replay costs one launch.  The schedule and the step
        counter are device-resident, so the graph holds no per-step constant.  One un-timed forward + backward runs first
        (without an update: parameters and optimiser state are untouched) to build the lazily created constants."""
        dev = self.device
        self._static = {k: batch[k].to(dev).clone().contiguous() for k in self._KEYS}
        st = self._static
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.forward_backward(st["audio_embedding"], st["image_latents"], st["text_embedding"], st["noise"], st["timesteps"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.procs.values():          # the casts of the TRAINABLE parameters must be part of the graph, not cache hits
            p._cache._c.clear()
            p._kv_cache.clear()
        graph = torch.cuda.CUDAGraph()
        n0 = ops._lib.launch_count()
        with torch.cuda.graph(graph):
            loss = self.forward_backward(st["audio_embedding"], st["image_latents"], st["text_embedding"], st["noise"],
                                         st["timesteps"])
            self.optimizer_step()
            self._graph_out = {"diffusion": loss.clone(), "grad_norm": self._scalars[1:2].clone()}
        # the capture pass launched nothing: undo its host-side bookkeeping
        self.step_count -= 1
        self._graph = graph
        self.graph_launches = int(ops._lib.launch_count() - n0)      # libc2d kernels one replay launches

    def release_graph(self) -> None:
        """Drop the captured graph (back to eager steps).  Do this before `torch.distributed.destroy_process_group()`: a live
        graph that holds NCCL kernels keeps the communicator's teardown waiting."""
        self._graph = None
        self._graph_out = None
        self._static = {}
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def _replay(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        for k in self._KEYS:
            src, dst = batch[k], self._static[k]
            if src.data_ptr() != dst.data_ptr():
                if tuple(src.shape) != tuple(dst.shape):
                    raise ValueError(f"train_step: the captured graph expects {k} of shape {tuple(dst.shape)}, got {tuple(src.shape)}")
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self.step_count += 1
        return {k: v.clone() for k, v in self._graph_out.items()}

    def named_grads(self) -> Dict[str, Dict[str, torch.Tensor]]:
        return {lvl: {k: self.grad[sl].view(dict(self.procs[lvl].named_parameters())[k].shape) for k, sl in self.slots[lvl].items()}
                for lvl in LEVELS}

    def state_dict(self) -> Dict[str, object]:
        """The layout scripts/inference.py reads as ``unet_adapter_final.pth`` (+ optimiser state)."""
        out: Dict[str, object] = {"mode": "add", "step": self.step_count}
        for lvl in LEVELS:
            out[f"processor_{lvl}"] = {k: v.detach().clone().cpu() for k, v in self.procs[lvl].state_dict().items()}
        out["optimizer_state_dict"] = {"exp_avg": self.exp_avg.cpu(), "exp_avg_sq": self.exp_avg_sq.cpu(), "step": self.step_count}
        return out
