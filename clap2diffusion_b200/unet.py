"""SD-1.5 UNet (diffusers ``UNet2DConditionModel`` semantics) executed entirely by
libc2d kernels.  Host side only: weight packing, buffer plumbing and the launch sequence.

Design points (B200-first, not a port of diffusers):
  * activations stay channels-last [B, H*W, C] for the whole network: 1x1 convs and all projections are
    plain GEMMs, 3x3 convs are implicit GEMMs fed by 4-D TMA boxes, no NCHW<->NHWC permutes;
  * skip concatenation is folded into the consuming GroupNorm (two-source read), the resnet shortcut
    runs first and is folded into conv2's epilogue as the residual;
  * every step-invariant quantity is hoisted: the time-embedding MLP and all 22 ``time_emb_proj``
    outputs for all timesteps (``time_table``), the audio injection and the attn2 K/V projections
    (``prepare_conditioning``);
  * attn1 uses one fused QKV GEMM; GEGLU is fused into the FF GEMM epilogue (bf16 mode);
  * bf16 mode: every tensor a GroupNorm will read leaves its producer's epilogue together with per-channel
    fixed-point statistics (``ops.linear/conv3x3(stats=...)``), so GroupNorm is ONE pass (``group_norm_apply``), and
    the shortcut GEMM reads [h | skip] through two tensor maps instead of a materialised concat;
  * the launch sequence has no host synchronisation, so a whole step is capturable in a CUDA graph.

The module exposes ``attn_processors`` / ``set_attn_processor`` with diffusers' processor names so the
reference's ``AudioProcessorManager`` logic applies unchanged.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import ops
from .models.audio_attention_processor import AudioAttnProcessor

BLOCK_OUT = (320, 640, 1280, 1280)
CROSS_DIM = 768
HEADS = 8
GROUPS = 32


class _Lin:
    """nn.Linear-shaped weight holder (``.weight`` [out,in], ``.bias``) living in the engine dtype."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None):
        self.weight = weight
        self.bias = bias
        self.out_features, self.in_features = weight.shape


class AttentionSite:
    """What a processor sees as ``attn`` (the subset of diffusers' ``Attention`` the reference's
    AudioAttnProcessor reads: to_q/to_k/to_v/to_out, heads, scale and the SD-1.5 flag values)."""

    spatial_norm = None
    norm_cross = None
    group_norm = None
    residual_connection = False
    rescale_output_factor = 1.0

    def __init__(self, name: str, to_q: _Lin, to_k: _Lin, to_v: _Lin, to_out: _Lin, heads: int = HEADS):
        self.name = name
        self.to_q, self.to_k, self.to_v = to_q, to_k, to_v
        self.to_out = [to_out]
        self.heads = heads
        self.scale = (to_q.out_features // heads) ** -0.5
        self.processor = None
        self.wqkv: Optional[torch.Tensor] = None     # fused [3C, C] for self-attention


class SelfAttnProcessor:
    """Default attn1 processor: fused QKV GEMM -> flash attention -> out projection (+ residual).
    Accepts and ignores cross_attention_kwargs such as ``audio`` (decision D5)."""

    def __call__(self, attn: AttentionSite, hidden_states, encoder_hidden_states=None, attention_mask=None,
                 temb=None, scale: float = 1.0, residual=None, ln_stats=None, row_stats=None, **kwargs):
        """ln_stats: hidden_states is the UN-normalised stream and norm1 is folded into the QKV GEMM (attn.ln_qkv);
        row_stats: accumulator for the row statistics of the output (the next LayerNorm's input)."""
        C = attn.to_q.out_features
        if ln_stats is not None:
            qkv = ops.linear(hidden_states, None, ln=attn.ln_qkv, ln_stats=ln_stats)
        else:
            qkv = ops.linear(hidden_states, attn.wqkv)
        o = ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], attn.heads, scale=attn.scale * scale)
        return ops.linear(o, attn.to_out[0].weight, attn.to_out[0].bias, residual=residual, row_stats=row_stats)


class CrossAttnProcessor:
    """Default attn2 processor (plain text cross-attention) with the prepare/attend split."""

    def prepare(self, attn: AttentionSite, encoder_hidden_states, audio=None):
        return ops.linear(encoder_hidden_states, attn.wkv)

    supports_ln_fold = True
    supports_packed_kv = True

    def attend(self, attn: AttentionSite, hidden_states, kv, residual=None, scale: float = 1.0, ln_stats=None,
               row_stats=None):
        """kv: cached [B,T,2C] tensor, or an ops.XattnKV (packed cache): then to_q + attention are ONE kernel."""
        C = attn.to_q.out_features
        if isinstance(kv, ops.XattnKV) and not ops.xattn_supported(hidden_states, attn.heads, kv.T, kv.T2):
            kv = kv.kv                       # token counts outside the fused kernel (tiny latents): three-kernel path
        if isinstance(kv, ops.XattnKV):
            if ln_stats is not None:
                o = ops.xattn(hidden_states, kv, ln=attn.ln_q, ln_stats=ln_stats, scale=attn.scale * scale)
            else:
                o = ops.xattn(hidden_states, kv, wq=attn.to_q.weight, scale=attn.scale * scale)
            return ops.linear(o, attn.to_out[0].weight, attn.to_out[0].bias, residual=residual, row_stats=row_stats)
        if ln_stats is not None:
            q = ops.linear(hidden_states, None, ln=attn.ln_q, ln_stats=ln_stats)
        else:
            q = ops.linear(hidden_states, attn.to_q.weight)
        o = ops.attention(q, kv[..., :C], kv[..., C:], attn.heads, scale=attn.scale * scale)
        return ops.linear(o, attn.to_out[0].weight, attn.to_out[0].bias, residual=residual, row_stats=row_stats)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0, **kwargs):
        return self.attend(attn, hidden_states, self.prepare(attn, encoder_hidden_states), scale=scale)


def unet_topology():
    """(resnet (cin, cout) lists, attention flags, samplers) of the SD-1.5 UNet."""
    down, cin = [], BLOCK_OUT[0]
    for i, cout in enumerate(BLOCK_OUT):
        down.append(dict(resnets=[(cin, cout), (cout, cout)], attn=i < 3, sample=i < 3, c=cout))
        cin = cout
    skips = [BLOCK_OUT[0]]
    for blk in down:
        skips += [blk["c"]] * (2 + (1 if blk["sample"] else 0))
    up, prev = [], BLOCK_OUT[-1]
    for i, cout in enumerate(reversed(BLOCK_OUT)):
        layers = []
        for j in range(3):
            layers.append(((prev if j == 0 else cout), skips.pop(), cout))
        up.append(dict(resnets=layers, attn=i > 0, sample=i < 3, c=cout))
        prev = cout
    return down, up


class _StatsArena:
    """Zero-initialised int64 arena for the per-tensor channel statistics of one forward pass ([B, C, 2] each)."""

    def __init__(self, device, capacity: int):
        self.buf = torch.zeros(capacity, device=device, dtype=torch.int64)
        self.off = 0

    def take_rows(self, M: int) -> torch.Tensor:
        return self.take(M, 1)

    def take(self, B: int, C: int) -> torch.Tensor:
        n = B * C * 2
        if self.off + n > self.buf.numel():
            raise RuntimeError("channel-statistics arena exhausted")
        v = self.buf[self.off:self.off + n]
        self.off += n
        return v


class SD15UNet:
    """``SD15UNet(state_dict, device, dtype)``; state-dict keys/shapes are diffusers' (fp32, any device)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", dtype=torch.bfloat16,
                 impl: int = ops.IMPL_AUTO, fused: Optional[bool] = None):
        """fused=False keeps every layer a separate kernel with plain weights (the layout the training step's backward
        walks, clap2diffusion_b200/train.py); default: all inference fusions on in bf16, off in the fp32 parity mode."""
        self.device = torch.device(device)
        self.dtype = dtype
        self.impl = impl
        fused = (dtype == torch.bfloat16) if fused is None else (bool(fused) and dtype == torch.bfloat16)
        self.fuse_geglu = fused
        self.fused_gn = fused       # producer-side GroupNorm statistics (tcgen05 epilogues)
        self.fold_ln = fused        # norm1/2/3 folded into the QKV / to_q / GEGLU GEMMs
        self.conv_in_tc = fused     # conv_in on the tensor cores over a channel-padded (4 -> 8) input
        self.fused_xattn = fused    # attn2 sites: to_q + attention in one kernel over a packed K/V cache
        self._gn_channels = 0
        self._sd = state_dict
        self.w: Dict[str, torch.Tensor] = {}
        self.sites: Dict[str, AttentionSite] = {}
        self._temb_offsets: Dict[str, int] = {}
        self._temb_total = 0
        self.down, self.up = unet_topology()
        self._pack()
        self._sd = None
        self._self_proc = SelfAttnProcessor()
        for s in self.sites.values():
            s.processor = self._self_proc if ".attn1" in s.name else CrossAttnProcessor()

    # ------------------------------------------------------------------ weights
    def _dev32(self, name):
        return self._sd[name].detach().to(self.device, torch.float32).contiguous()

    def _mat(self, name):      # matmul weight in the engine dtype ([out,in]; 1x1 conv kernels squeezed)
        w = self._dev32(name)
        if w.dim() == 4:
            w = w.reshape(w.shape[0], w.shape[1]).contiguous()
        return w if self.dtype == torch.float32 else ops.cast(w, self.dtype)

    def _conv(self, prefix):
        self.w[f"{prefix}.weight"] = ops.pack_conv3x3(self._dev32(f"{prefix}.weight"), self.dtype)
        self.w[f"{prefix}.bias"] = self._dev32(f"{prefix}.bias")

    def _linear(self, prefix, bias=True):
        self.w[f"{prefix}.weight"] = self._mat(f"{prefix}.weight")
        if bias:
            self.w[f"{prefix}.bias"] = self._dev32(f"{prefix}.bias")

    def _norm(self, prefix):
        self.w[f"{prefix}.weight"] = self._dev32(f"{prefix}.weight")
        self.w[f"{prefix}.bias"] = self._dev32(f"{prefix}.bias")
        if "transformer_blocks" not in prefix:             # GroupNorm site: bounds the statistics arena
            self._gn_channels += self.w[f"{prefix}.weight"].numel()

    def _pack_resnet(self, name, cin, cout):
        self._norm(f"{name}.norm1"); self._conv(f"{name}.conv1")
        self._norm(f"{name}.norm2"); self._conv(f"{name}.conv2")
        # time_emb_proj stays fp32: it is evaluated once per timestep table, not per step.  The timestep is
        # shared by the whole batch, so conv1's bias is folded into the table row and the row is passed
        # to conv1 as its bias vector.
        self.w[f"{name}.time_emb_proj.weight"] = self._dev32(f"{name}.time_emb_proj.weight")
        self.w[f"{name}.time_emb_proj.bias"] = ops.add(self._dev32(f"{name}.time_emb_proj.bias"),
                                                       self.w[f"{name}.conv1.bias"])
        self._temb_offsets[name] = self._temb_total
        self._temb_total += cout
        if cin != cout:
            self._linear(f"{name}.conv_shortcut")

    def _pack_transformer(self, name, c):
        tb = f"{name}.transformer_blocks.0"
        self._norm(f"{name}.norm"); self._linear(f"{name}.proj_in"); self._linear(f"{name}.proj_out")
        for n in ("norm1", "norm2", "norm3"):
            self._norm(f"{tb}.{n}")
        for a in ("attn1", "attn2"):
            lins = {}
            for p in ("to_q", "to_k", "to_v"):
                lins[p] = _Lin(self._mat(f"{tb}.{a}.{p}.weight"))
            lins["to_out"] = _Lin(self._mat(f"{tb}.{a}.to_out.0.weight"), self._dev32(f"{tb}.{a}.to_out.0.bias"))
            site = AttentionSite(f"{tb}.{a}", lins["to_q"], lins["to_k"], lins["to_v"], lins["to_out"])
            if a == "attn1":
                site.wqkv = torch.cat([lins["to_q"].weight, lins["to_k"].weight, lins["to_v"].weight], 0).contiguous()
                if self.fold_ln:
                    w32 = torch.cat([self._dev32(f"{tb}.{a}.{p}.weight") for p in ("to_q", "to_k", "to_v")], 0).contiguous()
                    site.ln_qkv = ops.pack_lnfold(w32, self.w[f"{tb}.norm1.weight"], self.w[f"{tb}.norm1.bias"], None, self.dtype)
            else:
                site.wkv = torch.cat([lins["to_k"].weight, lins["to_v"].weight], 0).contiguous()
                if self.fold_ln:
                    site.ln_q = ops.pack_lnfold(self._dev32(f"{tb}.{a}.to_q.weight"), self.w[f"{tb}.norm2.weight"],
                                                self.w[f"{tb}.norm2.bias"], None, self.dtype)
            self.sites[f"{tb}.{a}"] = site
        w, b = self._dev32(f"{tb}.ff.net.0.proj.weight"), self._dev32(f"{tb}.ff.net.0.proj.bias")
        if self.fuse_geglu:
            self.w[f"{tb}.ff.geglu.weight"], self.w[f"{tb}.ff.geglu.bias"] = ops.pack_geglu(w, b, self.dtype)
            if self.fold_ln:
                f = ops.pack_lnfold(w, self.w[f"{tb}.norm3.weight"], self.w[f"{tb}.norm3.bias"], b, self.dtype,
                                    out_dtype=torch.float32)
                wp, bp = ops.pack_geglu(f.w, f.bias, self.dtype)
                _, csp = ops.pack_geglu(f.w, f.colsum, self.dtype)
                self.w[f"{tb}.ff.geglu_ln"] = ops.LNFold(wp, csp, bp, 1e-5)
        else:
            self.w[f"{tb}.ff.net.0.proj.weight"] = w if self.dtype == torch.float32 else ops.cast(w, self.dtype)
            self.w[f"{tb}.ff.net.0.proj.bias"] = b
        self._linear(f"{tb}.ff.net.2")

    def _pack(self):
        # time MLP in fp32 (once per timestep table)
        for n in ("time_embedding.linear_1", "time_embedding.linear_2"):
            self.w[f"{n}.weight"] = self._dev32(f"{n}.weight")
            self.w[f"{n}.bias"] = self._dev32(f"{n}.bias")
        self._conv("conv_in")
        if self.conv_in_tc:
            w4 = self._dev32("conv_in.weight")
            w8 = torch.zeros(w4.shape[0], 8, 3, 3, device=self.device, dtype=torch.float32)
            w8[:, :4].copy_(w4)
            self.w["conv_in8.weight"] = ops.pack_conv3x3(w8, self.dtype)
        for i, blk in enumerate(self.down):
            for j, (cin, cout) in enumerate(blk["resnets"]):
                self._pack_resnet(f"down_blocks.{i}.resnets.{j}", cin, cout)
                if blk["attn"]:
                    self._pack_transformer(f"down_blocks.{i}.attentions.{j}", cout)
            if blk["sample"]:
                self._conv(f"down_blocks.{i}.downsamplers.0.conv")
        c = BLOCK_OUT[-1]
        self._pack_resnet("mid_block.resnets.0", c, c)
        self._pack_transformer("mid_block.attentions.0", c)
        self._pack_resnet("mid_block.resnets.1", c, c)
        for i, blk in enumerate(self.up):
            for j, (ch, cs, cout) in enumerate(blk["resnets"]):
                self._pack_resnet(f"up_blocks.{i}.resnets.{j}", ch + cs, cout)
                if blk["attn"]:
                    self._pack_transformer(f"up_blocks.{i}.attentions.{j}", cout)
            if blk["sample"]:
                self._conv(f"up_blocks.{i}.upsamplers.0.conv")
        self._norm("conv_norm_out")
        self._conv("conv_out")

    # ------------------------------------------------------------------ diffusers-style processor registry
    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{n}.processor": s.processor for n, s in self.sites.items()}

    def set_attn_processor(self, processor) -> None:
        if isinstance(processor, dict):
            missing = set(self.attn_processors) - set(processor)
            if missing:
                raise ValueError(f"set_attn_processor: {len(missing)} processor names missing, e.g. {sorted(missing)[0]}")
            for n, s in self.sites.items():
                s.processor = processor[f"{n}.processor"]
        else:
            for s in self.sites.values():
                s.processor = processor
        for p in {id(s.processor): s.processor for s in self.sites.values()}.values():
            if isinstance(p, torch.nn.Module):
                p.to(self.device)

    def get_submodule(self, target: str) -> AttentionSite:
        """nn.Module-style lookup of an attention site, e.g.
        ``down_blocks.0.attentions.0.transformer_blocks.0.attn2``."""
        if target not in self.sites:
            raise AttributeError(f"SD15UNet has no attention site {target!r}")
        return self.sites[target]

    def cross_sites(self) -> List[AttentionSite]:
        return [s for n, s in self.sites.items() if n.endswith("attn2")]

    # ------------------------------------------------------------------ step-invariant work
    def time_table(self, timesteps: Sequence[float]) -> torch.Tensor:
        """fp32 [S, sum(Cout)] : time_emb_proj(silu(time_mlp(t))) + conv1.bias of every resnet, per timestep."""
        t = torch.tensor([float(v) for v in timesteps], dtype=torch.float32, device=self.device)
        e = ops.timestep_embedding(t, BLOCK_OUT[0])
        e = ops.linear(e, self.w["time_embedding.linear_1.weight"], self.w["time_embedding.linear_1.bias"], act=ops.ACT_SILU)
        e = ops.linear(e, self.w["time_embedding.linear_2.weight"], self.w["time_embedding.linear_2.bias"], act=ops.ACT_SILU)
        table = torch.empty(len(timesteps), self._temb_total, device=self.device, dtype=torch.float32)
        for name, off in self._temb_offsets.items():
            w = self.w[f"{name}.time_emb_proj.weight"]
            ops.linear(e, w, self.w[f"{name}.time_emb_proj.bias"], out=table[:, off:off + w.shape[0]])
        return table

    def prepare_conditioning(self, encoder_hidden_states: torch.Tensor,
                             cross_attention_kwargs: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        """Per attn2 site: cached K/V [B, T', 2C] (audio injection + to_k/to_v), computed once per image."""
        kw = cross_attention_kwargs or {}
        ehs = encoder_hidden_states
        if ehs.dtype != self.dtype:
            ehs = ops.cast(ehs.contiguous(), self.dtype)
        ehs = ehs.contiguous()
        out, ctx_cache = {}, {}
        for n, s in self.sites.items():
            if not n.endswith("attn2"):
                continue
            p = s.processor
            if isinstance(p, AudioAttnProcessor) and p.mode != "decoupled":
                key = id(p)                     # one shared processor per level -> one context per level
                if key not in ctx_cache:
                    audio = kw.get("audio")
                    ctx_cache[key] = p.context(ehs, audio.get(p.level) if isinstance(audio, dict) else None)
                out[n] = ops.linear(ctx_cache[key], s.wkv)
            elif hasattr(p, "prepare"):
                out[n] = p.prepare(s, ehs, kw.get("audio"))
            else:
                out[n] = None                   # opaque processor: called with encoder_hidden_states each step
            kvt = out[n]
            if (self.fused_xattn and torch.is_tensor(kvt) and getattr(p, "supports_packed_kv", False)
                    and ops.xattn_packable(s.to_q.out_features, s.heads, kvt.shape[1], kvt.dtype)):
                out[n] = ops.xattn_pack_kv(kvt, s.heads)     # per-(batch, head) smem images for the fused kernel
        return out

    # ------------------------------------------------------------------ forward
    def _resnet(self, name, x, temb_row, skip=None, xs=None, skip_s=None, ar=None):
        """ResnetBlock2D.  bf16 mode: (x, xs) / (skip, skip_s) are tensors with their channel statistics and the
        block returns (out, out_stats); fp32 mode: statistics are None and GroupNorm is the two-pass kernel."""
        w = self.w
        B = x.shape[0]
        cin = x.shape[-1] + (0 if skip is None else skip.shape[-1])
        cout = w[f"{name}.conv1.weight"].shape[0]
        has_sc = f"{name}.conv_shortcut.weight" in w
        off = self._temb_offsets[name]
        if skip is not None and not has_sc:
            raise RuntimeError("concat resnet without shortcut is not part of SD-1.5")
        if ar is not None:
            h = ops.group_norm_apply(x, xs, w[f"{name}.norm1.weight"], w[f"{name}.norm1.bias"], GROUPS, 1e-5, True,
                                     x2=skip, stats2=skip_s)
            s1 = ar.take(B, cout)
            h = ops.conv3x3(h, w[f"{name}.conv1.weight"], temb_row[off:off + cout], impl=self.impl, stats=s1)
            h = ops.group_norm_apply(h, s1, w[f"{name}.norm2.weight"], w[f"{name}.norm2.bias"], GROUPS, 1e-5, True)
            sc = x
            if has_sc:
                sc = ops.linear(x, w[f"{name}.conv_shortcut.weight"], w[f"{name}.conv_shortcut.bias"], x2=skip, impl=self.impl)
            so = ar.take(B, cout)
            out = ops.conv3x3(h, w[f"{name}.conv2.weight"], w[f"{name}.conv2.bias"], residual=sc, impl=self.impl, stats=so)
            return out, so
        raw = None
        if skip is not None and has_sc:
            raw = torch.empty(*x.shape[:-1], cin, device=x.device, dtype=x.dtype)
        h = ops.group_norm(x, w[f"{name}.norm1.weight"], w[f"{name}.norm1.bias"], GROUPS, 1e-5, True, x2=skip, raw_cat=raw)
        h = ops.conv3x3(h, w[f"{name}.conv1.weight"], temb_row[off:off + cout], impl=self.impl)   # bias+temb row
        h = ops.group_norm(h, w[f"{name}.norm2.weight"], w[f"{name}.norm2.bias"], GROUPS, 1e-5, True)
        if has_sc:
            src = raw if raw is not None else x
            sc = ops.linear(src, w[f"{name}.conv_shortcut.weight"], w[f"{name}.conv_shortcut.bias"], impl=self.impl)
        else:
            sc = x
        return ops.conv3x3(h, w[f"{name}.conv2.weight"], w[f"{name}.conv2.bias"], residual=sc, impl=self.impl), None

    def _transformer(self, name, x, kv, ehs, kw, xs=None, ar=None):
        w = self.w
        B, H, W, C = x.shape
        tb = f"{name}.transformer_blocks.0"
        res = x.view(B, H * W, C)
        if ar is not None:
            h = ops.group_norm_apply(res, xs, w[f"{name}.norm.weight"], w[f"{name}.norm.bias"], GROUPS, 1e-6, False)
        else:
            h = ops.group_norm(res, w[f"{name}.norm.weight"], w[f"{name}.norm.bias"], GROUPS, 1e-6, False)
        s1, s2 = self.sites[f"{tb}.attn1"], self.sites[f"{tb}.attn2"]
        cached = kv.get(f"{tb}.attn2") if kv is not None else None
        if (ar is not None and self.fold_ln and isinstance(s1.processor, SelfAttnProcessor) and cached is not None
                and getattr(s2.processor, "supports_ln_fold", False)):
            # LayerNorm-free block: every GEMM that writes the residual stream also accumulates its row statistics,
            # every GEMM that reads LayerNorm(stream) takes the raw stream + statistics (ops.linear(ln=...))
            M = B * H * W
            rs0, rs1, rs2 = ar.take_rows(M), ar.take_rows(M), ar.take_rows(M)
            h = ops.linear(h, w[f"{name}.proj_in.weight"], w[f"{name}.proj_in.bias"], impl=self.impl, row_stats=rs0)
            h = s1.processor(s1, h, residual=h, ln_stats=rs0, row_stats=rs1, **kw)
            h = s2.processor.attend(s2, h, cached, residual=h, ln_stats=rs1, row_stats=rs2)
            g = ops.geglu_linear(h, None, None, ln=w[f"{tb}.ff.geglu_ln"], ln_stats=rs2)
            h = ops.linear(g, w[f"{tb}.ff.net.2.weight"], w[f"{tb}.ff.net.2.bias"], residual=h, impl=self.impl)
            so = ar.take(B, C)
            out = ops.linear(h, w[f"{name}.proj_out.weight"], w[f"{name}.proj_out.bias"], residual=res, impl=self.impl,
                             stats=so, stats_rows=H * W)
            return out.view(B, H, W, C), so
        h = ops.linear(h, w[f"{name}.proj_in.weight"], w[f"{name}.proj_in.bias"], impl=self.impl)
        # attn1
        n1 = ops.layer_norm(h, w[f"{tb}.norm1.weight"], w[f"{tb}.norm1.bias"])
        if isinstance(s1.processor, SelfAttnProcessor):
            h = s1.processor(s1, n1, residual=h, **kw)
        else:
            h = ops.add(h, s1.processor(s1, n1, **kw))
        # attn2
        n2 = ops.layer_norm(h, w[f"{tb}.norm2.weight"], w[f"{tb}.norm2.bias"])
        if cached is not None and hasattr(s2.processor, "attend"):
            h = s2.processor.attend(s2, n2, cached, residual=h)
        else:
            h = ops.add(h, s2.processor(s2, n2, encoder_hidden_states=ehs, **kw))
        # feed-forward
        n3 = ops.layer_norm(h, w[f"{tb}.norm3.weight"], w[f"{tb}.norm3.bias"])
        if self.fuse_geglu:
            g = ops.geglu_linear(n3, w[f"{tb}.ff.geglu.weight"], w[f"{tb}.ff.geglu.bias"])
        else:
            g = ops.geglu(ops.linear(n3, w[f"{tb}.ff.net.0.proj.weight"], w[f"{tb}.ff.net.0.proj.bias"], impl=self.impl))
        h = ops.linear(g, w[f"{tb}.ff.net.2.weight"], w[f"{tb}.ff.net.2.bias"], residual=h, impl=self.impl)
        so = ar.take(B, C) if ar is not None else None
        out = ops.linear(h, w[f"{name}.proj_out.weight"], w[f"{name}.proj_out.bias"], residual=res, impl=self.impl,
                         stats=so, stats_rows=H * W)
        return out.view(B, H, W, C), so

    def forward_nhwc(self, x: torch.Tensor, temb_row: torch.Tensor, kv: Optional[Dict[str, torch.Tensor]] = None,
                     encoder_hidden_states: Optional[torch.Tensor] = None,
                     cross_attention_kwargs: Optional[dict] = None, taps: Optional[dict] = None) -> torch.Tensor:
        """x [B,H,W,4] (or [B,H,W,8] with four zero padding channels, see conv_in_tc) in the engine dtype,
        temb_row fp32 [sum(Cout)] (one row of time_table) -> eps [B,H,W,4]."""
        w, kw = self.w, (cross_attention_kwargs or {})
        ehs = encoder_hidden_states
        B = x.shape[0]
        ar = None
        if self.fused_gn:
            # channel statistics of every GroupNorm input + row statistics of the three LayerNorm inputs of the
            # 5 / 5 / 5 / 1 transformer blocks at the four resolutions
            H0, W0 = x.shape[1], x.shape[2]
            rows = sum(n * (max(H0 >> l, 1) * max(W0 >> l, 1)) for l, n in enumerate((5, 5, 5, 1)))
            ar = _StatsArena(x.device, 2 * B * self._gn_channels + 2 * 3 * B * rows)
        if x.shape[-1] == 8:          # channel-padded input (sampler, bf16): tensor-core conv_in with fused statistics
            hs = ar.take(B, BLOCK_OUT[0]) if ar is not None else None
            h = ops.conv3x3(x, w["conv_in8.weight"], w["conv_in.bias"], impl=self.impl, stats=hs)
        else:
            h = ops.conv3x3(x, w["conv_in.weight"], w["conv_in.bias"], impl=self.impl)
            hs = ops.channel_stats(h, ar.take(B, h.shape[-1])) if ar is not None else None
        if taps is not None:
            taps["conv_in"] = h
        skips = [(h, hs)]
        for i, blk in enumerate(self.down):
            for j in range(2):
                h, hs = self._resnet(f"down_blocks.{i}.resnets.{j}", h, temb_row, xs=hs, ar=ar)
                if blk["attn"]:
                    h, hs = self._transformer(f"down_blocks.{i}.attentions.{j}", h, kv, ehs, kw, xs=hs, ar=ar)
                skips.append((h, hs))
            if blk["sample"]:
                n = f"down_blocks.{i}.downsamplers.0.conv"
                hs = ar.take(B, w[f"{n}.weight"].shape[0]) if ar is not None else None
                h = ops.conv3x3(h, w[f"{n}.weight"], w[f"{n}.bias"], stride=2, impl=self.impl, stats=hs)
                skips.append((h, hs))
        h, hs = self._resnet("mid_block.resnets.0", h, temb_row, xs=hs, ar=ar)
        h, hs = self._transformer("mid_block.attentions.0", h, kv, ehs, kw, xs=hs, ar=ar)
        h, hs = self._resnet("mid_block.resnets.1", h, temb_row, xs=hs, ar=ar)
        if taps is not None:
            taps["mid"] = h
        for i, blk in enumerate(self.up):
            for j in range(3):
                sk, sks = skips.pop()
                h, hs = self._resnet(f"up_blocks.{i}.resnets.{j}", h, temb_row, skip=sk, xs=hs, skip_s=sks, ar=ar)
                if blk["attn"]:
                    h, hs = self._transformer(f"up_blocks.{i}.attentions.{j}", h, kv, ehs, kw, xs=hs, ar=ar)
            if blk["sample"]:
                n = f"up_blocks.{i}.upsamplers.0.conv"
                if self.dtype == torch.bfloat16 or ar is not None:
                    hs = ar.take(B, w[f"{n}.weight"].shape[0]) if ar is not None else None
                    h = ops.conv3x3(ops.upsample2x(h), w[f"{n}.weight"], w[f"{n}.bias"], impl=self.impl, stats=hs)
                else:
                    h = ops.conv3x3(h, w[f"{n}.weight"], w[f"{n}.bias"], upsample=True, impl=self.impl)
        B, H, W, C = h.shape
        if ar is not None:
            h = ops.group_norm_apply(h.view(B, H * W, C), hs, w["conv_norm_out.weight"], w["conv_norm_out.bias"], GROUPS, 1e-5, True)
        else:
            h = ops.group_norm(h.view(B, H * W, C), w["conv_norm_out.weight"], w["conv_norm_out.bias"], GROUPS, 1e-5, True)
        return ops.conv3x3(h.view(B, H, W, C), w["conv_out.weight"], w["conv_out.bias"], impl=self.impl)

    def __call__(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                 cross_attention_kwargs: Optional[dict] = None, taps: Optional[dict] = None) -> torch.Tensor:
        """diffusers-style convenience call: sample fp32 NCHW [B,4,H,W], scalar timestep,
        encoder_hidden_states [B,77,768] -> eps fp32 NCHW.  (The sampler uses the hoisted pieces directly.)"""
        t = float(timestep) if not torch.is_tensor(timestep) else float(timestep.reshape(-1)[0])
        table = self.time_table([t])
        kv = self.prepare_conditioning(encoder_hidden_states, cross_attention_kwargs)
        x = ops.nchw_to_nhwc(sample.contiguous().float(), self.dtype)
        ehs = encoder_hidden_states
        if ehs.dtype != self.dtype:
            ehs = ops.cast(ehs.contiguous(), self.dtype)
        eps = self.forward_nhwc(x, table[0], kv, ehs, cross_attention_kwargs, taps=taps)
        return ops.nhwc_to_nchw(eps)


# ------------------------------------------------------------------------------------------------------
def param_shapes() -> Dict[str, tuple]:
    """diffusers state-dict names -> shapes of the SD-1.5 UNet (859,520,964 parameters)."""
    out: Dict[str, tuple] = {}

    def lin(n, cin, cout, bias=True):
        out[f"{n}.weight"] = (cout, cin)
        if bias:
            out[f"{n}.bias"] = (cout,)

    def conv(n, cin, cout, k):
        out[f"{n}.weight"] = (cout, cin, k, k)
        out[f"{n}.bias"] = (cout,)

    def norm(n, c):
        out[f"{n}.weight"] = (c,)
        out[f"{n}.bias"] = (c,)

    def resnet(n, cin, cout):
        norm(f"{n}.norm1", cin); conv(f"{n}.conv1", cin, cout, 3); lin(f"{n}.time_emb_proj", 1280, cout)
        norm(f"{n}.norm2", cout); conv(f"{n}.conv2", cout, cout, 3)
        if cin != cout:
            conv(f"{n}.conv_shortcut", cin, cout, 1)

    def transformer(n, c):
        tb = f"{n}.transformer_blocks.0"
        norm(f"{n}.norm", c); conv(f"{n}.proj_in", c, c, 1)
        for a, kd in (("attn1", c), ("attn2", CROSS_DIM)):
            norm(f"{tb}.norm{1 if a == 'attn1' else 2}", c)
            lin(f"{tb}.{a}.to_q", c, c, False); lin(f"{tb}.{a}.to_k", kd, c, False)
            lin(f"{tb}.{a}.to_v", kd, c, False); lin(f"{tb}.{a}.to_out.0", c, c)
        norm(f"{tb}.norm3", c); lin(f"{tb}.ff.net.0.proj", c, 8 * c); lin(f"{tb}.ff.net.2", 4 * c, c)
        conv(f"{n}.proj_out", c, c, 1)

    lin("time_embedding.linear_1", BLOCK_OUT[0], 1280); lin("time_embedding.linear_2", 1280, 1280)
    conv("conv_in", 4, BLOCK_OUT[0], 3)
    down, up = unet_topology()
    for i, blk in enumerate(down):
        for j, (cin, cout) in enumerate(blk["resnets"]):
            resnet(f"down_blocks.{i}.resnets.{j}", cin, cout)
            if blk["attn"]:
                transformer(f"down_blocks.{i}.attentions.{j}", cout)
        if blk["sample"]:
            conv(f"down_blocks.{i}.downsamplers.0.conv", blk["c"], blk["c"], 3)
    c = BLOCK_OUT[-1]
    resnet("mid_block.resnets.0", c, c); transformer("mid_block.attentions.0", c); resnet("mid_block.resnets.1", c, c)
    for i, blk in enumerate(up):
        for j, (ch, cs, cout) in enumerate(blk["resnets"]):
            resnet(f"up_blocks.{i}.resnets.{j}", ch + cs, cout)
            if blk["attn"]:
                transformer(f"up_blocks.{i}.attentions.{j}", cout)
        if blk["sample"]:
            conv(f"up_blocks.{i}.upsamplers.0.conv", blk["c"], blk["c"], 3)
    norm("conv_norm_out", BLOCK_OUT[0]); conv("conv_out", BLOCK_OUT[0], 4, 3)
    return out
