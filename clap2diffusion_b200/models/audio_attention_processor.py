"""Drop-in for the reference's ``models/audio_attention_processor.py`` (AudioAttnProcessor :13-145,
AudioProcessorManager :148-267) with the compute behind libc2d's C ABI.

Same class names, constructor arguments, ``state_dict()`` keys (``alpha``, ``audio_proj.0.*``,
``audio_proj.3.*``) and the diffusers processor call convention
``proc(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
**cross_attention_kwargs)`` with audio arriving as ``cross_attention_kwargs['audio'][level]``.

Design extension (no reference counterpart, opt-in): ``mode="decoupled"`` -- the text keys stay untouched and the
projected audio tokens get their own softmax through the site's to_k / to_v; the audio branch is scaled by
sigmoid(alpha) and added before to_out (BASELINE.json north_star "fused decoupled text+audio cross-attention").
It runs on the fused bf16 kernel only (``ops.xattn``); the reference treats the string as an unknown mode.

What differs from the reference (by design, results identical):
  * the audio injection + K/V projections are step-invariant, so they are exposed separately
    (``prepare`` -> cached K/V, ``attend`` -> per-step work); ``__call__`` = prepare + attend;
  * attention probabilities are never materialised (flash kernel), no head permutes;
  * the stand-alone ``__call__`` (the path a real diffusers UNet takes) keeps the projected / packed K/V of a site in
    a small identity-keyed cache: a denoising loop passes the SAME ``encoder_hidden_states`` / audio tensors every
    step, so K/V are projected once per image there too (any in-place change bumps ``_version`` and misses);
  * ``attention_mask`` (reference :129 hands it to ``attn.get_attention_scores``, i.e. an additive bias): key-padding
    masks -- boolean, or additive 0 / -inf of shape [B,T], [B,1,T] or [B*heads,1,T] -- are honoured; a bias that
    differs per head or per query row raises ``C2DError`` (no kernel for it, never silently ignored);
  * CUDA only -- a CPU tensor raises ``C2DError`` (no fallback).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from .. import ops
from .._lib import C2DError

_MODES = {"add": ops.AUDIO_ADD, "concat": ops.AUDIO_CONCAT}


class _CastCache:
    """Device copies of parameters in the compute dtype, refreshed when the parameter changes."""

    def __init__(self):
        self._c: Dict[Any, Any] = {}

    def get(self, t: torch.Tensor, dtype: torch.dtype, key=None) -> torch.Tensor:
        t = t.detach()
        if t.dtype == dtype and t.is_contiguous():
            return t
        k = (key if key is not None else id(t), dtype)
        hit = self._c.get(k)
        if hit is not None and hit[0] == (t._version, t.data_ptr()):
            return hit[1]
        out = ops.cast(t.contiguous(), dtype) if t.dtype in (torch.float32, torch.bfloat16) else t.to(dtype)
        self._c[k] = ((t._version, t.data_ptr()), out)
        return out


def _weight(mod) -> torch.Tensor:
    return mod.weight


def _to_out_linear(attn):
    to_out = attn.to_out
    return to_out[0] if isinstance(to_out, (list, tuple, nn.ModuleList, nn.Sequential)) else to_out


class AudioAttnProcessor(nn.Module):
    """Audio-conditioned cross-attention processor (Add-FiLM or KV-concat)."""

    def __init__(self, level: str, audio_dim: int = 768, hidden_dim: int = 768, mode: str = "add",
                 dropout: float = 0.1, bottleneck_dim: int = 64):
        super().__init__()
        self.level = level
        self.mode = mode
        # parameter containers only (indices 0 and 3 keep the reference's state-dict keys)
        self.audio_proj = nn.Sequential(nn.Linear(audio_dim, bottleneck_dim), nn.GELU(), nn.Dropout(dropout),
                                        nn.Linear(bottleneck_dim, hidden_dim))
        self.alpha = nn.Parameter(torch.zeros(1))
        self._cache = _CastCache()
        self._kv_cache: Dict[int, Any] = {}      # id(attn) -> (identity key, K/V) of the last stand-alone call

    # ---- step-invariant part -------------------------------------------------------------------
    def context(self, encoder_hidden_states: torch.Tensor, audio_tokens: Optional[torch.Tensor]) -> torch.Tensor:
        """Text states with the audio injected (reference :85-109).  Falls through unchanged when there
        is no audio for this level or the mode is unknown (reference :80,:86,:91,:99)."""
        if audio_tokens is None or self.mode not in _MODES:
            return encoder_hidden_states
        dt = encoder_hidden_states.dtype
        a = audio_tokens if audio_tokens.dtype == dt else ops.cast(audio_tokens.contiguous(), dt)
        l0, l3 = self.audio_proj[0], self.audio_proj[3]
        c = self._cache
        return ops.audio_context(encoder_hidden_states.contiguous(), a.contiguous(),
                                 c.get(l0.weight, dt, "w1"), c.get(l0.bias, torch.float32, "b1"),
                                 c.get(l3.weight, dt, "w2"), c.get(l3.bias, torch.float32, "b2"),
                                 c.get(self.alpha, torch.float32, "alpha"), _MODES[self.mode])

    def _kv_weight(self, attn, dt) -> torch.Tensor:
        wk, wv = _weight(attn.to_k).detach(), _weight(attn.to_v).detach()
        key = ("wkv", id(attn), dt)
        hit = self._cache._c.get(key)
        ver = (wk._version, wv._version, wk.data_ptr(), wv.data_ptr())
        if hit is not None and hit[0] == ver:
            return hit[1]
        w = torch.cat([self._cache.get(wk, dt, ("wk", id(attn))), self._cache.get(wv, dt, ("wv", id(attn)))], dim=0)
        self._cache._c[key] = (ver, w.contiguous())
        return self._cache._c[key][1]

    def audio_features(self, audio_tokens: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
        """audio_proj(audio_tokens) [B,K,hidden] (reference :88, eval mode: no dropout)."""
        a = audio_tokens if audio_tokens.dtype == dt else ops.cast(audio_tokens.contiguous(), dt)
        l0, l3 = self.audio_proj[0], self.audio_proj[3]
        c = self._cache
        h = ops.linear(a.contiguous(), c.get(l0.weight, dt, "w1"), c.get(l0.bias, torch.float32, "b1"), act=ops.ACT_GELU)
        return ops.linear(h, c.get(l3.weight, dt, "w2"), c.get(l3.bias, torch.float32, "b2"))

    def prepare(self, attn, encoder_hidden_states: torch.Tensor, audio: Optional[Dict[str, torch.Tensor]] = None):
        """K/V for this site: [B, T', 2C] = context(ehs, audio[level]) @ [Wk; Wv]^T (reference :120-121).
        mode="decoupled" (extension): a packed ops.XattnKV holding the text K/V and the audio branch's own K/V."""
        tokens = audio.get(self.level) if isinstance(audio, dict) else None
        if self.mode == "decoupled" and tokens is not None:
            dt = encoder_hidden_states.dtype
            wkv = self._kv_weight(attn, dt)
            C = wkv.shape[0] // 2
            if tokens.shape[1] > 16 or not ops.xattn_packable(C, attn.heads, encoder_hidden_states.shape[1], dt, tokens.shape[1]):
                raise C2DError("mode='decoupled' runs on the fused bf16 cross-attention kernel only "
                               f"(<= 16 audio tokens, <= 96 text keys; got dtype {dt}, {tokens.shape[1]} audio tokens)")
            kv = ops.linear(encoder_hidden_states.contiguous(), wkv)
            kv2 = ops.linear(self.audio_features(tokens, dt), wkv)
            lam = float(torch.sigmoid(self.alpha.detach().float()))     # once per image, outside the captured step
            return ops.xattn_pack_kv(kv, attn.heads, kv2, lambda2=lam)
        ehs = self.context(encoder_hidden_states, tokens)
        return ops.linear(ehs, self._kv_weight(attn, ehs.dtype))

    # ---- per-step part -------------------------------------------------------------------------
    supports_ln_fold = True
    supports_packed_kv = True

    def attend(self, attn, hidden_states: torch.Tensor, kv, residual: Optional[torch.Tensor] = None,
               scale: float = 1.0, ln_stats: Optional[torch.Tensor] = None,
               row_stats: Optional[torch.Tensor] = None, key_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """to_q -> softmax(q k^T d^-1/2) v -> to_out[0] (+ residual).  hidden_states [B,N,C].
        ln_stats: the engine's LayerNorm-free mode -- hidden_states is the un-normalised residual stream, norm2 is
        folded into the to_q GEMM (attn.ln_q, prepared by SD15UNet); row_stats: statistics accumulator of the output."""
        dt = hidden_states.dtype
        c = self._cache
        out_lin = _to_out_linear(attn)
        if key_mask is not None and isinstance(kv, ops.XattnKV):
            if kv.kv is None:
                raise C2DError("attention_mask with mode='decoupled' is not supported")
            kv = kv.kv                       # masked keys: the flash kernel with a key mask (the fused kernel has none)
        if isinstance(kv, ops.XattnKV) and not ops.xattn_supported(hidden_states, attn.heads, kv.T, kv.T2):
            if kv.kv is None:
                raise C2DError(f"mode='decoupled' needs the fused kernel; {hidden_states.shape[1]} tokens per sample is "
                               "outside it (multiples of 128, or 32 / 64 / 96)")
            kv = kv.kv                       # token counts outside the fused kernel (tiny latents): three-kernel path
        if isinstance(kv, ops.XattnKV):
            # packed K/V cache (SD15UNet.prepare_conditioning): to_q + attention core are ONE kernel, Q stays on chip
            sc = float(getattr(attn, "scale", (kv.C // attn.heads) ** -0.5)) * scale
            if ln_stats is not None:
                o = ops.xattn(hidden_states, kv, ln=attn.ln_q, ln_stats=ln_stats, scale=sc)
            else:
                o = ops.xattn(hidden_states, kv, wq=c.get(_weight(attn.to_q), dt, ("wq", id(attn))), scale=sc)
            bias = None if out_lin.bias is None else c.get(out_lin.bias, torch.float32, ("bo", id(attn)))
            return ops.linear(o, c.get(out_lin.weight, dt, ("wo", id(attn))), bias, residual=residual, row_stats=row_stats)
        C = kv.shape[-1] // 2
        if ln_stats is not None:
            q = ops.linear(hidden_states, None, ln=attn.ln_q, ln_stats=ln_stats)
        else:
            q = ops.linear(hidden_states, c.get(_weight(attn.to_q), dt, ("wq", id(attn))))
        d = C // attn.heads
        o = ops.attention(q, kv[..., :C], kv[..., C:], attn.heads, scale=float(getattr(attn, "scale", d ** -0.5)) * scale,
                          mask=key_mask)
        bias = None if out_lin.bias is None else c.get(out_lin.bias, torch.float32, ("bo", id(attn)))
        return ops.linear(o, c.get(out_lin.weight, dt, ("wo", id(attn))), bias, residual=residual, row_stats=row_stats)

    # ---- attention_mask (reference :129) ------------------------------------------------------
    @staticmethod
    def _key_mask(attention_mask: torch.Tensor, B: int, heads: int, T: int) -> torch.Tensor:
        """Reduce diffusers' attention mask (additive bias added to q k^T, or boolean keep-mask) to a uint8 [B,T] key
        mask.  Raises when the bias is not a pure key-padding mask."""
        m = attention_mask
        if m.dim() == 2:
            m = m[:, None, :]
        if m.dim() != 3 or m.shape[-1] != T:
            raise C2DError(f"attention_mask of shape {tuple(attention_mask.shape)} does not match {T} keys")
        if m.shape[1] != 1:
            if not bool((m == m[:, :1]).all()):
                raise C2DError("attention_mask differs per query row: only key-padding masks are supported")
            m = m[:, :1]
        if m.shape[0] == B * heads:
            mh = m.reshape(B, heads, 1, T)
            if not bool((mh == mh[:, :1]).all()):
                raise C2DError("attention_mask differs per head: only key-padding masks are supported")
            m = mh[:, 0]
        elif m.shape[0] != B:
            raise C2DError(f"attention_mask batch {m.shape[0]} is neither B={B} nor B*heads={B * heads}")
        m = m.reshape(B, T)
        if m.dtype == torch.bool:
            keep = m
        else:
            mf = m.float()
            keep = mf > -1e4
            if not bool(((mf == 0) | ~keep).all()):
                raise C2DError("attention_mask is a general additive bias: only 0 / -inf key-padding masks are supported")
        if not bool(keep.any(dim=1).all()):
            raise C2DError("attention_mask removes every key of a sample")
        return keep.to(torch.uint8).contiguous()

    def _cached_kv(self, attn, x: torch.Tensor, ehs: torch.Tensor, audio, pack: bool):
        """K/V of a site for the stand-alone call, reused while the inputs are the same tensors (same storage and
        version counters) -- what a denoising loop over a diffusers UNet passes at every step."""
        tokens = audio.get(self.level) if isinstance(audio, dict) else None
        params = [self.alpha, self.audio_proj[0].weight, self.audio_proj[0].bias, self.audio_proj[3].weight,
                  self.audio_proj[3].bias, _weight(attn.to_k), _weight(attn.to_v)]
        ident = (ehs.data_ptr(), ehs._version, tuple(ehs.shape), ehs.dtype, x.dtype,
                 None if tokens is None else (tokens.data_ptr(), tokens._version, tuple(tokens.shape), tokens.dtype),
                 self.mode, pack, tuple((t.data_ptr(), t._version) for t in params))
        hit = self._kv_cache.get(id(attn))
        if hit is not None and hit[0] == ident:
            return hit[1]
        e = ehs if ehs.dtype == x.dtype else ops.cast(ehs.contiguous(), x.dtype)
        kv = self.prepare(attn, e, audio)
        self.kv_projections = getattr(self, "kv_projections", 0) + 1     # how often K/V were actually (re)computed
        if pack and torch.is_tensor(kv) and ops.xattn_supported(x, attn.heads, kv.shape[1]) and ops.xattn_packable(
                x.shape[-1], attn.heads, kv.shape[1], x.dtype):
            kv = ops.xattn_pack_kv(kv, attn.heads)       # bf16, SD-1.5 shapes: to_q + attention core in one kernel
        # keep the source tensors alive next to the entry: a freed-and-reused address must not look like a hit
        self._kv_cache[id(attn)] = (ident, kv, ehs, tokens)
        return kv

    # ---- diffusers processor protocol ----------------------------------------------------------
    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, temb: Optional[torch.Tensor] = None,
                 scale: float = 1.0, **cross_attention_kwargs) -> torch.Tensor:
        ops.require_cuda(hidden_states, "AudioAttnProcessor")
        if getattr(attn, "spatial_norm", None) is not None or getattr(attn, "norm_cross", None):
            raise C2DError("spatial_norm / norm_cross attention variants are not on the SD-1.5 path")
        x = hidden_states
        nd = x.dim()
        if nd == 4:                                   # [B,C,H,W] -> [B,HW,C]   (reference :67-70)
            B, Cc, H, W = x.shape
            x = ops.transpose(x.reshape(B, Cc, H * W).contiguous())
        x = x.contiguous()
        if encoder_hidden_states is None:
            # reference quirk (:117-121): K/V are projected from the Q-projected states
            dt = x.dtype
            c = self._cache
            q = ops.linear(x, c.get(_weight(attn.to_q), dt, ("wq", id(attn))))
            if scale != 1.0:
                raise C2DError("scale != 1 with encoder_hidden_states=None is not supported")
            k = ops.linear(q, c.get(_weight(attn.to_k), dt, ("wk", id(attn))))
            v = ops.linear(q, c.get(_weight(attn.to_v), dt, ("wv", id(attn))))
            km = None if attention_mask is None else self._key_mask(attention_mask, q.shape[0], attn.heads, k.shape[1])
            o = ops.attention(q, k, v, attn.heads, scale=float(getattr(attn, "scale", (q.shape[-1] // attn.heads) ** -0.5)), mask=km)
            out_lin = _to_out_linear(attn)
            bias = None if out_lin.bias is None else c.get(out_lin.bias, torch.float32, ("bo", id(attn)))
            out = ops.linear(o, c.get(out_lin.weight, dt, ("wo", id(attn))), bias)
        else:
            kv = self._cached_kv(attn, x, encoder_hidden_states, cross_attention_kwargs.get("audio"), pack=attention_mask is None)
            km = None
            if attention_mask is not None:
                T = kv.T if isinstance(kv, ops.XattnKV) else kv.shape[1]
                km = self._key_mask(attention_mask, x.shape[0], attn.heads, T)
            res = x if getattr(attn, "residual_connection", False) and nd != 4 else None
            out = self.attend(attn, x, kv, residual=res, scale=scale, key_mask=km)
        if nd == 4:                                   # back to [B,C,H,W] (reference :137-138)
            out = ops.transpose(out).reshape(B, Cc, H, W)
            if getattr(attn, "residual_connection", False):
                out = ops.add(out, hidden_states.contiguous())
        if float(getattr(attn, "rescale_output_factor", 1.0)) != 1.0:
            raise C2DError("rescale_output_factor != 1 is not on the SD-1.5 path")
        return out

    forward = __call__


class AudioProcessorManager:
    """Maps one shared AudioAttnProcessor per level onto the UNet's attn2 sites (reference :148-267)."""

    LEVEL_RULES = (("mid_block", "mid"), ("down_blocks.0", "early"), ("down_blocks.1", "early"),
                   ("down_blocks.2", "late"), ("down_blocks.3", "late"), ("up_blocks.0", "late"),
                   ("up_blocks.1", "late"), ("up_blocks.2", "mid"), ("up_blocks.3", "mid"))

    def __init__(self, unet):
        self.unet = unet
        self.processors: Dict[str, Any] = {}
        self.level_mapping = self._create_level_mapping()

    def _create_level_mapping(self) -> Dict[str, list]:
        mapping: Dict[str, list] = {"early": [], "mid": [], "late": []}
        for name in self.unet.attn_processors.keys():
            if "attn1" in name:                       # self-attention keeps its processor
                continue
            level = "mid"
            for key, lvl in self.LEVEL_RULES:
                if key in name:
                    level = lvl
                    break
            mapping[level].append(name)
        return mapping

    def setup_processors(self, audio_dim: int = 768, hidden_dim: Optional[int] = None, mode: str = "add",
                         dropout: float = 0.1):
        if hidden_dim is None:
            hidden_dim = 768
            for name in self.unet.attn_processors:
                if "attn2" not in name:
                    continue
                try:
                    mod = self.unet.get_submodule(name.rsplit(".", 1)[0])
                except (AttributeError, KeyError):
                    continue
                if hasattr(mod, "to_k"):
                    w = mod.to_k.weight
                    hidden_dim = getattr(mod.to_k, "in_features", w.shape[1])
                    break
        new_processors = dict(self.unet.attn_processors)
        for level, names in self.level_mapping.items():
            proc = AudioAttnProcessor(level=level, audio_dim=audio_dim, hidden_dim=hidden_dim, mode=mode, dropout=dropout)
            for name in names:
                new_processors[name] = proc
        self.unet.set_attn_processor(new_processors)
        self.processors = new_processors
        print("Setup audio processors:")
        for lvl in ("early", "mid", "late"):
            print(f"  {lvl.capitalize()} blocks: {len(self.level_mapping[lvl])}")

    def get_audio_kwargs(self, routed_tokens: Dict[str, torch.Tensor]) -> Dict:
        return {"audio": routed_tokens}
