"""Drop-in for the reference's ``models/hierarchical_audio_v4.py``: foreground / background / ambience
decomposition of the CLAP embedding, soft level assignment, adaptive level weights, routing to the
early / mid / late UNet scales and the Perceiver-style 10 -> 77 token projector.

Class names, constructor signatures, ``state_dict`` keys and buffers follow the reference
(TemperatureScheduler :20-76, SoftHierarchicalDecomposition :79-238, AdaptiveHierarchyWeights :241-290,
LevelToUNetRouter :293-369, CrossAttentionBlock :375-414, AudioProjectionTransformer77 :417-492,
CrossHierarchyAttention :495-591, ImprovedHierarchicalAudioEncoder :594-772, legacy
HierarchicalAudioDecomposition :776-882 and HierarchicalAudioV4 :885-932).  Forward passes run on
libc2d kernels (CUDA only; dropout is the identity).  ``ImprovedHierarchicalAudioEncoder.encode`` is the
sync-free entry the sampler uses; ``forward`` keeps the reference's return structure.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ._nnops import Native, require_cuda


class TemperatureScheduler:
    """Anneals the decomposer's softmax temperature from T_max to T_min (cosine or linear) after a warm-up."""

    def __init__(self, decomposer: nn.Module, T_max: float = 2.0, T_min: float = 0.5, total_steps: int = 5000,
                 warmup_steps: int = 200, mode: str = "cosine"):
        self.decomposer = decomposer
        self.T_max, self.T_min = T_max, T_min
        self.total_steps, self.warmup_steps, self.mode = total_steps, warmup_steps, mode
        decomposer.set_temperature(T_max)

    def temperature_at(self, current_step: int) -> float:
        if current_step < self.warmup_steps:
            return self.T_max
        if current_step >= self.total_steps or self.total_steps <= self.warmup_steps:
            return self.T_min
        frac = (current_step - self.warmup_steps) / (self.total_steps - self.warmup_steps)
        span = self.T_max - self.T_min
        if self.mode == "cosine":
            return self.T_min + span * 0.5 * (1.0 + math.cos(math.pi * frac))
        if self.mode == "linear":
            return self.T_max - span * frac
        raise ValueError(f"Unknown annealing mode: {self.mode}")

    def step(self, current_step: int):
        self.decomposer.set_temperature(self.temperature_at(current_step))


class CrossHierarchyAttention(Native):
    """Self-attention + MLP across the hierarchy tokens in a bottleneck space, with an outer residual."""

    def __init__(self, dim: int, num_heads: int = 4, dropout: float = 0.1, bottleneck_dim: int = 256,
                 mlp_ratio: float = 2.0):
        super().__init__()
        if bottleneck_dim % num_heads != 0:
            raise ValueError(f"bottleneck_dim ({bottleneck_dim}) must be divisible by num_heads ({num_heads})")
        self.dim, self.bottleneck_dim, self.num_heads = dim, bottleneck_dim, num_heads
        self.head_dim = bottleneck_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.input_proj = nn.Linear(dim, bottleneck_dim)
        self.norm1 = nn.LayerNorm(bottleneck_dim)
        self.qkv = nn.Linear(bottleneck_dim, bottleneck_dim * 3, bias=True)
        self.attn_drop = nn.Dropout(dropout)
        self.proj = nn.Linear(bottleneck_dim, bottleneck_dim)
        self.proj_drop = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(bottleneck_dim)
        hidden = int(bottleneck_dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(bottleneck_dim, hidden), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden, bottleneck_dim), nn.Dropout(dropout))
        self.output_proj = nn.Linear(bottleneck_dim, dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        require_cuda(x, "CrossHierarchyAttention")
        x = x.contiguous()
        c = self.bottleneck_dim
        h = self.lin(self.input_proj, x)
        qkv = self.lin(self.qkv, self.ln(self.norm1, h))
        o = ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], self.num_heads, scale=self.scale)
        h = self.lin(self.proj, o, residual=h)
        m = self.lin(self.mlp[0], self.ln(self.norm2, h), act=ops.ACT_GELU)
        h = self.lin(self.mlp[3], m, residual=h)
        return self.lin(self.output_proj, h, residual=x)


class SoftHierarchicalDecomposition(Native):
    """CLAP embedding -> K tokens with soft (temperature-scaled) assignment to L semantic levels."""

    def __init__(self, audio_dim: int = 512, text_dim: int = 768, num_tokens: int = 10, num_levels: int = 3,
                 dropout: float = 0.1, initial_temperature: float = 2.0):
        super().__init__()
        self.audio_dim, self.text_dim, self.num_tokens, self.num_levels = audio_dim, text_dim, num_tokens, num_levels
        self.shared_mlp = nn.Sequential(nn.Linear(audio_dim, 512), nn.GELU(), nn.LayerNorm(512), nn.Dropout(dropout),
                                        nn.Linear(512, text_dim))
        self.token_offsets = nn.Parameter(torch.randn(num_tokens, text_dim) * 0.02)
        self.level_anchors = nn.Parameter(torch.randn(num_levels, text_dim) * 0.02)
        self.gating_head = nn.Sequential(nn.Linear(text_dim, 10), nn.GELU(), nn.Linear(10, num_levels))
        self.register_buffer("temperature", torch.tensor(initial_temperature))
        self.register_buffer("level_prior", torch.tensor([5.0, 3.0, 2.0]) / 10.0)
        self.cross_hierarchy_attn = CrossHierarchyAttention(dim=text_dim, num_heads=4, dropout=dropout,
                                                            bottleneck_dim=192, mlp_ratio=1.5)
        self.norm = nn.LayerNorm(text_dim)

    @torch.no_grad()
    def set_temperature(self, temperature: float):
        self.temperature.fill_(max(temperature, 0.1))

    def compute_assignments(self, tokens: torch.Tensor) -> torch.Tensor:
        """[B,K,D] -> fp32 [B,K,L] soft assignment probabilities (one fused kernel)."""
        dt = tokens.dtype
        g0, g2 = self.gating_head[0], self.gating_head[2]
        return ops.hier_assign(tokens.contiguous(), self._w(self.level_anchors, dt), self._w(g0.weight, dt),
                               self._f(g0.bias), self._w(g2.weight, dt), self._f(g2.bias),
                               self._f(self.temperature).reshape(1))

    def tokens_and_assignments(self, audio_features: torch.Tensor):
        x = audio_features.contiguous()
        B, K, D = x.shape[0], self.num_tokens, self.text_dim
        s = self.lin(self.shared_mlp[4], self.ln(self.shared_mlp[2], self.lin(self.shared_mlp[0], x, act=ops.ACT_GELU)))
        tokens = ops.bcast_add(s, self._w(self.token_offsets, x.dtype), B, K, D, 1, 2)
        assignments = self.compute_assignments(tokens)
        tokens_out = self.ln(self.norm, self.cross_hierarchy_attn(tokens))
        return tokens_out, assignments

    def forward(self, audio_features: torch.Tensor, return_stats: bool = False) -> Tuple[torch.Tensor, Dict]:
        require_cuda(audio_features, "SoftHierarchicalDecomposition")
        tokens_out, assignments = self.tokens_and_assignments(audio_features)
        info = {"tokens": tokens_out, "assignments": assignments, "temperature": self.temperature.item(),
                "level_anchors": self.level_anchors}
        if return_stats:       # monitoring only (host syncs, like the reference :229-236)
            with torch.no_grad():
                entropy = -(assignments * (assignments + 1e-8).log()).sum(dim=-1).mean()
                info["stats"] = {"avg_assignment": assignments.mean(dim=[0, 1]), "entropy": entropy.item(),
                                 "effective_levels": torch.exp(entropy).item()}
        return tokens_out, info


class AdaptiveHierarchyWeights(Native):
    """Per-sample level weights softmax(MLP(clap)) [B,L] (or global learnable weights)."""

    def __init__(self, audio_dim: int = 512, hidden_dim: int = 6, num_levels: int = 3, use_audio_context: bool = True):
        super().__init__()
        self.num_levels, self.use_audio_context = num_levels, use_audio_context
        if use_audio_context:
            self.weight_network = nn.Sequential(nn.Linear(audio_dim, hidden_dim), nn.GELU(), nn.LayerNorm(hidden_dim),
                                                nn.Linear(hidden_dim, num_levels))
        else:
            self.weights = nn.Parameter(torch.tensor([0.5, 0.3, 0.2]))

    def forward(self, audio_features: torch.Tensor) -> torch.Tensor:
        require_cuda(audio_features, "AdaptiveHierarchyWeights")
        x = audio_features.contiguous()
        if self.use_audio_context:
            net = self.weight_network
            logits = self.lin(net[3], self.ln(net[2], self.lin(net[0], x, act=ops.ACT_GELU)))
            w = ops.softmax_rows(logits)
        else:
            w = ops.softmax_rows(self._w(self.weights, x.dtype).reshape(1, -1)).expand(x.shape[0], -1).contiguous()
        return w if w.dtype == torch.float32 else ops.cast(w, torch.float32)


class LevelToUNetRouter(Native):
    """Routes level assignments to the early / mid / late UNet scales."""

    def __init__(self, num_levels: int = 3, text_dim: int = 768):
        super().__init__()
        self.num_levels, self.text_dim = num_levels, text_dim
        self.level_gates = nn.ParameterDict({k: nn.Parameter(torch.zeros(1)) for k in ("early", "mid", "late")})
        self.routing_matrix = nn.Parameter(torch.tensor([[0.1, 0.3, 0.6], [0.2, 0.6, 0.2], [0.6, 0.3, 0.1]]))

    def forward(self, tokens: torch.Tensor, assignments: torch.Tensor,
                hierarchy_weights: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        require_cuda(tokens, "LevelToUNetRouter")
        gates = torch.cat([self.level_gates[k].detach().float().reshape(1) for k in ("early", "mid", "late")])
        a = assignments if assignments.dtype == torch.float32 else ops.cast(assignments.contiguous(), torch.float32)
        hw = None if hierarchy_weights is None else hierarchy_weights.float().contiguous()
        e, m, l = ops.hier_route(tokens.contiguous(), a.contiguous(), hw, self._f(self.routing_matrix), gates)
        return {"early": e, "mid": m, "late": l}


class CrossAttentionBlock(Native):
    """One Perceiver decoder block: pre-LN multi-head cross-attention (+res) and a 2x FFN (+res)."""

    def __init__(self, d_model: int, num_heads: int, dropout: float = 0.1):
        super().__init__()
        self.ln_q = nn.LayerNorm(d_model)
        self.ln_kv = nn.LayerNorm(d_model)
        self.cross_attn = nn.MultiheadAttention(embed_dim=d_model, num_heads=num_heads, dropout=dropout, batch_first=True)
        self.ffn = nn.Sequential(nn.LayerNorm(d_model), nn.Linear(d_model, d_model * 2), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(d_model * 2, d_model), nn.Dropout(dropout))

    def forward(self, queries: torch.Tensor, keys_values: torch.Tensor) -> torch.Tensor:
        require_cuda(queries, "CrossAttentionBlock")
        mha = self.cross_attn
        E = mha.embed_dim
        w, b = mha.in_proj_weight, mha.in_proj_bias
        q = self.lin_w(w[:E], b[:E], self.ln(self.ln_q, queries))
        kvn = self.ln(self.ln_kv, keys_values)
        kv = self.lin_w(w[E:], b[E:], kvn)                       # fused K,V projection [B, n, 2E]
        o = ops.attention(q, kv[..., :E], kv[..., E:], mha.num_heads)
        queries = self.lin(mha.out_proj, o, residual=queries.contiguous())
        h = self.lin(self.ffn[1], self.ln(self.ffn[0], queries), act=ops.ACT_GELU)
        return self.lin(self.ffn[4], h, residual=queries)


class AudioProjectionTransformer77(Native):
    """10 hierarchical audio tokens -> 77 CLIP-shaped tokens via learned queries and cross-attention blocks."""

    def __init__(self, audio_dim: int = 768, clip_dim: int = 768, bottleneck_dim: int = 256, num_heads: int = 8,
                 num_layers: int = 4, dropout: float = 0.1) -> None:
        super().__init__()
        self.audio_dim, self.clip_dim, self.bottleneck_dim = audio_dim, clip_dim, bottleneck_dim
        self.audio_proj = nn.Linear(audio_dim, bottleneck_dim)
        self.queries = nn.Parameter(torch.randn(77, bottleneck_dim) * 0.02)
        self.query_pos = nn.Parameter(torch.zeros(77, bottleneck_dim))
        self.blocks = nn.ModuleList(CrossAttentionBlock(bottleneck_dim, num_heads, dropout) for _ in range(num_layers))
        self.out_proj = nn.Linear(bottleneck_dim, clip_dim)
        self.out_norm = nn.LayerNorm(clip_dim)
        self.clip_pos_embed = nn.Parameter(torch.zeros(1, 77, clip_dim))
        nn.init.trunc_normal_(self.clip_pos_embed, std=0.02)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        require_cuda(x, "AudioProjectionTransformer77")
        x = x.contiguous()
        B, dt, E = x.shape[0], x.dtype, self.bottleneck_dim
        feats = self.lin(self.audio_proj, x)
        q = ops.bcast_add(self._w(self.queries, dt), self._w(self.query_pos, dt), B, 77, E, 2, 2)
        for blk in self.blocks:
            q = blk(q, feats)
        out = self.lin(self.out_proj, q)
        out = ops.bcast_add(out, self._w(self.clip_pos_embed, dt), B, 77, self.clip_dim, 0, 2)
        return self.ln(self.out_norm, out)


class ImprovedHierarchicalAudioEncoder(Native):
    """Soft decomposition + adaptive weights + router + 77-token projector."""

    def __init__(self, audio_dim: int = 512, text_dim: int = 768, num_tokens: int = 10, num_levels: int = 3,
                 out_tokens: int = 77, dropout: float = 0.1, use_adaptive_weights: bool = True,
                 use_soft_decomposition: bool = True):
        super().__init__()
        self.use_soft_decomposition = use_soft_decomposition
        if use_soft_decomposition:
            self.decomposer = SoftHierarchicalDecomposition(audio_dim=audio_dim, text_dim=text_dim, num_tokens=num_tokens,
                                                            num_levels=num_levels, dropout=dropout)
        else:
            self.decomposer = HierarchicalAudioDecomposition(audio_dim=audio_dim, text_dim=text_dim, dropout=dropout)
        self.adaptive_weights = (AdaptiveHierarchyWeights(audio_dim=audio_dim, hidden_dim=6, num_levels=num_levels,
                                                          use_audio_context=True) if use_adaptive_weights else None)
        self.router = LevelToUNetRouter(num_levels=num_levels, text_dim=text_dim)
        self.projector = AudioProjectionTransformer77(audio_dim=text_dim, clip_dim=text_dim, bottleneck_dim=256,
                                                      num_heads=8, num_layers=4)
        self.temperature_scheduler = None

    def encode(self, audio_features: torch.Tensor, with_tokens77: bool = True) -> Dict[str, torch.Tensor]:
        """Sync-free inference entry: dict(tokens_10, assignments, hierarchy_weights, routed{early,mid,late}
        [, tokens_77]).  No host reads, no losses -- safe inside the sampling loop / a CUDA graph."""
        require_cuda(audio_features, "ImprovedHierarchicalAudioEncoder")
        if self.use_soft_decomposition:
            tokens_10, assignments = self.decomposer.tokens_and_assignments(audio_features)
        else:
            tokens_10 = self.decomposer(audio_features)
            assignments = torch.zeros(tokens_10.shape[0], tokens_10.shape[1], 3, device=tokens_10.device)
        hw = self.adaptive_weights(audio_features) if self.adaptive_weights is not None else None
        out = {"tokens_10": tokens_10, "assignments": assignments, "hierarchy_weights": hw,
               "routed": self.router(tokens_10, assignments, hw)}
        if with_tokens77:
            out["tokens_77"] = self.projector(tokens_10)
        return out

    def compute_losses(self, assignments: torch.Tensor, tokens: torch.Tensor,
                       hierarchy_weights: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Stage-2 regularisers (training only, not on the inference hot path; reference :661-711).
        Plain tensor algebra on the tiny [B,K,L] / [B,K,K] statistics."""
        a = assignments.float()
        t = F.normalize(tokens.float(), p=2, dim=-1)
        eye = torch.eye(t.shape[1], device=t.device).expand(t.shape[0], -1, -1)
        losses = {"entropy": -(a * (a + 1e-8).log()).sum(dim=-1).mean(),
                  "orthogonality": F.mse_loss(torch.bmm(t, t.transpose(1, 2)), eye)}
        if self.use_soft_decomposition and hasattr(self.decomposer, "level_prior"):
            avg = a.mean(dim=1)
            prior = self.decomposer.level_prior.unsqueeze(0).expand_as(avg)
            losses["prior"] = F.kl_div(prior.log(), avg, reduction="batchmean")
        else:
            losses["prior"] = torch.tensor(0.0, device=tokens.device)
        return losses

    def forward(self, audio_features: torch.Tensor,
                return_all: bool = False) -> Union[torch.Tensor, Tuple[torch.Tensor, Dict]]:
        enc = self.encode(audio_features)
        if not return_all:
            return enc["tokens_77"]
        a = enc["assignments"]
        stats, temperature = {}, 1.0
        if self.use_soft_decomposition:
            with torch.no_grad():
                ent = -(a * (a + 1e-8).log()).sum(dim=-1).mean()
                stats = {"avg_assignment": a.mean(dim=[0, 1]), "entropy": ent.item(),
                         "effective_levels": torch.exp(ent).item()}
            temperature = self.decomposer.temperature.item()
        info = {"tokens_10": enc["tokens_10"], "tokens_77": enc["tokens_77"], "assignments": a,
                "routed": enc["routed"], "hierarchy_weights": enc["hierarchy_weights"],
                "losses": self.compute_losses(a, enc["tokens_10"], enc["hierarchy_weights"]),
                "stats": stats, "temperature": temperature}
        return enc["tokens_77"], info


class HierarchicalAudioDecomposition(Native):
    """Legacy rigid 5-3-2 decomposition (what ``scripts/inference.py`` instantiates through HierarchicalAudioV4)."""

    def __init__(self, audio_dim: int = 512, text_dim: int = 768, num_foreground: int = 5, num_background: int = 3,
                 num_ambience: int = 2, dropout: float = 0.1):
        super().__init__()
        self.audio_dim, self.text_dim = audio_dim, text_dim
        self.num_foreground, self.num_background, self.num_ambience = num_foreground, num_background, num_ambience
        self.total_tokens = num_foreground + num_background + num_ambience

        def head(hidden, n):
            return nn.Sequential(nn.Linear(audio_dim, hidden), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden, text_dim * n))
        self.foreground_proj = head(text_dim * 2, num_foreground)
        self.background_proj = head(text_dim, num_background)
        self.ambience_proj = head(text_dim // 2, num_ambience)
        self.hierarchy_weights = nn.Parameter(torch.tensor([0.5, 0.3, 0.2], dtype=torch.float32))
        self.layer_norm = nn.LayerNorm(text_dim)
        self.cross_hierarchy_attn = CrossHierarchyAttention(text_dim, num_heads=4, dropout=dropout, bottleneck_dim=192)

    def forward(self, audio_features: torch.Tensor,
                return_hierarchy: bool = False) -> Union[torch.Tensor, Tuple[torch.Tensor, Dict]]:
        require_cuda(audio_features, "HierarchicalAudioDecomposition")
        x = audio_features.contiguous()
        parts = [self.lin(p[3], self.lin(p[0], x, act=ops.ACT_GELU))
                 for p in (self.foreground_proj, self.background_proj, self.ambience_proj)]
        cat = ops.legacy_combine(parts[0], parts[1], parts[2], self._f(self.hierarchy_weights), self.text_dim)
        tokens = self.ln(self.layer_norm, self.cross_hierarchy_attn(cat))
        if not return_hierarchy:
            return tokens
        nf, nb = self.num_foreground, self.num_background
        weights = ops.softmax_rows(self._f(self.hierarchy_weights).reshape(1, -1)).reshape(-1)
        return tokens, {"foreground": cat[:, :nf], "background": cat[:, nf:nf + nb], "ambience": cat[:, nf + nb:],
                        "weights": weights, "combined": tokens}


class HierarchicalAudioV4(Native):
    """Legacy Stage-1 encoder: rigid decomposition + projection to 77 tokens."""

    def __init__(self, audio_dim: int = 512, text_dim: int = 768, num_foreground: int = 5, num_background: int = 3,
                 num_ambience: int = 2, out_tokens: int = 77, projector_layers: int = 4, projector_heads: int = 8,
                 projector_mlp_ratio: float = 4.0, dropout: float = 0.1) -> None:
        super().__init__()
        self.decomposer = HierarchicalAudioDecomposition(audio_dim=audio_dim, text_dim=text_dim,
                                                         num_foreground=num_foreground, num_background=num_background,
                                                         num_ambience=num_ambience, dropout=dropout)
        self.projector = AudioProjectionTransformer77(audio_dim=text_dim, clip_dim=text_dim, bottleneck_dim=256,
                                                      num_heads=projector_heads, num_layers=projector_layers,
                                                      dropout=dropout)

    def forward(self, clap_features: torch.Tensor, return_intermediate: bool = False):
        tokens10, hierarchy = self.decomposer(clap_features, return_hierarchy=True)
        tokens77 = self.projector(tokens10)
        if return_intermediate:
            hierarchy = dict(hierarchy)
            hierarchy["tokens10"] = tokens10
            return tokens77, hierarchy
        return tokens77

    def encode(self, clap_features: torch.Tensor, with_tokens77: bool = True) -> Dict[str, torch.Tensor]:
        """Conditioning entry with the interface of ImprovedHierarchicalAudioEncoder.encode, so that a stage-1/2/3
        checkpoint of THIS class (the only hierarchical model the reference's scripts save: train_stage2.py:183-189,
        train_stage3.py:262-271) can drive the audio attention processors.  The rigid 5-3-2 groups go to the UNet
        levels the reference names for them (models/hierarchical_audio_v4.py:311-313, :319-321): ambience -> early
        blocks, background -> mid blocks, foreground -> late blocks."""
        tokens10, hierarchy = self.decomposer(clap_features, return_hierarchy=True)
        out = {"tokens_10": tokens10,
               "routed": {"early": hierarchy["ambience"].contiguous(), "mid": hierarchy["background"].contiguous(),
                          "late": hierarchy["foreground"].contiguous()}}
        if with_tokens77:
            out["tokens_77"] = self.projector(tokens10)
        return out
