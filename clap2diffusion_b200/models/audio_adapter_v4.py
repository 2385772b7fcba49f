"""Drop-in for the reference's ``models/audio_adapter_v4.py`` -- the "audio projector": CLAP embedding
[B,512] -> 16 audio tokens [B,16,768], plus the gated audio cross-attention layer for UNet blocks.

Same classes / constructor signatures / ``state_dict`` keys as the reference
(AudioTokenGenerator :13-119, AudioSelfAttention :122-165, AudioCrossAttention :168-261,
AudioAdapter :264-301); every forward runs on libc2d kernels (CUDA only, inference semantics:
dropout is the identity).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ._nnops import Native, require_cuda


class AudioSelfAttention(Native):
    """Multi-head self-attention over the audio tokens: bias-free fused QKV, biased output projection."""

    def __init__(self, hidden_dim: int, num_heads: int, dropout: float = 0.1):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = hidden_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.to_qkv = nn.Linear(hidden_dim, hidden_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.Dropout(dropout))

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        require_cuda(x, "AudioSelfAttention")
        c = x.shape[-1]
        qkv = self.lin(self.to_qkv, x)
        o = ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], self.num_heads, scale=self.scale)
        return self.lin(self.to_out[0], o, residual=residual)


class AudioTokenGenerator(Native):
    """Learned queries attend (single head) to K/V synthesised from the CLAP embedding, then are refined by
    ``num_layers`` pre-LN self-attention layers and an output projection + LayerNorm."""

    def __init__(self, audio_dim: int = 512, hidden_dim: int = 768, num_tokens: int = 16, num_layers: int = 4,
                 num_heads: int = 8, dropout: float = 0.1):
        super().__init__()
        self.num_tokens, self.hidden_dim = num_tokens, hidden_dim
        self.audio_queries = nn.Parameter(torch.randn(num_tokens, hidden_dim))
        self.pos_embed = nn.Parameter(torch.randn(num_tokens, hidden_dim))
        self.audio_to_kv = nn.Sequential(nn.Linear(audio_dim, 256), nn.GELU(), nn.Dropout(dropout),
                                         nn.Linear(256, hidden_dim * 2 * num_tokens))
        self.self_attn_layers = nn.ModuleList(AudioSelfAttention(hidden_dim, num_heads, dropout) for _ in range(num_layers))
        self.layer_norms = nn.ModuleList(nn.LayerNorm(hidden_dim) for _ in range(num_layers))
        self.output_proj = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.LayerNorm(hidden_dim))
        self._init_weights()

    def _init_weights(self):
        # reference :71-78 -- Xavier-uniform for the queries and every Linear, zero biases; pos_embed stays randn
        nn.init.xavier_uniform_(self.audio_queries)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, audio_embedding: torch.Tensor) -> torch.Tensor:
        require_cuda(audio_embedding, "AudioTokenGenerator")
        x = audio_embedding.contiguous()
        B, n, d = x.shape[0], self.num_tokens, self.hidden_dim
        dt = x.dtype
        # queries + positions, shared by the whole batch: [1, n, d]
        q0 = ops.bcast_add(self._w(self.audio_queries, dt), self._w(self.pos_embed, dt), 1, n, d, 0, 0)
        kv = self.lin(self.audio_to_kv[3], self.lin(self.audio_to_kv[0], x, act=ops.ACT_GELU)).view(B, n, 2, d)
        # single-head cross-attention of the shared queries onto per-sample K/V (index 0 = K, 1 = V)
        att = ops.attention(q0.expand(B, n, d), kv[:, :, 0, :], kv[:, :, 1, :], 1, scale=d ** -0.5)
        tok = ops.bcast_add(att, q0.view(n, d), B, n, d, 0, 2)
        for attn, norm in zip(self.self_attn_layers, self.layer_norms):
            tok = attn(self.ln(norm, tok), residual=tok)
        return self.ln(self.output_proj[1], self.lin(self.output_proj[0], tok))


class AudioCrossAttention(Native):
    """Gated audio cross-attention for UNet blocks: ``h + sigmoid(gate) * Attn(LN(h), audio)`` with 8 x 64
    heads regardless of the query width (the "decoupled audio branch")."""

    def __init__(self, query_dim: int, context_dim: int = 768, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, gate_init: float = -5.0):
        super().__init__()
        inner = dim_head * heads
        self.scale, self.heads = dim_head ** -0.5, heads
        self.norm = nn.LayerNorm(query_dim)
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))
        self.gate = nn.Parameter(torch.tensor(gate_init))

    def forward(self, hidden_states: torch.Tensor, audio_context: torch.Tensor,
                attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        require_cuda(hidden_states, "AudioCrossAttention")
        h = hidden_states.contiguous()
        a = audio_context.contiguous()
        if a.dtype != h.dtype:
            a = ops.cast(a, h.dtype)
        B, nk = a.shape[0], a.shape[1]
        mask = None
        if attention_mask is not None:
            # the reference broadcasts a boolean mask over heads and queries (masked_fill on [B,h,N,K], :244-245)
            if attention_mask.numel() != B * nk:
                raise NotImplementedError("only key-padding masks broadcastable from [B,1,1,K] are supported")
            mask = attention_mask.reshape(B, nk).to(torch.uint8).contiguous()
        q = self.lin(self.to_q, self.ln(self.norm, h))
        k, v = self.lin(self.to_k, a), self.lin(self.to_v, a)
        o = ops.attention(q, k, v, self.heads, scale=self.scale, mask=mask)
        # out = h + sigmoid(gate) * (o Wo^T + bo): fold the gate into the projection weights (scalar)
        g = float(torch.sigmoid(self.gate.detach().float()))
        lin = self.to_out[0]
        key = ("gated", g, h.dtype)
        hit = self._cc._c.get(key)
        ver = (lin.weight._version, lin.weight.data_ptr(), lin.bias._version)
        if hit is None or hit[0] != ver:
            wg = (lin.weight.detach().float() * g).to(h.dtype).contiguous()
            bg = (lin.bias.detach().float() * g).contiguous()
            self._cc._c[key] = (ver, (wg, bg))
        wg, bg = self._cc._c[key][1]
        return ops.linear(o, wg, bg, residual=h)


class AudioAdapter(Native):
    """CLAP embedding [B, audio_dim] -> audio tokens [B, num_tokens, hidden_dim]."""

    def __init__(self, audio_dim: int = 512, hidden_dim: int = 768, num_tokens: int = 16, num_layers: int = 4,
                 num_heads: int = 8, dropout: float = 0.1):
        super().__init__()
        self.token_generator = AudioTokenGenerator(audio_dim=audio_dim, hidden_dim=hidden_dim, num_tokens=num_tokens,
                                                   num_layers=num_layers, num_heads=num_heads, dropout=dropout)

    def forward(self, audio_embedding: torch.Tensor) -> torch.Tensor:
        return self.token_generator(audio_embedding)
