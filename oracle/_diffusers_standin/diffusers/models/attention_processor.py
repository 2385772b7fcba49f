"""Stand-in for diffusers.models.attention_processor (diffusers==0.23.1), ORACLE-ONLY.

Only what the reference's AudioAttnProcessor touches (models/audio_attention_processor.py:64-143):
to_q/to_k/to_v (bias-free), to_out=[Linear, Dropout], heads, scale, head_to_batch_dim,
batch_to_head_dim, get_attention_scores (baddbmm alpha=scale + softmax), and the SD-1.5 flag values
spatial_norm=None, norm_cross=None, residual_connection=False, rescale_output_factor=1.
"""
import torch
import torch.nn as nn


class AttnProcessor:
    pass


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False):
        super().__init__()
        inner = heads * dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.spatial_norm = None
        self.norm_cross = None
        self.group_norm = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = None

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def get_attention_scores(self, query, key, attention_mask=None):
        if attention_mask is None:
            base = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
            beta = 0
        else:
            base, beta = attention_mask, 1
        scores = torch.baddbmm(base, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        return scores.softmax(dim=-1).to(query.dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)
