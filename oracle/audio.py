"""fp32 restatement of the reference's audio-conditioning modules.  TEST INFRASTRUCTURE
(see oracle/__init__.py).

Functional style: ``fn(sd, x)`` where ``sd`` maps the reference's state-dict keys
to fp32 tensors.  Pinned against the unmodified reference modules by ``oracle/make_golden.py``
(vectors under ``tests/golden/``).  All file:line citations are into /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .weights import P, linear, norm

T = torch.Tensor


def _lin(sd, name, x):
    return F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"))


def _ln(sd, name, x, eps=1e-5):
    w = sd[f"{name}.weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[f"{name}.bias"], eps)


def _heads_attention(q: T, k: T, v: T, heads: int) -> T:
    """softmax(q k^T d^-1/2) v, tensors [B, N, heads*d]."""
    B, Nq, C = q.shape
    d = C // heads
    q = q.reshape(B, Nq, heads, d).transpose(1, 2)
    k = k.reshape(B, -1, heads, d).transpose(1, 2)
    v = v.reshape(B, -1, heads, d).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, Nq, C)


# =======================================================================================
# AudioAdapter / AudioTokenGenerator  (models/audio_adapter_v4.py:13-119, 122-165, 264-301)
# =======================================================================================
def audio_adapter_spec(audio_dim=512, hidden=768, tokens=16, layers=4) -> List[P]:
    g = "token_generator"
    xav = lambda fi, fo: math.sqrt(2.0 / (fi + fo))      # xavier std (audio_adapter_v4.py:71-78)
    s = [P(f"{g}.audio_queries", (tokens, hidden), "emb", scale=xav(hidden, tokens)),
         P(f"{g}.pos_embed", (tokens, hidden), "emb", scale=1.0)]
    s += linear(f"{g}.audio_to_kv.0", audio_dim, 256) + linear(f"{g}.audio_to_kv.3", 256, hidden * 2 * tokens)
    for i in range(layers):
        s += linear(f"{g}.self_attn_layers.{i}.to_qkv", hidden, 3 * hidden, False)
        s += linear(f"{g}.self_attn_layers.{i}.to_out.0", hidden, hidden)
    for i in range(layers):
        s += norm(f"{g}.layer_norms.{i}", hidden)
    s += linear(f"{g}.output_proj.0", hidden, hidden) + norm(f"{g}.output_proj.1", hidden)
    return s


def audio_adapter_forward(sd: Dict[str, T], clap: T, heads: int = 8) -> T:
    """[B,512] -> [B,16,768]  (audio_adapter_v4.py:80-119; AudioSelfAttention :141-165)."""
    g = "token_generator"
    q0 = sd[f"{g}.audio_queries"] + sd[f"{g}.pos_embed"]                 # :91-93
    n, d = q0.shape
    B = clap.shape[0]
    kv = _lin(sd, f"{g}.audio_to_kv.3", F.gelu(_lin(sd, f"{g}.audio_to_kv.0", clap)))   # :96
    kv = kv.view(B, n, 2, d)                                              # :97 (index 0 = K, 1 = V)
    k, v = kv[:, :, 0], kv[:, :, 1]
    sc = torch.einsum("nd,bmd->bnm", q0, k) * d ** -0.5                   # :101-103 single head
    tok = torch.softmax(sc, dim=-1) @ v + q0[None]                        # :104-108
    i = 0
    while f"{g}.self_attn_layers.{i}.to_qkv.weight" in sd:               # :111-114
        x = _ln(sd, f"{g}.layer_norms.{i}", tok)
        qkv = _lin(sd, f"{g}.self_attn_layers.{i}.to_qkv", x)
        qq, kk, vv = qkv.chunk(3, dim=-1)
        o = _heads_attention(qq, kk, vv, heads)
        tok = _lin(sd, f"{g}.self_attn_layers.{i}.to_out.0", o) + tok
        i += 1
    return _ln(sd, f"{g}.output_proj.1", _lin(sd, f"{g}.output_proj.0", tok))    # :117


def norm60(tokens: T, target: float = 60.0, per_sample: bool = False) -> T:
    """scripts/inference.py:92-99: x * target / mean(||x||_2 over last dim) (mean over batch AND
    tokens).  ``per_sample=True`` is decision D3: mean over tokens only, so results do
    not depend on how the batch is sharded; identical for batch 1."""
    nrm = tokens.norm(dim=-1, keepdim=True)
    m = nrm.mean(dim=(1, 2), keepdim=True) if per_sample else nrm.mean()
    scale = torch.where(m > 0, target / m, torch.ones_like(m))
    return tokens * scale


# =======================================================================================
# CrossHierarchyAttention (hierarchical_audio_v4.py:495-591)
# =======================================================================================
def cross_hierarchy_spec(name: str, dim=768, bott=192, mlp_hidden=288) -> List[P]:
    s = linear(f"{name}.input_proj", dim, bott) + norm(f"{name}.norm1", bott)
    s += linear(f"{name}.qkv", bott, 3 * bott) + linear(f"{name}.proj", bott, bott) + norm(f"{name}.norm2", bott)
    s += linear(f"{name}.mlp.0", bott, mlp_hidden) + linear(f"{name}.mlp.3", mlp_hidden, bott)
    s += linear(f"{name}.output_proj", bott, dim)
    return s


def cross_hierarchy_forward(sd, name: str, x: T, heads: int = 4) -> T:
    h = _lin(sd, f"{name}.input_proj", x)                                 # :558
    qkv = _lin(sd, f"{name}.qkv", _ln(sd, f"{name}.norm1", h))           # :562-566
    q, k, v = qkv.chunk(3, dim=-1)      # reshape(B,N,3,H,d) == chunk(3) then split heads
    h = h + _lin(sd, f"{name}.proj", _heads_attention(q, k, v, heads))    # :569-580
    m = _lin(sd, f"{name}.mlp.3", F.gelu(_lin(sd, f"{name}.mlp.0", _ln(sd, f"{name}.norm2", h))))
    h = h + m                                                             # :583-586
    return x + _lin(sd, f"{name}.output_proj", h)                         # :589-591


# =======================================================================================
# AudioProjectionTransformer77 (hierarchical_audio_v4.py:417-492; CrossAttentionBlock :375-414)
# =======================================================================================
def projector77_spec(name: str, audio_dim=768, clip_dim=768, bott=256, layers=4) -> List[P]:
    s = linear(f"{name}.audio_proj", audio_dim, bott)
    s += [P(f"{name}.queries", (77, bott), "emb", scale=0.02),
          P(f"{name}.query_pos", (77, bott), "emb", scale=0.02)]
    for i in range(layers):
        b = f"{name}.blocks.{i}"
        s += norm(f"{b}.ln_q", bott) + norm(f"{b}.ln_kv", bott)
        s += [P(f"{b}.cross_attn.in_proj_weight", (3 * bott, bott), "w"),
              P(f"{b}.cross_attn.in_proj_bias", (3 * bott,), "b", fan_in=bott)]
        s += linear(f"{b}.cross_attn.out_proj", bott, bott)
        s += norm(f"{b}.ffn.0", bott) + linear(f"{b}.ffn.1", bott, 2 * bott) + linear(f"{b}.ffn.4", 2 * bott, bott)
    s += linear(f"{name}.out_proj", bott, clip_dim) + norm(f"{name}.out_norm", clip_dim)
    s += [P(f"{name}.clip_pos_embed", (1, 77, clip_dim), "emb", scale=0.02)]
    return s


def projector77_forward(sd, name: str, x: T, heads: int = 8) -> T:
    """[B,10,768] -> [B,77,768]."""
    B = x.shape[0]
    a = _lin(sd, f"{name}.audio_proj", x)                                 # :476
    q = (sd[f"{name}.queries"] + sd[f"{name}.query_pos"])[None].expand(B, -1, -1)   # :479-480
    E = q.shape[-1]
    i = 0
    while f"{name}.blocks.{i}.ln_q.weight" in sd:
        b = f"{name}.blocks.{i}"
        w, bias = sd[f"{b}.cross_attn.in_proj_weight"], sd[f"{b}.cross_attn.in_proj_bias"]
        qn, kvn = _ln(sd, f"{b}.ln_q", q), _ln(sd, f"{b}.ln_kv", a)      # :408-409
        qq = F.linear(qn, w[:E], bias[:E])
        kk = F.linear(kvn, w[E:2 * E], bias[E:2 * E])
        vv = F.linear(kvn, w[2 * E:], bias[2 * E:])
        q = q + _lin(sd, f"{b}.cross_attn.out_proj", _heads_attention(qq, kk, vv, heads))   # :410-411
        f = _lin(sd, f"{b}.ffn.4", F.gelu(_lin(sd, f"{b}.ffn.1", _ln(sd, f"{b}.ffn.0", q))))
        q = q + f                                                         # :414
        i += 1
    out = _lin(sd, f"{name}.out_proj", q) + sd[f"{name}.clip_pos_embed"]  # :487-490
    return _ln(sd, f"{name}.out_norm", out)


# =======================================================================================
# ImprovedHierarchicalAudioEncoder (hierarchical_audio_v4.py:594-772)
# =======================================================================================
def improved_hier_spec(audio_dim=512, text_dim=768, tokens=10, levels=3) -> List[P]:
    d = "decomposer"
    s = [P(f"{d}.token_offsets", (tokens, text_dim), "emb", scale=0.02),
         P(f"{d}.level_anchors", (levels, text_dim), "emb", scale=0.02)]
    s += linear(f"{d}.shared_mlp.0", audio_dim, 512) + norm(f"{d}.shared_mlp.2", 512)
    s += linear(f"{d}.shared_mlp.4", 512, text_dim)
    s += linear(f"{d}.gating_head.0", text_dim, 10) + linear(f"{d}.gating_head.2", 10, levels)
    s += cross_hierarchy_spec(f"{d}.cross_hierarchy_attn", text_dim, 192, 288)
    s += norm(f"{d}.norm", text_dim)
    a = "adaptive_weights.weight_network"
    s += linear(f"{a}.0", audio_dim, 6) + norm(f"{a}.2", 6) + linear(f"{a}.3", 6, levels)
    s += [P("router.routing_matrix", (3, 3), "scalar"),
          P("router.level_gates.early", (1,), "scalar"),
          P("router.level_gates.mid", (1,), "scalar"),
          P("router.level_gates.late", (1,), "scalar")]
    s += projector77_spec("projector", text_dim, text_dim)
    return s


IMPROVED_BUFFERS = {"decomposer.temperature": 2.0, "decomposer.level_prior": (0.5, 0.3, 0.2)}


def improved_hier_forward(sd, clap: T, temperature: Optional[float] = None) -> Dict[str, T]:
    """[B,512] -> dict(tokens_77 [B,77,768], tokens_10, assignments [B,10,3], hierarchy_weights [B,3],
    routed{early,mid,late} [B,10,768])  (forward :713-772 with return_all=True)."""
    d = "decomposer"
    temp = float(sd[f"{d}.temperature"]) if temperature is None else temperature
    # SoftHierarchicalDecomposition.forward :184-238
    s = _lin(sd, f"{d}.shared_mlp.0", clap)
    s = _lin(sd, f"{d}.shared_mlp.4", _ln(sd, f"{d}.shared_mlp.2", F.gelu(s)))     # :203
    tokens = s[:, None, :] + sd[f"{d}.token_offsets"][None]                         # :204-205
    # compute_assignments :154-182
    tn = F.normalize(tokens, p=2, dim=-1)
    an = F.normalize(sd[f"{d}.level_anchors"], p=2, dim=-1)
    sim = torch.einsum("bkd,ld->bkl", tn, an) * 10.0
    gate = _lin(sd, f"{d}.gating_head.2", F.gelu(_lin(sd, f"{d}.gating_head.0", tokens)))
    assign = torch.softmax((sim + gate) / temp, dim=-1)
    tok10 = _ln(sd, f"{d}.norm", cross_hierarchy_forward(sd, f"{d}.cross_hierarchy_attn", tokens))  # :211-212
    # AdaptiveHierarchyWeights :271-290
    a = "adaptive_weights.weight_network"
    w = torch.softmax(_lin(sd, f"{a}.3", _ln(sd, f"{a}.2", F.gelu(_lin(sd, f"{a}.0", clap)))), dim=-1)
    # LevelToUNetRouter :325-369
    am = assign * w[:, None, :]
    am = am / (am.sum(dim=-1, keepdim=True) + 1e-8)
    r = am @ torch.softmax(sd["router.routing_matrix"], dim=1)
    routed = {}
    for i, lvl in enumerate(("early", "mid", "late")):
        routed[lvl] = tok10 * r[:, :, i:i + 1] * torch.sigmoid(sd[f"router.level_gates.{lvl}"])
    tok77 = projector77_forward(sd, "projector", tok10)
    return dict(tokens_77=tok77, tokens_10=tok10, assignments=assign, hierarchy_weights=w, routed=routed)


# =======================================================================================
# Legacy HierarchicalAudioV4 (hierarchical_audio_v4.py:776-932) -- what scripts/inference.py:56 builds
# =======================================================================================
def legacy_hier_spec(audio_dim=512, text_dim=768, nf=5, nb=3, na=2) -> List[P]:
    d = "decomposer"
    s = linear(f"{d}.foreground_proj.0", audio_dim, 2 * text_dim) + linear(f"{d}.foreground_proj.3", 2 * text_dim, text_dim * nf)
    s += linear(f"{d}.background_proj.0", audio_dim, text_dim) + linear(f"{d}.background_proj.3", text_dim, text_dim * nb)
    s += linear(f"{d}.ambience_proj.0", audio_dim, text_dim // 2) + linear(f"{d}.ambience_proj.3", text_dim // 2, text_dim * na)
    s += [P(f"{d}.hierarchy_weights", (3,), "scalar")]
    s += norm(f"{d}.layer_norm", text_dim)
    s += cross_hierarchy_spec(f"{d}.cross_hierarchy_attn", text_dim, 192, 384)
    s += projector77_spec("projector", text_dim, text_dim)
    return s


def legacy_hier_forward(sd, clap: T) -> Dict[str, T]:
    d = "decomposer"
    B = clap.shape[0]
    D = sd[f"{d}.layer_norm.weight"].shape[0]
    w = torch.softmax(sd[f"{d}.hierarchy_weights"], dim=0)                # :858
    parts = []
    for i, nm in enumerate(("foreground", "background", "ambience")):
        z = _lin(sd, f"{d}.{nm}_proj.3", F.gelu(_lin(sd, f"{d}.{nm}_proj.0", clap)))   # :844-846
        parts.append(z.view(B, -1, D) * w[i])                              # :849-861
    cat = torch.cat(parts, dim=1)                                          # :864
    tok10 = _ln(sd, f"{d}.layer_norm", cross_hierarchy_forward(sd, f"{d}.cross_hierarchy_attn", cat))
    tok77 = projector77_forward(sd, "projector", tok10)
    return dict(tokens_77=tok77, tokens10=tok10, foreground=parts[0], background=parts[1],
                ambience=parts[2], weights=w, combined=tok10)


# =======================================================================================
# AudioAttnProcessor (models/audio_attention_processor.py:13-145)
# =======================================================================================
def attn_processor_spec(audio_dim=768, hidden=768, bott=64) -> List[P]:
    return ([P("alpha", (1,), "scalar")] + linear("audio_proj.0", audio_dim, bott)
            + linear("audio_proj.3", bott, hidden))


def attn_site_spec(c: int, cross_dim: int = 768) -> List[P]:
    """Weights of one diffusers cross-``Attention`` module (bias-free q/k/v, biased out)."""
    return (linear("to_q", c, c, False) + linear("to_k", cross_dim, c, False)
            + linear("to_v", cross_dim, c, False) + linear("to_out.0", c, c))


def pool_tokens(x: T, out_len: int) -> T:
    """F.adaptive_avg_pool1d over the token axis of [B,K,D] (:103-108)."""
    return F.adaptive_avg_pool1d(x.transpose(1, 2), out_len).transpose(1, 2)


def processor_modify_context(psd, ehs: T, audio_tokens: Optional[T], mode: str = "add") -> T:
    """Audio injection into the text states (:85-109).  Returns ehs' ([B,77,768] add / [B,81,768] concat)."""
    if audio_tokens is None:
        return ehs
    ap = _lin(psd, "audio_proj.3", F.gelu(_lin(psd, "audio_proj.0", audio_tokens)))   # :88 (eval: no dropout)
    if mode == "add":
        return ehs + torch.sigmoid(psd["alpha"]) * ap.mean(dim=1, keepdim=True)      # :92-97
    if mode == "concat":
        if ap.shape[1] > 4:
            ap = pool_tokens(ap, 4)                                                  # :101-108
        return torch.cat([ehs, ap], dim=1)                                           # :109
    return ehs


def processor_call(psd, attn: Dict[str, T], heads: int, h: T, ehs: Optional[T],
                   audio_tokens: Optional[T], mode: str = "add", scale: float = 1.0) -> T:
    """Whole AudioAttnProcessor.__call__ for SD-1.5 attention flags (no spatial_norm / norm_cross /
    residual / rescale).  ``attn`` holds to_q/to_k/to_v/to_out.0 weights."""
    if ehs is not None:
        ehs = processor_modify_context(psd, ehs, audio_tokens, mode)
    q = F.linear(h, attn["to_q.weight"]) * scale                                      # :115
    src = q if ehs is None else ehs                                                   # :117-118 quirk
    k = F.linear(src, attn["to_k.weight"])
    v = F.linear(src, attn["to_v.weight"])
    o = _heads_attention(q, k, v, heads)                                              # :124-131
    return F.linear(o, attn["to_out.0.weight"], attn["to_out.0.bias"])                # :134


def processor_call_decoupled(psd, attn: Dict[str, T], heads: int, h: T, ehs: T, audio_tokens: T) -> T:
    """DESIGN EXTENSION -- no counterpart in the reference (parity unpinned; pinned only by this definition).
    mode="decoupled" of the product's AudioAttnProcessor (BASELINE.json north_star: "fused decoupled text+audio
    cross-attention ... audio-branch scale and add"): the text keys stay untouched, the projected audio tokens
    (the reference's audio_proj, :88) get their OWN softmax through the site's to_k / to_v, and the audio branch is
    scaled by the reference's gate sigmoid(alpha) (:92) and added before to_out:
        o = softmax(q K_t^T s) V_t + sigmoid(alpha) * softmax(q K_a^T s) V_a."""
    ap = _lin(psd, "audio_proj.3", F.gelu(_lin(psd, "audio_proj.0", audio_tokens)))
    q = F.linear(h, attn["to_q.weight"])
    o = _heads_attention(q, F.linear(ehs, attn["to_k.weight"]), F.linear(ehs, attn["to_v.weight"]), heads)
    o2 = _heads_attention(q, F.linear(ap, attn["to_k.weight"]), F.linear(ap, attn["to_v.weight"]), heads)
    return F.linear(o + torch.sigmoid(psd["alpha"]) * o2, attn["to_out.0.weight"], attn["to_out.0.bias"])


LEVEL_OF_SITE_RULES = (("mid_block", "mid"), ("down_blocks.0", "early"), ("down_blocks.1", "early"),
                       ("down_blocks.2", "late"), ("down_blocks.3", "late"), ("up_blocks.0", "late"),
                       ("up_blocks.1", "late"), ("up_blocks.2", "mid"), ("up_blocks.3", "mid"))


def level_of_site(name: str) -> str:
    """AudioProcessorManager._create_level_mapping (:158-193) for one attn2 site name."""
    for key, lvl in LEVEL_OF_SITE_RULES:
        if key in name:
            return lvl
    return "mid"


# =======================================================================================
# AudioCrossAttention -- gated audio branch (models/audio_adapter_v4.py:168-261)
# =======================================================================================
def gated_xattn_spec(query_dim: int, context_dim=768, heads=8, dim_head=64) -> List[P]:
    inner = heads * dim_head
    return ([P("gate", (), "scalar", shift=-1.0)] + norm("norm", query_dim)
            + linear("to_q", query_dim, inner, False) + linear("to_k", context_dim, inner, False)
            + linear("to_v", context_dim, inner, False) + linear("to_out.0", inner, query_dim))


def gated_xattn_forward(sd, h: T, audio: T, heads: int = 8, mask: Optional[T] = None) -> T:
    hn = _ln(sd, "norm", h)
    q, k, v = _lin(sd, "to_q", hn), _lin(sd, "to_k", audio), _lin(sd, "to_v", audio)
    B, N, C = q.shape
    d = C // heads
    qh = q.view(B, N, heads, d).transpose(1, 2)
    kh = k.view(B, -1, heads, d).transpose(1, 2)
    vh = v.view(B, -1, heads, d).transpose(1, 2)
    dots = qh @ kh.transpose(-1, -2) * d ** -0.5
    if mask is not None:
        dots = dots.masked_fill(~mask, -torch.finfo(dots.dtype).max)                 # :244-245
    o = (torch.softmax(dots, dim=-1) @ vh).transpose(1, 2).reshape(B, N, C)
    return h + torch.sigmoid(sd["gate"]) * _lin(sd, "to_out.0", o)                   # :258-259
