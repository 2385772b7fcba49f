"""Drop-in modules on the GPU through the C ABI vs the golden vectors generated from the UNMODIFIED
reference modules (-m gpu).  fp32: <= 1e-5 rel-L2; bf16 activations: <= 2e-2."""
import numpy as np
import pytest
import torch

from oracle import audio as A
from oracle.pipeline import np_randn, rel_l2, to_torch
from oracle.weights import synth_state_dict

from clap2diffusion_b200 import ops
from clap2diffusion_b200.models import audio_adapter_v4 as padapter
from clap2diffusion_b200.models import audio_attention_processor as pproc
from clap2diffusion_b200.models import hierarchical_audio_v4 as phier

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _load(mod, sd):
    mod.load_state_dict(to_torch(sd))
    return mod.to(DEV).eval()


def test_audio_adapter(gold):
    g = gold("audio_adapter.npz")
    m = _load(padapter.AudioAdapter(), synth_state_dict(A.audio_adapter_spec(), int(g["seed"])))
    with torch.no_grad():
        out = m(_t(g["clap"]).to(DEV))
        assert rel_l2(out, _t(g["tokens"])) < 1e-5
        assert rel_l2(ops.norm_scale(out, 60.0, False), _t(g["tokens_norm60"])) < 1e-5
        # batch-256 (config 4 size) agrees with batch-3 rows
        big = m(_t(g["clap"]).to(DEV).repeat(86, 1)[:256])
        assert rel_l2(big[:3], _t(g["tokens"])) < 1e-5


def test_improved_hier(gold):
    g = gold("improved_hier.npz")
    sd = synth_state_dict(A.improved_hier_spec(), int(g["seed"]))
    for k, v in A.IMPROVED_BUFFERS.items():
        sd[k] = np.asarray(v, dtype=np.float32)
    m = _load(phier.ImprovedHierarchicalAudioEncoder(), sd)
    with torch.no_grad():
        t77, info = m(_t(g["clap"]).to(DEV), return_all=True)
    assert rel_l2(t77, _t(g["tokens_77"])) < 1e-5
    for k in ("tokens_10", "assignments", "hierarchy_weights"):
        assert rel_l2(info[k], _t(g[k])) < 1e-5, k
    for lvl in ("early", "mid", "late"):
        assert rel_l2(info["routed"][lvl], _t(g[f"routed_{lvl}"])) < 1e-5
    m.decomposer.set_temperature(0.5)
    with torch.no_grad():
        enc = m.encode(_t(g["clap"]).to(DEV))
    assert rel_l2(enc["assignments"], _t(g["assignments_T05"])) < 1e-5


def _kernels(fn):
    ops.PROFILE = []
    out = fn()
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    names = {}
    for r in rec:
        names[r[0]] = names.get(r[0], 0) + 1
    return out, names


def test_audio_side_bf16_on_tensor_cores(gold):
    """SURVEY 2.1 K8 / north star "projector MLPs": in bf16 the projector and decomposer GEMMs of
    ImprovedHierarchicalAudioEncoder and AudioAdapter run on the tcgen05 kernel (weights streamed once as bf16) and stay
    within 2e-2 of the goldens produced by the unmodified reference modules; at the benchmark batch (8) and at the
    config-4 batch (256) a sample's tokens do not depend on its batch neighbours."""
    g = gold("improved_hier.npz")
    sd = synth_state_dict(A.improved_hier_spec(), int(g["seed"]))
    for k, v in A.IMPROVED_BUFFERS.items():
        sd[k] = np.asarray(v, dtype=np.float32)
    m = _load(phier.ImprovedHierarchicalAudioEncoder(), sd)
    clap = _t(g["clap"]).to(DEV)
    rep = clap.repeat(86, 1)[:256].contiguous()
    with torch.no_grad():
        enc, ks = _kernels(lambda: m.encode(rep[:8].to(torch.bfloat16)))
        enc256 = m.encode(rep.to(torch.bfloat16))
    assert ks.get("linear_tc", 0) >= 20 and ks.get("attn_tc", 0) >= 4, ks      # the dense contractions are on tcgen05
    for k in ("tokens_10", "assignments", "hierarchy_weights", "tokens_77"):
        assert enc[k].shape[0] == 8 and rel_l2(enc[k][:3].float(), _t(g[k])) < 2e-2, k
        assert rel_l2(enc256[k][:3].float(), _t(g[k])) < 2e-2, k
    for lvl in ("early", "mid", "late"):
        assert rel_l2(enc["routed"][lvl][:3].float(), _t(g[f"routed_{lvl}"])) < 2e-2
        assert enc["routed"][lvl].dtype == torch.bfloat16
    ga = gold("audio_adapter.npz")
    ad = _load(padapter.AudioAdapter(), synth_state_dict(A.audio_adapter_spec(), int(ga["seed"])))
    ca = _t(ga["clap"]).to(DEV).repeat(86, 1)[:256].contiguous()
    with torch.no_grad():
        out, ks = _kernels(lambda: ad(ca[:8].to(torch.bfloat16)))
        out256 = ad(ca.to(torch.bfloat16))
    assert ks.get("linear_tc", 0) >= 8, ks           # 512 -> 256 -> 24576 token generator, the four self-attention blocks
    assert rel_l2(out[:3].float(), _t(ga["tokens"])) < 2e-2 and rel_l2(out256[:3].float(), _t(ga["tokens"])) < 2e-2


def test_legacy_hier(gold):
    g = gold("legacy_hier.npz")
    m = _load(phier.HierarchicalAudioV4(), synth_state_dict(A.legacy_hier_spec(), int(g["seed"])))
    with torch.no_grad():
        t77, hz = m(_t(g["clap"]).to(DEV), return_intermediate=True)
    assert rel_l2(t77, _t(g["tokens_77"])) < 1e-5
    for k in ("tokens10", "foreground", "background", "ambience", "weights"):
        assert rel_l2(hz[k], _t(g[k])) < 1e-5, k


class _Site:
    spatial_norm = None
    norm_cross = None
    residual_connection = False
    rescale_output_factor = 1.0

    def __init__(self, sd, heads=8):
        lin = lambda w, b=None: type("L", (), {"weight": w, "bias": b})()
        self.to_q, self.to_k, self.to_v = lin(sd["to_q.weight"]), lin(sd["to_k.weight"]), lin(sd["to_v.weight"])
        self.to_out = [lin(sd["to_out.0.weight"], sd["to_out.0.bias"])]
        self.heads = heads
        self.scale = (sd["to_q.weight"].shape[0] // heads) ** -0.5


@pytest.mark.parametrize("N,C", [(4096, 320), (1024, 640), (256, 1280), (64, 1280)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_attn_processor_sites(gold, N, C, dtype, tol):
    g = gold("attn_processor.npz")
    seed = int(g["seed"])
    psd = to_torch(synth_state_dict(A.attn_processor_spec(), seed))
    ehs = _t(np_randn("ehs", (2, 77, 768))).to(DEV)
    audio = (_t(np_randn("audio10", (2, 10, 768))) * 0.3).to(DEV)
    asd = synth_state_dict(A.attn_site_spec(C), seed, prefix=f"site{C}.")
    site = _Site({k.split(".", 1)[1]: torch.from_numpy(v).to(DEV) for k, v in asd.items()})
    h = _t(np_randn(f"h_{N}_{C}", (2, N, C))).to(DEV).to(dtype)
    rows = torch.from_numpy(g[f"rows_{N}_{C}"]).to(DEV)
    for mode in ("add", "concat"):
        proc = pproc.AudioAttnProcessor(level="mid", mode=mode)
        proc.load_state_dict(psd)
        proc = proc.to(DEV).eval()
        with torch.no_grad():
            out = proc(site, h, encoder_hidden_states=ehs, audio={"mid": audio})
            out_na = proc(site, h, encoder_hidden_states=ehs)
        assert out.dtype == dtype and tuple(out.shape) == (2, N, C)
        assert rel_l2(out[:, rows].float(), _t(g[f"out_{mode}_{N}_{C}"])) < tol, mode
        assert rel_l2(out_na[:, rows].float(), _t(g[f"out_noaudio_{N}_{C}"])) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_attn_processor_attention_mask_and_kv_cache(gold, dtype, tol):
    """attention_mask pass-through (reference :129) against the UNMODIFIED reference processor's output (additive
    [B*heads,1,T] key-padding mask as diffusers prepares it), the boolean [B,T] form, loud refusal of general biases, and
    the identity-keyed K/V cache of the stand-alone call (a denoising loop projects K/V once per image)."""
    g = gold("attn_processor_mask.npz")
    seed, N, C, heads = int(g["seed"]), 256, 1280, 8
    psd = to_torch(synth_state_dict(A.attn_processor_spec(), seed))
    ehs = _t(np_randn("ehs", (2, 77, 768))).to(DEV)
    audio = (_t(np_randn("audio10", (2, 10, 768))) * 0.3).to(DEV)
    asd = synth_state_dict(A.attn_site_spec(C), seed, prefix=f"site{C}.")
    site = _Site({k.split(".", 1)[1]: torch.from_numpy(v).to(DEV) for k, v in asd.items()})
    h = _t(np_randn(f"h_{N}_{C}", (2, N, C))).to(DEV).to(dtype)
    keep = _t(g["keep"]).to(DEV)
    bias = torch.zeros(2, 77, device=DEV).masked_fill(~keep, -10000.0)
    mask = bias[:, None, None, :].expand(2, heads, 1, 77).reshape(2 * heads, 1, 77).contiguous()
    rows = torch.from_numpy(g["rows"]).to(DEV)
    proc = pproc.AudioAttnProcessor(level="mid", mode="add")
    proc.load_state_dict(psd)
    proc = proc.to(DEV).eval()
    with torch.no_grad():
        out = proc(site, h, encoder_hidden_states=ehs, attention_mask=mask, audio={"mid": audio})
        out_b = proc(site, h, encoder_hidden_states=ehs, attention_mask=keep, audio={"mid": audio})
        assert rel_l2(out[:, rows].float(), _t(g["out"])) < tol
        assert torch.equal(out, out_b)
        with pytest.raises(Exception):          # a per-query bias is not a key-padding mask: refused, never ignored
            proc(site, h, encoder_hidden_states=ehs, attention_mask=torch.randn(2 * heads, N, 77, device=DEV), audio={"mid": audio})
        with pytest.raises(Exception):
            proc(site, h, encoder_hidden_states=ehs, attention_mask=torch.zeros(2, 50, device=DEV), audio={"mid": audio})
        # K/V cache: same ehs / audio tensors -> one projection for many "steps"; an in-place change is a miss
        aud = {"mid": audio}
        n0 = proc.kv_projections
        o1 = proc(site, h, encoder_hidden_states=ehs, audio=aud)
        for _ in range(3):
            o2 = proc(site, h, encoder_hidden_states=ehs, audio=aud)
        assert proc.kv_projections == n0 + 1 and torch.equal(o1, o2)
        ehs.mul_(1.25)
        o3 = proc(site, h, encoder_hidden_states=ehs, audio=aud)
        assert proc.kv_projections == n0 + 2 and not torch.equal(o1, o3)
        proc.alpha.fill_(1.5)          # (no_grad) in-place parameter update bumps the version counter
        proc(site, h, encoder_hidden_states=ehs, audio=aud)
        assert proc.kv_projections == n0 + 3


def test_gated_xattn(gold):
    g = gold("gated_xattn.npz")
    m = _load(padapter.AudioCrossAttention(320), synth_state_dict(A.gated_xattn_spec(320), int(g["seed"])))
    h, a = _t(np_randn("gx_h", (2, 256, 320))).to(DEV), _t(np_randn("gx_a16", (2, 16, 768))).to(DEV)
    with torch.no_grad():
        assert rel_l2(m(h, a), _t(g["out"])) < 1e-5
        assert rel_l2(m(h, a, _t(g["mask"]).to(DEV)), _t(g["out_masked"])) < 1e-5


@pytest.mark.parametrize("N,C", [(4096, 320), (1024, 640), (256, 1280), (64, 1280)])
def test_attn_processor_decoupled_mode(N, C):
    """Design extension (no reference counterpart): separate text / audio softmaxes, audio branch scaled by
    sigmoid(alpha) and added, all inside the fused kernel -- vs the oracle's definition in fp32."""
    seed = 11
    psd = to_torch(synth_state_dict(A.attn_processor_spec(), seed))
    psd["alpha"] = torch.tensor([0.7])
    ehs = _t(np_randn("ehs", (2, 77, 768)))
    audio = _t(np_randn("audio10", (2, 10, 768))) * 0.3
    asd = to_torch({k.split(".", 1)[1]: v for k, v in synth_state_dict(A.attn_site_spec(C), seed, prefix=f"site{C}.").items()})
    h = _t(np_randn(f"h_{N}_{C}", (2, N, C)))
    ref = A.processor_call_decoupled(psd, asd, 8, h, ehs, audio)
    ref_na = A.processor_call(psd, asd, 8, h, ehs, None)
    site = _Site({k: v.to(DEV) for k, v in asd.items()})
    proc = pproc.AudioAttnProcessor(level="mid", mode="decoupled")
    proc.load_state_dict(psd)
    proc = proc.to(DEV).eval()
    n0 = ops._lib.launch_count()
    with torch.no_grad():
        out = proc(site, h.to(DEV).to(torch.bfloat16), encoder_hidden_states=ehs.to(DEV), audio={"mid": audio.to(DEV)})
        assert (ops.lib.c2d_last_kernel() or b"").decode() != ""
        out_na = proc(site, h.to(DEV).to(torch.bfloat16), encoder_hidden_states=ehs.to(DEV))
    assert ops._lib.launch_count() > n0
    assert rel_l2(out.float().cpu(), ref) < 2e-2
    assert rel_l2(out_na.float().cpu(), ref_na) < 2e-2
    assert rel_l2(ref, ref_na) > 5e-2           # the audio branch matters at this gate value
    with pytest.raises(Exception):              # fp32 parity mode has no decoupled kernel: refused, not re-routed
        proc(site, h.to(DEV), encoder_hidden_states=ehs.to(DEV), audio={"mid": audio.to(DEV)})
