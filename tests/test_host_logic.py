"""HOST logic of the product (weight packing, topology, skip wiring, hoisting, processor protocol, drop-in
module plumbing, sampler, schedulers) checked on CPU against the oracle and the reference-generated golden
vectors, with the libc2d-backed ops swapped for the torch test double (tests/torch_ops.py).  The CUDA
kernels themselves are checked by the -m gpu tests."""
import os

import numpy as np
import pytest
import torch

import torch_ops
from oracle import audio as A
from oracle import pipeline as PL
from oracle import sd15
from oracle.pipeline import np_randn, rel_l2, to_torch
from oracle.weights import synth_state_dict

from clap2diffusion_b200 import schedulers
from clap2diffusion_b200.models import audio_adapter_v4 as padapter
from clap2diffusion_b200.models import audio_attention_processor as pproc
from clap2diffusion_b200.models import hierarchical_audio_v4 as phier


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def W():
    return PL.build_weights(seed=0, with_vae=True)


def test_state_dict_contract():
    """Same keys and shapes as the reference modules, via the oracle specs."""
    for mod, spec, extra in ((padapter.AudioAdapter(), A.audio_adapter_spec(), {}),
                             (phier.ImprovedHierarchicalAudioEncoder(), A.improved_hier_spec(), A.IMPROVED_BUFFERS),
                             (phier.HierarchicalAudioV4(), A.legacy_hier_spec(), {}),
                             (pproc.AudioAttnProcessor("mid"), A.attn_processor_spec(), {}),
                             (padapter.AudioCrossAttention(320), A.gated_xattn_spec(320), {})):
        sd = mod.state_dict()
        want = {p.name: tuple(p.shape) for p in spec}
        want.update({k: tuple(np.asarray(v).shape) for k, v in extra.items()})
        assert {k: tuple(v.shape) for k, v in sd.items()} == want, type(mod).__name__


def test_default_init_statistics():
    """KAT (iv): tokens out of the default-init adapter have norm ~ sqrt(768); gates at their reference init."""
    m = padapter.AudioCrossAttention(320)
    assert abs(float(m.gate) + 5.0) < 1e-6
    r = phier.LevelToUNetRouter()
    assert torch.allclose(r.routing_matrix.detach(), torch.tensor([[.1, .3, .6], [.2, .6, .2], [.6, .3, .1]]))
    assert float(pproc.AudioAttnProcessor("early").alpha) == 0.0


def test_temperature_scheduler(gold):
    g = gold("temperature_kat.npz")
    dec = phier.SoftHierarchicalDecomposition()
    sch = phier.TemperatureScheduler(dec, T_max=2.0, T_min=0.5, total_steps=2000)
    for step, temp in zip(g["steps"].tolist(), g["temps"].tolist()):
        sch.step(step)
        assert abs(float(dec.temperature) - temp) < 1e-6
    with pytest.raises(ValueError):
        phier.TemperatureScheduler(dec, mode="bogus").step(300)
    with pytest.raises(ValueError):
        phier.CrossHierarchyAttention(768, num_heads=4, bottleneck_dim=190)


def test_audio_adapter_module(gold):
    g = gold("audio_adapter.npz")
    m = padapter.AudioAdapter().eval()
    m.load_state_dict(to_torch(synth_state_dict(A.audio_adapter_spec(), int(g["seed"]))))
    with torch_ops.installed(), torch.no_grad():
        out = m(_t(g["clap"]))
        n60 = torch_ops.norm_scale(out, 60.0, False)
    assert rel_l2(out, _t(g["tokens"])) < 5e-6
    assert rel_l2(n60, _t(g["tokens_norm60"])) < 5e-6


def test_improved_hier_module(gold):
    g = gold("improved_hier.npz")
    sd = synth_state_dict(A.improved_hier_spec(), int(g["seed"]))
    for k, v in A.IMPROVED_BUFFERS.items():
        sd[k] = np.asarray(v, dtype=np.float32)
    m = phier.ImprovedHierarchicalAudioEncoder().eval()
    m.load_state_dict(to_torch(sd))
    with torch_ops.installed(), torch.no_grad():
        t77, info = m(_t(g["clap"]), return_all=True)
        t77b = m(_t(g["clap"]))
    assert rel_l2(t77, _t(g["tokens_77"])) < 5e-6 and rel_l2(t77b, _t(g["tokens_77"])) < 5e-6
    assert rel_l2(info["tokens_10"], _t(g["tokens_10"])) < 5e-6
    assert rel_l2(info["assignments"], _t(g["assignments"])) < 5e-6
    assert rel_l2(info["hierarchy_weights"], _t(g["hierarchy_weights"])) < 5e-6
    for lvl in ("early", "mid", "late"):
        assert rel_l2(info["routed"][lvl], _t(g[f"routed_{lvl}"])) < 5e-6
    assert set(info["losses"]) == {"entropy", "orthogonality", "prior"}
    assert set(info["stats"]) == {"avg_assignment", "entropy", "effective_levels"}
    assert info["temperature"] == 2.0


def test_legacy_hier_module(gold):
    g = gold("legacy_hier.npz")
    m = phier.HierarchicalAudioV4().eval()
    m.load_state_dict(to_torch(synth_state_dict(A.legacy_hier_spec(), int(g["seed"]))))
    with torch_ops.installed(), torch.no_grad():
        t77, hz = m(_t(g["clap"]), return_intermediate=True)
    assert rel_l2(t77, _t(g["tokens_77"])) < 5e-6
    for k in ("tokens10", "foreground", "background", "ambience", "weights"):
        assert rel_l2(hz[k], _t(g[k])) < 5e-6, k
    assert tuple(hz["foreground"].shape) == (3, 5, 768) and tuple(hz["ambience"].shape) == (3, 2, 768)


class _Site:
    """Duck-typed diffusers Attention (what the processor reads)."""
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


def test_attn_processor_module(gold):
    g = gold("attn_processor.npz")
    seed = int(g["seed"])
    psd = to_torch(synth_state_dict(A.attn_processor_spec(), seed))
    ehs = _t(np_randn("ehs", (2, 77, 768)))
    audio = _t(np_randn("audio10", (2, 10, 768))) * 0.3
    for (N, C) in ((256, 1280), (64, 1280)):
        asd = synth_state_dict(A.attn_site_spec(C), seed, prefix=f"site{C}.")
        site = _Site(to_torch({k.split(".", 1)[1]: v for k, v in asd.items()}))
        h = _t(np_randn(f"h_{N}_{C}", (2, N, C)))
        rows = g[f"rows_{N}_{C}"]
        for mode in ("add", "concat"):
            proc = pproc.AudioAttnProcessor(level="mid", mode=mode).eval()
            proc.load_state_dict(psd)
            with torch_ops.installed(), torch.no_grad():
                out = proc(site, h, encoder_hidden_states=ehs, audio={"mid": audio})
                out_na = proc(site, h, encoder_hidden_states=ehs)                 # fall-through: no audio
                out_lv = proc(site, h, encoder_hidden_states=ehs, audio={"late": audio})   # other level only
            assert rel_l2(out[:, rows], _t(g[f"out_{mode}_{N}_{C}"])) < 5e-6
            assert rel_l2(out_na[:, rows], _t(g[f"out_noaudio_{N}_{C}"])) < 5e-6
            assert rel_l2(out_lv, out_na) == 0.0
        if N == 64:   # 4-D input path
            proc = pproc.AudioAttnProcessor(level="mid", mode="add").eval()
            proc.load_state_dict(psd)
            h4 = h.transpose(1, 2).reshape(2, C, 8, 8).contiguous()
            with torch_ops.installed(), torch.no_grad():
                o4 = proc(site, h4, encoder_hidden_states=ehs, audio={"mid": audio})
            assert tuple(o4.shape) == (2, C, 8, 8)
            assert rel_l2(o4.reshape(2, C, 64).transpose(1, 2)[:, rows], _t(g[f"out_add_{N}_{C}"])) < 5e-6


def test_attn_processor_mask_and_kv_cache_host_logic(gold):
    """attention_mask reduction (additive [B*heads,1,T] / boolean [B,T] key-padding masks; general biases refused) against
    the unmodified reference's output, and the identity-keyed K/V cache of the stand-alone call."""
    g = gold("attn_processor_mask.npz")
    seed, N, C, heads = int(g["seed"]), 256, 1280, 8
    psd = to_torch(synth_state_dict(A.attn_processor_spec(), seed))
    ehs = _t(np_randn("ehs", (2, 77, 768)))
    audio = _t(np_randn("audio10", (2, 10, 768))) * 0.3
    asd = synth_state_dict(A.attn_site_spec(C), seed, prefix=f"site{C}.")
    site = _Site(to_torch({k.split(".", 1)[1]: v for k, v in asd.items()}))
    h = _t(np_randn(f"h_{N}_{C}", (2, N, C)))
    keep = _t(g["keep"])
    bias = torch.zeros(2, 77).masked_fill(~keep, -10000.0)
    mask = bias[:, None, None, :].expand(2, heads, 1, 77).reshape(2 * heads, 1, 77).contiguous()
    proc = pproc.AudioAttnProcessor(level="mid", mode="add").eval()
    proc.load_state_dict(psd)
    aud = {"mid": audio}
    with torch_ops.installed(), torch.no_grad():
        out = proc(site, h, encoder_hidden_states=ehs, attention_mask=mask, audio=aud)
        out_b = proc(site, h, encoder_hidden_states=ehs, attention_mask=keep, audio=aud)
        assert rel_l2(out[:, g["rows"]], _t(g["out"])) < 5e-6
        assert torch.equal(out, out_b)
        for bad in (torch.randn(2 * heads, N, 77), torch.zeros(2, 50), bias[:, None, :] * 0.5 - 1.0,
                    torch.full((2, 77), -10000.0)):
            with pytest.raises(Exception):
                proc(site, h, encoder_hidden_states=ehs, attention_mask=bad, audio=aud)
        n0 = proc.kv_projections
        o1 = proc(site, h, encoder_hidden_states=ehs, audio=aud)
        o2 = proc(site, h, encoder_hidden_states=ehs, audio=aud)
        assert proc.kv_projections == n0 + 1 and torch.equal(o1, o2)
        ehs.mul_(1.25)                                   # in-place change of the text states: version bump -> miss
        o3 = proc(site, h, encoder_hidden_states=ehs, audio=aud)
        assert proc.kv_projections == n0 + 2 and not torch.equal(o1, o3)
        proc.alpha.fill_(1.5)                            # parameter update -> miss
        proc(site, h, encoder_hidden_states=ehs, audio=aud)
        assert proc.kv_projections == n0 + 3
        proc(site, h, encoder_hidden_states=ehs, audio={"mid": audio.clone()})      # another audio tensor -> miss
        assert proc.kv_projections == n0 + 4


def test_attn_processor_decoupled_mode():
    """Design extension (no reference counterpart): text and audio branches with separate softmaxes, audio branch
    scaled by sigmoid(alpha) and added -- host logic vs the oracle's definition."""
    seed = 11
    psd = to_torch(synth_state_dict(A.attn_processor_spec(), seed))
    psd["alpha"] = torch.tensor([0.7])
    ehs = _t(np_randn("ehs", (2, 77, 768)))
    audio = _t(np_randn("audio10", (2, 10, 768))) * 0.3
    C, N = 320, 128
    asd = to_torch({k.split(".", 1)[1]: v for k, v in synth_state_dict(A.attn_site_spec(C), seed, prefix=f"site{C}.").items()})
    site = _Site(asd)
    h = _t(np_randn(f"h_{N}_{C}", (2, N, C)))
    proc = pproc.AudioAttnProcessor(level="mid", mode="decoupled").eval()
    proc.load_state_dict(psd)
    with torch_ops.installed(), torch.no_grad():
        out = proc(site, h, encoder_hidden_states=ehs, audio={"mid": audio})
        out_na = proc(site, h, encoder_hidden_states=ehs)                          # no audio: plain text cross-attention
    assert rel_l2(out, A.processor_call_decoupled(psd, asd, 8, h, ehs, audio)) < 5e-6
    assert rel_l2(out_na, A.processor_call(psd, asd, 8, h, ehs, None)) < 5e-6
    assert rel_l2(out, out_na) > 1e-3


def test_gated_xattn_module(gold):
    g = gold("gated_xattn.npz")
    m = padapter.AudioCrossAttention(320).eval()
    m.load_state_dict(to_torch(synth_state_dict(A.gated_xattn_spec(320), int(g["seed"]))))
    h, a = _t(np_randn("gx_h", (2, 256, 320))), _t(np_randn("gx_a16", (2, 16, 768)))
    with torch_ops.installed(), torch.no_grad():
        assert rel_l2(m(h, a), _t(g["out"])) < 5e-6
        assert rel_l2(m(h, a, _t(g["mask"])), _t(g["out_masked"])) < 5e-6


def test_schedulers_match_oracle():
    for n in (20, 50):
        plan = schedulers.ddim_plan(n)
        ref = sd15.ddim_coeffs(n)
        assert plan.timesteps == [float(t) for t, _, _ in ref]
        assert np.allclose(plan.coef[:, 0], [a for _, a, _ in ref], rtol=1e-6)
        assert np.allclose(plan.coef[:, 1], [b for _, _, b in ref], rtol=1e-6)
    ts, sig = sd15.euler_sigmas(50)
    plan = schedulers.euler_plan(50)
    assert plan.timesteps == ts and abs(plan.init_scale - sig[0]) < 1e-6
    assert np.allclose(plan.coef[:, 1], np.diff(np.asarray(sig)), rtol=1e-6)


def _build_unet(W, dtype=torch.float32, fused=False):
    from clap2diffusion_b200 import unet as unet_mod
    from clap2diffusion_b200.unet import SD15UNet
    if fused:
        # the bf16 product wiring (fused GroupNorm statistics, folded LayerNorms, fused GEGLU) in fp32 arithmetic
        class _Fused(SD15UNet):
            def __init__(self, *a, **k):
                self._force = True
                super().__init__(*a, **k)

            def __setattr__(self, k, v):
                if k in ("fused_gn", "fold_ln", "fuse_geglu", "fused_xattn") and getattr(self, "_force", False):
                    v = True
                object.__setattr__(self, k, v)
        unet = _Fused(W["unet"], device="cpu", dtype=dtype)
    else:
        unet = SD15UNet(W["unet"], device="cpu", dtype=dtype)
    mgr = pproc.AudioProcessorManager(unet)
    mgr.setup_processors(mode="add")
    assert {k: len(v) for k, v in mgr.level_mapping.items()} == {"early": 4, "mid": 7, "late": 5}
    for lvl, names in mgr.level_mapping.items():
        unet.sites[names[0][:-len(".processor")]].processor.load_state_dict(W[f"proc_{lvl}"])
    return unet, mgr


def test_unet_host_logic_vs_oracle(gold, W):
    g = gold("unet_16x16.npz")
    clap = _t(PL.clap_embedding(0))[None]
    with torch_ops.installed(), torch.no_grad():
        unet, mgr = _build_unet(W)
        assert len(unet.attn_processors) == 32
        assert sorted(unet.attn_processors) == sorted(sd15.attn_processor_names())
        enc = phier.ImprovedHierarchicalAudioEncoder().eval()
        enc.load_state_dict(W["hier"])
        routed = enc.encode(clap, with_tokens77=False)["routed"]
        x = _t(PL.init_noise(5, 16, 16))[None]
        taps = {}
        eps = unet(x, float(g["t"]), _t(PL.text_states("a beach"))[None],
                   cross_attention_kwargs=mgr.get_audio_kwargs(routed), taps=taps)
    assert rel_l2(taps["conv_in"].permute(0, 3, 1, 2), _t(g["conv_in"])) < 1e-5
    assert rel_l2(taps["mid"].permute(0, 3, 1, 2), _t(g["mid"])) < 1e-5
    assert rel_l2(eps, _t(g["eps"])) < 2e-5


def test_unet_fused_groupnorm_host_logic_vs_oracle(gold, W):
    """The bf16 product path's wiring (producer-side channel statistics -> one-pass GroupNorm, K-concatenated
    shortcut GEMM) run in fp32 through the test double must reproduce the oracle UNet."""
    g = gold("unet_16x16.npz")
    clap = _t(PL.clap_embedding(0))[None]
    with torch_ops.installed(), torch.no_grad():
        unet, mgr = _build_unet(W, fused=True)
        enc = phier.ImprovedHierarchicalAudioEncoder().eval()
        enc.load_state_dict(W["hier"])
        routed = enc.encode(clap, with_tokens77=False)["routed"]
        x = _t(PL.init_noise(5, 16, 16))[None]

        def _no_norm(*a, **k):
            raise AssertionError("the fused UNet path must not launch stand-alone LayerNorm / two-pass GroupNorm kernels")
        from clap2diffusion_b200 import ops as _ops
        _ops.layer_norm = _no_norm          # torch_ops.installed() restores both on exit
        _ops.group_norm = _no_norm
        fused_calls = []
        _xattn = _ops.xattn
        _ops.xattn = lambda *a, **k: (fused_calls.append(1), _xattn(*a, **k))[1]
        kv = unet.prepare_conditioning(_t(PL.text_states("a beach"))[None], mgr.get_audio_kwargs(routed))
        assert all(isinstance(v, _ops.XattnKV) for v in kv.values()) and len(kv) == 16      # packed K/V cache per attn2 site
        eps = unet(x, float(g["t"]), _t(PL.text_states("a beach"))[None],
                   cross_attention_kwargs=mgr.get_audio_kwargs(routed))
    assert len(fused_calls) == 16           # every attn2 site went through the fused to_q + attention entry point
    assert rel_l2(eps, _t(g["eps"])) < 1e-4


def test_sampler_host_logic_vs_oracle(W):
    """3 DDIM + 3 Euler steps with CFG on an 8x8 latent: product loop (hoisted tables, fused CFG/scheduler
    contract, D2 audio tiling) == oracle loop."""
    from clap2diffusion_b200.sampler import Sampler
    clap = _t(np.stack([PL.clap_embedding(1), PL.clap_embedding(2)]))
    cc = _t(np.stack([PL.text_states("a beach"), PL.text_states("a city")]))
    cu = _t(np.stack([PL.text_states("")] * 2))
    noise = _t(np.stack([PL.init_noise(1, 8, 8), PL.init_noise(2, 8, 8)]))
    for sched in ("ddim", "euler"):
        ref = PL.sample(W, clap, cc, cu, noise, steps=50, guidance=7.5, scheduler=sched, max_steps=3)
        with torch_ops.installed(), torch.no_grad():
            unet, _ = _build_unet(W)
            enc = phier.ImprovedHierarchicalAudioEncoder().eval()
            enc.load_state_dict(W["hier"])
            out = Sampler(unet, enc, None, use_graph=False).sample(clap, cc, cu, noise, steps=50, guidance=7.5,
                                                                   scheduler=sched, decode=False, trace=True, max_steps=3)
        for i in range(3):
            assert rel_l2(out["trace"][i], ref["latents"][i]) < 5e-5, (sched, i)


def test_vae_host_logic_vs_oracle(W):
    from clap2diffusion_b200.vae import VAEDecoder
    z = _t(PL.init_noise(3, 8, 8))[None] * 0.5
    with torch.no_grad():
        ref = sd15.vae_decode(W["vae"], z)
    with torch_ops.installed(), torch.no_grad():
        img = VAEDecoder(W["vae"], device="cpu", dtype=torch.float32).decode(z)
    assert tuple(img.shape) == (1, 3, 64, 64)
    assert rel_l2(img, ref) < 2e-5


def test_vae_legacy_attention_key_names(W):
    """The published SD-1.5 VAE files use the pre-0.19 diffusers attention names (query / key / value / proj_attn, stored
    as 1x1 convolutions): accepted as aliases; a state dict with neither raises a KeyError that lists the block's keys."""
    from clap2diffusion_b200.vae import VAEDecoder
    a = "decoder.mid_block.attentions.0"
    legacy = {}
    for k, v in W["vae"].items():
        for new, old in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
            if k.startswith(f"{a}.{new}."):
                k = k.replace(f"{a}.{new}.", f"{a}.{old}.")
                v = v[:, :, None, None] if k.endswith("weight") else v
        legacy[k] = v
    assert f"{a}.query.weight" in legacy and f"{a}.to_q.weight" not in legacy and legacy[f"{a}.proj_attn.weight"].dim() == 4
    z = _t(PL.init_noise(3, 8, 8))[None] * 0.5
    with torch_ops.installed(), torch.no_grad():
        ref = VAEDecoder(W["vae"], device="cpu", dtype=torch.float32).decode(z)
        img = VAEDecoder(legacy, device="cpu", dtype=torch.float32).decode(z)
        assert torch.equal(img, ref)
        broken = {k: v for k, v in legacy.items() if not k.startswith(f"{a}.key.")}
        with pytest.raises(KeyError, match="to_k / key"):
            VAEDecoder(broken, device="cpu", dtype=torch.float32)


def test_vae_encoder_host_logic_vs_oracle():
    """AutoencoderKL encoder: product launch sequence (Downsample2D as conv3x3_down, folded latent
    scaling) == oracle restatement, and the state-dict contract (34,163,592 + 72 parameters)."""
    from clap2diffusion_b200.vae import VAEEncoder, encoder_param_shapes
    from oracle.weights import synth_state_dict
    spec = sd15.vae_encoder_spec()
    assert sum(int(np.prod(p.shape)) for p in spec) == 34_163_592 + 72
    shapes = encoder_param_shapes()
    assert {p.name: tuple(p.shape) for p in spec} == shapes
    sd = to_torch(synth_state_dict(spec, 3))
    img = torch.tanh(_t(np_randn("vae_img", (2, 3, 64, 64))))
    with torch.no_grad():
        mean, logvar = sd15.vae_encode(sd, img)
    with torch_ops.installed(), torch.no_grad():
        enc = VAEEncoder(sd, device="cpu", dtype=torch.float32)
        m2, lv2 = enc.moments(img)
        z = enc.encode(img)
    assert tuple(z.shape) == (2, 4, 8, 8)
    assert rel_l2(m2, mean) < 2e-5 and rel_l2(lv2, logvar) < 2e-5
    assert rel_l2(z, mean * sd15.VAE_SCALING) < 2e-5


def test_inference_cli_surface():
    """Same flags / defaults and class API as the reference's scripts/inference.py (:21-214)."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(__file__), "..", "scripts", "inference.py")).read()
    for flag in ("--audio", "--text", "--output", "--checkpoint_dir", "--steps", "--cfg_scale", "--seed", "--no_hierarchical"):
        assert f'"{flag}"' in src, flag
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "AudioToImageInference"][0]
    methods = {n.name for n in cls.body if isinstance(n, ast.FunctionDef)}
    assert {"load_models", "load_audio", "extract_clap_embedding", "apply_normalization", "generate", "batch_generate"} <= methods


def test_clap_tower_host_logic_vs_oracle(gold):
    """ClapAudioTower's packing (fused QKV, expanded relative-position bias, folded BatchNorm, DFT / mel matrices) and
    launch sequence through the torch double == the oracle restatement == the HF goldens."""
    from clap2diffusion_b200.clap import ClapAudioTower, param_shapes
    from clap2diffusion_b200.synthetic import synthetic_audio
    from oracle import clap as C
    from oracle.weights import synth_state_dict
    g = gold("clap_audio.npz")
    spec = C.clap_audio_spec()
    assert {p.name: tuple(p.shape) for p in spec} == param_shapes()
    sd = {k: _t(v) for k, v in synth_state_dict(spec, 4321).items()}
    waves = np.stack([synthetic_audio(s) for s in (0, 1)])
    waves[1] *= np.linspace(0.05, 1.0, waves.shape[1], dtype=np.float32)
    with torch_ops.installed(), torch.no_grad():
        tower = ClapAudioTower(sd, device="cpu", dtype=torch.float32, clip_chunk=1)
        taps = {}
        emb = tower.encode(_t(waves), taps)
    # front end vs the HF feature extractor (undo the folded BatchNorm)
    mel = (taps["mel_bn"] - tower.w["bn_b"]) / tower.w["bn_a"]
    assert rel_l2(mel, _t(g["mel"][:, 0])) < 2e-5
    assert rel_l2(taps["patch_embed"][:, ::64], _t(g["patch_embed"])) < 1e-4
    assert rel_l2(taps["stage0"][:, ::64], _t(g["stage0"])) < 1e-4
    assert rel_l2(taps["stage2"][:, ::16], _t(g["stage2"])) < 1e-4
    assert rel_l2(taps["pooled"], _t(g["pooled"])) < 1e-4
    assert rel_l2(emb, _t(g["embedding"])) < 1e-4
    assert int(g["n_params"]) == 28190872


def test_clap_mel_filter_bank_and_feature_config(tmp_path):
    """Slaney filter bank for a configurable range against Hugging Face's audio_utils.mel_filter_bank, and the resolution
    order of the feature-extractor settings: caller > preprocessor_config.json > published laion/clap-htsat-* > class defaults."""
    import json
    from transformers.audio_utils import mel_filter_bank
    from clap2diffusion_b200 import clap as pclap
    from clap2diffusion_b200.models.audio_encoder import CLAPAudioEncoder
    for fmin, fmax in ((0.0, 14000.0), (50.0, 14000.0), (20.0, 12000.0)):
        ref = mel_filter_bank(num_frequency_bins=513, num_mel_filters=64, min_frequency=fmin, max_frequency=fmax,
                              sampling_rate=48000, norm="slaney", mel_scale="slaney")
        assert np.allclose(pclap._slaney_mel_filters(fmin, fmax), ref, rtol=1e-9, atol=1e-12), (fmin, fmax)
    assert CLAPAudioEncoder._feature_config("laion/clap-htsat-unfused", None)["frequency_min"] == 50.0
    assert CLAPAudioEncoder._feature_config("some/other-model", None)["frequency_min"] == 0.0
    with open(tmp_path / "preprocessor_config.json", "w") as f:
        json.dump({"frequency_min": 30, "frequency_max": 13000, "truncation": "rand_trunc", "padding": "repeatpad"}, f)
    cfg = CLAPAudioEncoder._feature_config(str(tmp_path), None)
    assert (cfg["frequency_min"], cfg["frequency_max"]) == (30.0, 13000.0) and cfg["source"].endswith("preprocessor_config.json")
    assert CLAPAudioEncoder._feature_config(str(tmp_path), {"frequency_min": 10})["frequency_min"] == 10.0


def test_evaluate_and_gradio_surface_without_gpu():
    """Class / method / CLI names of the reference's evaluate.py and gradio_app.py (no device work)."""
    import importlib.util
    import inspect
    import os
    root = os.path.join(os.path.dirname(__file__), "..")
    for name, path, cls, methods in (("ev", "scripts/evaluate.py", "Evaluator", ("compute_clip_score", "compute_audio_alignment", "evaluate_single",
                                                                                  "evaluate_dataset", "print_results")),
                                     ("gr", "app/gradio_app.py", "AudioToImageGenerator", ("load_models", "generate"))):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        c = getattr(mod, cls)
        for m in methods:
            assert callable(getattr(c, m))
        assert callable(mod.main)
    sig = inspect.signature(mod.AudioToImageGenerator.generate)
    assert list(sig.parameters)[1:] == ["audio_path", "text_prompt", "norm_value", "num_steps", "cfg_scale", "seed", "model_type"]


def test_train_stage3_cli_surface():
    """scripts/train_stage3.py keeps the reference's configuration keys and defaults (main :275-286, gradient accumulation
    aside: the global batch comes from the GPUs) and refuses to run without a CUDA device."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("train_stage3_cli", os.path.join(ROOT, "scripts", "train_stage3.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    assert cli.DEFAULTS == {"learning_rate": 1e-5, "weight_decay": 0.01, "num_steps": 1000, "batch_size": 2,
                            "gradient_clipping": 0.5, "checkpoint_dir": "../checkpoints", "save_interval": 500, "log_interval": 50}
    b = cli.Batches(None, 3, 8, rank=1, seed=0)(5)
    assert tuple(b["audio_embedding"].shape) == (3, 512) and tuple(b["image_latents"].shape) == (3, 4, 8, 8)
    assert tuple(b["text_embedding"].shape) == (3, 77, 768) and tuple(b["noise"].shape) == (3, 4, 8, 8)
    assert b["timesteps"].dtype == torch.int64 and int(b["timesteps"].max()) < 1000
    if not torch.cuda.is_available():
        with pytest.raises(SystemExit):
            cli.main(["--num-steps", "1"])
