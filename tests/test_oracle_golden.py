"""The oracle restatement reproduces the vectors generated from the UNMODIFIED reference modules
(oracle/make_golden.py, build container) -- this is what pins the oracle."""
import json
import os

import numpy as np
import torch

from oracle import audio as A
from oracle import sd15
from oracle.pipeline import clap_embedding, np_randn, rel_l2, to_torch
from oracle.weights import count, synth_state_dict

TOL = 2e-6


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_param_counts():
    assert count(sd15.unet_spec()) == 859_520_964
    assert count(sd15.vae_decoder_spec()) == 49_490_179 + 20
    assert count(A.audio_adapter_spec()) == 16_510_464
    assert count(A.improved_hier_spec()) == 3_840_766
    assert count(A.legacy_hier_spec()) == 12_843_395
    assert count(A.attn_processor_spec()) == 99_137
    assert count(A.gated_xattn_spec(320)) == 1_115_073


def test_attention_site_census_and_levels(gold):
    names = sd15.attn_processor_names()
    assert len(names) == 32 and sum("attn1" in n for n in names) == 16
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "level_mapping.json")))
    assert {k: len(v) for k, v in ref.items()} == {"early": 4, "mid": 7, "late": 5}
    for lvl, sites in ref.items():
        for n in sites:
            assert A.level_of_site(n) == lvl


def test_ddim_timesteps():
    ts = sd15.ddim_timesteps(50)
    assert ts[:3] == [981, 961, 941] and ts[-1] == 1 and len(ts) == 50
    assert sd15.ddim_timesteps(20)[:2] == [951, 901]


def test_audio_adapter_golden(gold):
    g = gold("audio_adapter.npz")
    sd = to_torch(synth_state_dict(A.audio_adapter_spec(), int(g["seed"])))
    out = A.audio_adapter_forward(sd, _t(g["clap"]))
    assert rel_l2(out, _t(g["tokens"])) < TOL
    assert rel_l2(A.norm60(out), _t(g["tokens_norm60"])) < TOL
    # every token has norm ~ sqrt(768) after the final LayerNorm; Norm-60 scale ~ 2.165 (SURVEY §8c KAT iv)
    assert abs(float(out.norm(dim=-1).mean()) - 27.7) < 1.5


def test_improved_hier_golden(gold):
    g = gold("improved_hier.npz")
    sd = synth_state_dict(A.improved_hier_spec(), int(g["seed"]))
    for k, v in A.IMPROVED_BUFFERS.items():
        sd[k] = np.asarray(v, dtype=np.float32)
    out = A.improved_hier_forward(to_torch(sd), _t(g["clap"]))
    for k in ("tokens_77", "tokens_10", "assignments", "hierarchy_weights"):
        assert rel_l2(out[k], _t(g[k])) < TOL, k
    for lvl in ("early", "mid", "late"):
        assert rel_l2(out["routed"][lvl], _t(g[f"routed_{lvl}"])) < TOL
    out05 = A.improved_hier_forward(to_torch(sd), _t(g["clap"]), temperature=0.5)
    assert rel_l2(out05["assignments"], _t(g["assignments_T05"])) < TOL


def test_legacy_hier_golden(gold):
    g = gold("legacy_hier.npz")
    sd = to_torch(synth_state_dict(A.legacy_hier_spec(), int(g["seed"])))
    out = A.legacy_hier_forward(sd, _t(g["clap"]))
    assert rel_l2(out["tokens_77"], _t(g["tokens_77"])) < TOL
    for k in ("tokens10", "foreground", "background", "ambience", "weights"):
        assert rel_l2(out[k], _t(g[k])) < TOL, k


def test_attn_processor_golden(gold):
    g = gold("attn_processor.npz")
    seed = int(g["seed"])
    psd = to_torch(synth_state_dict(A.attn_processor_spec(), seed))
    ehs = _t(np_randn("ehs", (2, 77, 768)))
    audio = _t(np_randn("audio10", (2, 10, 768))) * 0.3
    for (N, C) in ((1024, 640), (256, 1280), (64, 1280)):      # (4096,320) is covered by the GPU tests
        asd = synth_state_dict(A.attn_site_spec(C), seed, prefix=f"site{C}.")
        asd = to_torch({k.split(".", 1)[1]: v for k, v in asd.items()})
        h = _t(np_randn(f"h_{N}_{C}", (2, N, C)))
        rows = g[f"rows_{N}_{C}"]
        for mode in ("add", "concat"):
            out = A.processor_call(psd, asd, 8, h, ehs, audio, mode)
            assert rel_l2(out[:, rows], _t(g[f"out_{mode}_{N}_{C}"])) < 5e-6, (mode, N)
        out = A.processor_call(psd, asd, 8, h, ehs, None, "add")
        assert rel_l2(out[:, rows], _t(g[f"out_noaudio_{N}_{C}"])) < 5e-6


def test_gated_xattn_golden(gold):
    g = gold("gated_xattn.npz")
    sd = to_torch(synth_state_dict(A.gated_xattn_spec(320), int(g["seed"])))
    h, a = _t(np_randn("gx_h", (2, 256, 320))), _t(np_randn("gx_a16", (2, 16, 768)))
    assert rel_l2(A.gated_xattn_forward(sd, h, a), _t(g["out"])) < TOL
    assert rel_l2(A.gated_xattn_forward(sd, h, a, mask=_t(g["mask"])), _t(g["out_masked"])) < TOL


def test_temperature_kat(gold):
    g = gold("temperature_kat.npz")
    kat = dict(zip(g["steps"].tolist(), g["temps"].tolist()))
    assert kat[0] == 2.0 and kat[200] == 2.0 and abs(kat[500] - 1.8995) < 1e-4
    assert abs(kat[1100] - 1.25) < 1e-6 and kat[2000] == 0.5


def test_unet_oracle_is_deterministic(gold):
    """UNPINNED part: the restated UNet at least reproduces its own committed vector on this machine."""
    from oracle import pipeline as PL
    g = gold("unet_16x16.npz")
    W = PL.build_weights(seed=0, with_vae=False)
    clap = _t(clap_embedding(0))[None]
    hier = A.improved_hier_forward(W["hier"], clap)
    hook = PL.make_attn2_hook(W, hier["routed"], "add")
    x = _t(PL.init_noise(5, 16, 16))[None]
    with torch.no_grad():
        eps = sd15.unet_forward(W["unet"], x, float(g["t"]), _t(PL.text_states("a beach"))[None], hook)
    assert rel_l2(eps, _t(g["eps"])) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# Known answers for the UNPINNED restatement (oracle/sd15.py: diffusers is absent, so these published /
# definitional constants of SD-1.5 are what anchors it beyond the parameter counts).
# ---------------------------------------------------------------------------------------------------------
def test_sd15_scheduler_known_answers():
    ac = sd15.alphas_cumprod()
    assert len(ac) == 1000
    assert abs(float(ac[0]) - 0.99915) < 1e-7                  # 1 - beta_start (scaled_linear 0.00085 .. 0.012)
    assert abs(float(ac[999]) - 0.004660) < 5e-6               # SD-1.x terminal alpha-bar (published 0.00466)
    sig = ((1 - ac) / ac) ** 0.5
    assert abs(float(sig[999]) - 14.6146) < 1e-3               # EulerDiscrete sigma_max of SD-1.5
    assert abs(float(sig[0]) - 0.0292) < 1e-4                  # sigma_min
    # DDIM, eta 0, steps_offset 1, set_alpha_to_one False: the last step (t = 1) lands on alpha-bar[0], not on 1
    t, ca, cb = sd15.ddim_coeffs(50)[-1]
    assert t == 1
    assert abs(ca - (float(ac[0]) / float(ac[1])) ** 0.5) < 1e-9
    # x_prev = sqrt(a_p) x0 + sqrt(1 - a_p) eps with x0 = (x - sqrt(1 - a_t) eps) / sqrt(a_t), checked on numbers
    a_t, a_p = float(ac[1]), float(ac[0])
    x, e = 0.7, -0.3
    x0 = (x - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
    assert abs(ca * x + cb * e - (a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * e)) < 1e-9
    # Euler ('leading' spacing): first sigma of the 50-step plan is sigma(t = 981), sigmas end at 0
    ts, sg = sd15.euler_sigmas(50)
    assert ts[0] == 981.0 and sg[-1] == 0.0 and abs(sg[0] - float(sig[981])) < 1e-6
    # CFG: uncond first, cond second; eps = eu + g (ec - eu)
    e2 = torch.tensor([[1.0], [3.0]])
    assert float(sd15.cfg_combine(e2, 7.5)) == 1.0 + 7.5 * 2.0
    assert sd15.VAE_SCALING == 0.18215


def test_sd15_timestep_embedding_order():
    """flip_sin_to_cos=True, freq_shift=0 (SD-1.5 UNet config): [cos | sin], frequencies 10000^(-i/160)."""
    e = sd15.timestep_embedding(torch.tensor([0.0, 1.0, 500.0]), 320)
    assert e.shape == (3, 320)
    assert torch.all(e[0, :160] == 1.0) and torch.all(e[0, 160:] == 0.0)
    assert abs(float(e[1, 0]) - np.cos(1.0)) < 1e-6 and abs(float(e[1, 160]) - np.sin(1.0)) < 1e-6
    assert abs(float(e[2, 80]) - np.cos(500.0 * 10000.0 ** (-0.5))) < 1e-4
    assert abs(float(e[2, 160 + 159]) - np.sin(500.0 * 10000.0 ** (-159.0 / 160.0))) < 1e-5


def test_sd15_geglu_chunk_order_and_block_wiring():
    """GEGLU = value * gelu(gate) with value = FIRST half of the projection (diffusers GEGLU.forward), exact-erf GELU;
    proj_in / proj_out are 1x1 convs around the block with the outer residual added after proj_out."""
    C, name = 64, "t"
    tb = f"{name}.transformer_blocks.0"
    sd = {f"{name}.norm.weight": torch.ones(C), f"{name}.norm.bias": torch.zeros(C),
          f"{name}.proj_in.weight": torch.zeros(C, C, 1, 1), f"{name}.proj_in.bias": torch.full((C,), 0.25),
          f"{name}.proj_out.weight": torch.eye(C).reshape(C, C, 1, 1), f"{name}.proj_out.bias": torch.zeros(C)}
    for n in ("norm1", "norm2", "norm3"):
        sd[f"{tb}.{n}.weight"] = torch.ones(C)
        sd[f"{tb}.{n}.bias"] = torch.zeros(C)
    for a, kd in (("attn1", C), ("attn2", 768)):
        sd[f"{tb}.{a}.to_q.weight"] = torch.zeros(C, C)
        sd[f"{tb}.{a}.to_k.weight"] = torch.zeros(C, kd)
        sd[f"{tb}.{a}.to_v.weight"] = torch.zeros(C, kd)
        sd[f"{tb}.{a}.to_out.0.weight"] = torch.zeros(C, C)
        sd[f"{tb}.{a}.to_out.0.bias"] = torch.zeros(C)
    sd[f"{tb}.ff.net.0.proj.weight"] = torch.zeros(8 * C, C)
    sd[f"{tb}.ff.net.0.proj.bias"] = torch.cat([torch.full((4 * C,), 2.0), torch.full((4 * C,), 1.0)])   # value 2, gate 1
    w2 = torch.zeros(C, 4 * C)
    w2[torch.arange(C), torch.arange(C)] = 1.0
    sd[f"{tb}.ff.net.2.weight"] = w2
    sd[f"{tb}.ff.net.2.bias"] = torch.zeros(C)
    x = torch.zeros(1, C, 4, 4)
    out = sd15.transformer_2d(sd, name, x, torch.zeros(1, 77, 768))
    want = 0.25 + 2.0 * 0.5 * (1.0 + float(torch.erf(torch.tensor(1.0 / 2 ** 0.5))))       # 2 * gelu(1) = 1.68269
    assert abs(want - 0.25 - 1.682689) < 1e-5
    assert torch.allclose(out, torch.full_like(out, want), atol=1e-6)
    assert abs(float(out[0, 0, 0, 0]) - (0.25 + 1.0 * 0.5 * (1.0 + float(torch.erf(torch.tensor(2.0 / 2 ** 0.5)))))) > 0.2  # swapped order would give gate*gelu(value)
