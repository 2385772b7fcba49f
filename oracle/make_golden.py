"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (imported from
/root/reference, build container only) on portable synthetic weights, and check the oracle
restatement (oracle/audio.py) against them.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden            # audio-side fixtures (seconds)
    python -m oracle.make_golden --unet     # + oracle-generated UNet/pipeline fixtures (minutes, UNPINNED)

The fixtures hold only inputs and reference outputs; weights are regenerated from
(name, shape, kind, seed) by oracle/weights.py on any machine.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
SEED = 1234


def _load_ref():
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted; goldens can only be regenerated in the build container")
    sys.path.insert(0, os.path.join(HERE, "_diffusers_standin"))
    sys.path.insert(0, REF)
    import models.audio_adapter_v4 as ad
    import models.hierarchical_audio_v4 as hi
    import models.audio_attention_processor as ap
    import diffusers.models.attention_processor as dap
    return ad, hi, ap, dap


def _load(module, sd_np, strict=True):
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
    cur = module.state_dict()
    assert set(cur.keys()) == set(sd.keys()), (sorted(set(cur) ^ set(sd)))
    for k in cur:
        assert tuple(cur[k].shape) == tuple(sd[k].shape), (k, cur[k].shape, sd[k].shape)
    module.load_state_dict(sd, strict=strict)
    module.eval()


def _chk(name, a, b, tol=2e-6):
    from oracle.pipeline import rel_l2
    e = rel_l2(a, b)
    print(f"  {name:38s} rel_l2 oracle-vs-reference = {e:.2e}")
    assert e < tol, (name, e)


def audio_goldens():
    from oracle import audio as A
    from oracle.pipeline import clap_embedding, to_torch
    from oracle.weights import synth_state_dict, count
    ad, hi, ap, dap = _load_ref()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    clap = torch.from_numpy(np.stack([clap_embedding(s) for s in (11, 12, 13)]))

    # ---- AudioAdapter -----------------------------------------------------------------
    m = ad.AudioAdapter()
    spec = A.audio_adapter_spec()
    assert count(spec) == 16_510_464 == sum(p.numel() for p in m.parameters())
    sd = synth_state_dict(spec, SEED)
    _load(m, sd)
    with torch.no_grad():
        ref = m(clap)
    ora = A.audio_adapter_forward(to_torch(sd), clap)
    _chk("AudioAdapter", ora, ref)
    # Norm-60 (scripts/inference.py:92-99), restated inline here because that script needs librosa
    nrm = torch.norm(ref, dim=-1, keepdim=True).mean()
    ref60 = ref * (60.0 / nrm)
    _chk("norm60 (batch-coupled)", A.norm60(ref), ref60)
    np.savez_compressed(os.path.join(GOLD, "audio_adapter.npz"), seed=SEED, clap=clap.numpy(),
                        tokens=ref.numpy(), tokens_norm60=ref60.numpy())

    # ---- ImprovedHierarchicalAudioEncoder ----------------------------------------------
    m = hi.ImprovedHierarchicalAudioEncoder()
    spec = A.improved_hier_spec()
    assert count(spec) == 3_840_766 == sum(p.numel() for p in m.parameters())
    sd = synth_state_dict(spec, SEED)
    for k, v in A.IMPROVED_BUFFERS.items():
        sd[k] = np.asarray(v, dtype=np.float32)
    _load(m, sd)
    with torch.no_grad():
        t77, info = m(clap, return_all=True)
    ora = A.improved_hier_forward(to_torch(sd), clap)
    _chk("Improved.tokens_77", ora["tokens_77"], t77)
    _chk("Improved.tokens_10", ora["tokens_10"], info["tokens_10"])
    _chk("Improved.assignments", ora["assignments"], info["assignments"])
    _chk("Improved.hierarchy_weights", ora["hierarchy_weights"], info["hierarchy_weights"])
    for lvl in ("early", "mid", "late"):
        _chk(f"Improved.routed[{lvl}]", ora["routed"][lvl], info["routed"][lvl])
    # temperature 0.5 variant (TemperatureScheduler end state)
    m.decomposer.set_temperature(0.5)
    with torch.no_grad():
        _, info05 = m(clap, return_all=True)
    ora05 = A.improved_hier_forward(to_torch(sd), clap, temperature=0.5)
    _chk("Improved.assignments(T=0.5)", ora05["assignments"], info05["assignments"])
    np.savez_compressed(
        os.path.join(GOLD, "improved_hier.npz"), seed=SEED, clap=clap.numpy(), tokens_77=t77.numpy(),
        tokens_10=info["tokens_10"].numpy(), assignments=info["assignments"].numpy(),
        hierarchy_weights=info["hierarchy_weights"].numpy(),
        routed_early=info["routed"]["early"].numpy(), routed_mid=info["routed"]["mid"].numpy(),
        routed_late=info["routed"]["late"].numpy(), assignments_T05=info05["assignments"].numpy())

    # TemperatureScheduler known answers (hierarchical_audio_v4.py:20-76)
    sch = hi.TemperatureScheduler(m.decomposer, T_max=2.0, T_min=0.5, total_steps=2000)
    temps = {}
    for step in (0, 100, 200, 500, 1100, 1500, 2000, 2500):
        sch.step(step)
        temps[step] = float(m.decomposer.temperature)
    print("  temperature KAT:", temps)
    np.savez(os.path.join(GOLD, "temperature_kat.npz"), steps=np.array(list(temps.keys())),
             temps=np.array(list(temps.values()), dtype=np.float64))

    # ---- legacy HierarchicalAudioV4 -----------------------------------------------------
    m = hi.HierarchicalAudioV4()
    spec = A.legacy_hier_spec()
    assert count(spec) == 12_843_395 == sum(p.numel() for p in m.parameters())
    sd = synth_state_dict(spec, SEED)
    _load(m, sd)
    with torch.no_grad():
        t77, hz = m(clap, return_intermediate=True)
    ora = A.legacy_hier_forward(to_torch(sd), clap)
    _chk("Legacy.tokens77", ora["tokens_77"], t77)
    for k in ("foreground", "background", "ambience", "weights", "tokens10"):
        _chk(f"Legacy.{k}", ora[k], hz[k])
    np.savez_compressed(os.path.join(GOLD, "legacy_hier.npz"), seed=SEED, clap=clap.numpy(),
                        tokens_77=t77.numpy(), tokens10=hz["tokens10"].numpy(),
                        foreground=hz["foreground"].numpy(), background=hz["background"].numpy(),
                        ambience=hz["ambience"].numpy(), weights=hz["weights"].numpy())

    # ---- AudioAttnProcessor at the four SD-1.5 (N,C) shapes, both modes ----------------
    from oracle.pipeline import np_randn
    g = torch.Generator().manual_seed(7)
    ehs = torch.from_numpy(np_randn("ehs", (2, 77, 768)))
    audio = torch.from_numpy(np_randn("audio10", (2, 10, 768))) * 0.3
    out = dict(seed=SEED)
    psd = synth_state_dict(A.attn_processor_spec(), SEED)
    assert count(A.attn_processor_spec()) == 99_137
    for (N, C) in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
        attn = dap.Attention(C, cross_attention_dim=768, heads=8, dim_head=C // 8)
        asd = synth_state_dict(A.attn_site_spec(C), SEED, prefix=f"site{C}.")
        asd = {k.split(".", 1)[1]: v for k, v in asd.items()}
        _load(attn, asd)
        h = torch.from_numpy(np_randn(f"h_{N}_{C}", (2, N, C)))
        rows = np.arange(0, N, max(1, N // 64))
        out[f"rows_{N}_{C}"] = rows
        for mode in ("add", "concat"):
            proc = ap.AudioAttnProcessor(level="mid", mode=mode)
            _load(proc, psd)
            with torch.no_grad():
                ref = proc(attn, h, encoder_hidden_states=ehs, audio={"mid": audio})
                ref_noaudio = proc(attn, h, encoder_hidden_states=ehs)
            ora = A.processor_call(to_torch(psd), to_torch(asd), 8, h, ehs, audio, mode)
            _chk(f"AudioAttnProcessor[{mode}] N={N} C={C}", ora, ref, tol=5e-6)
            ora_na = A.processor_call(to_torch(psd), to_torch(asd), 8, h, ehs, None, mode)
            _chk(f"  no-audio fall-through N={N}", ora_na, ref_noaudio, tol=5e-6)
            out[f"out_{mode}_{N}_{C}"] = ref.numpy()[:, rows]       # sampled query rows
        out[f"out_noaudio_{N}_{C}"] = ref_noaudio.numpy()[:, rows]
        # 4-D input path (:67-70, :137-138) at the smallest site
        if N == 64:
            h4 = h.transpose(1, 2).reshape(2, C, 8, 8).contiguous()
            proc = ap.AudioAttnProcessor(level="mid", mode="add"); _load(proc, psd)
            with torch.no_grad():
                ref4 = proc(attn, h4, encoder_hidden_states=ehs, audio={"mid": audio})
            _chk("AudioAttnProcessor 4-D input", ref4.reshape(2, C, 64).transpose(1, 2),
                 A.processor_call(to_torch(psd), to_torch(asd), 8, h, ehs, audio, "add"), tol=5e-6)
    np.savez_compressed(os.path.join(GOLD, "attn_processor.npz"), **out)

    # ---- AudioCrossAttention (gated branch) --------------------------------------------
    C, N = 320, 256
    m = ad.AudioCrossAttention(query_dim=C)
    spec = A.gated_xattn_spec(C)
    assert count(spec) == 1_115_073 == sum(p.numel() for p in m.parameters())
    sd = synth_state_dict(spec, SEED)
    _load(m, sd)
    h = torch.from_numpy(np_randn("gx_h", (2, N, C)))
    a16 = torch.from_numpy(np_randn("gx_a16", (2, 16, 768)))
    mask = torch.ones(2, 1, 1, 16, dtype=torch.bool); mask[:, :, :, 12:] = False
    with torch.no_grad():
        ref = m(h, a16)
        refm = m(h, a16, mask)
    _chk("AudioCrossAttention", A.gated_xattn_forward(to_torch(sd), h, a16), ref)
    _chk("AudioCrossAttention(mask)", A.gated_xattn_forward(to_torch(sd), h, a16, mask=mask), refm)
    np.savez_compressed(os.path.join(GOLD, "gated_xattn.npz"), seed=SEED, out=ref.numpy(), out_masked=refm.numpy(), mask=mask.numpy())

    # ---- AudioProcessorManager level census --------------------------------------------
    from oracle import sd15

    class _FakeUNet:
        def __init__(self):
            self.attn_processors = {n: object() for n in sd15.attn_processor_names()}
    mgr = ap.AudioProcessorManager(_FakeUNet())
    census = {k: len(v) for k, v in mgr.level_mapping.items()}
    print("  level census:", census)
    assert census == {"early": 4, "mid": 7, "late": 5}
    for lvl, names in mgr.level_mapping.items():
        for n in names:
            assert A.level_of_site(n) == lvl, (n, lvl)
    import json
    with open(os.path.join(GOLD, "level_mapping.json"), "w") as f:
        json.dump({k: sorted(v) for k, v in mgr.level_mapping.items()}, f, indent=1)
    print("audio goldens written to", GOLD)


def mask_goldens():
    """attention_mask pass-through (reference audio_attention_processor.py:129 -> attn.get_attention_scores): the
    UNMODIFIED reference processor with a key-padding mask in the form diffusers prepares it ([B*heads, 1, T], additive
    0 / -10000), 'add' mode, at the (256, 1280) site.  Sample 0 keeps 60 keys, sample 1 keeps 77."""
    from oracle import audio as A
    from oracle.pipeline import np_randn
    from oracle.weights import synth_state_dict
    ad, hi, ap, dap = _load_ref()
    N, C, heads = 256, 1280, 8
    ehs = torch.from_numpy(np_randn("ehs", (2, 77, 768)))
    audio = torch.from_numpy(np_randn("audio10", (2, 10, 768))) * 0.3
    psd = synth_state_dict(A.attn_processor_spec(), SEED)
    attn = dap.Attention(C, cross_attention_dim=768, heads=heads, dim_head=C // heads)
    asd = synth_state_dict(A.attn_site_spec(C), SEED, prefix=f"site{C}.")
    asd = {k.split(".", 1)[1]: v for k, v in asd.items()}
    _load(attn, asd)
    h = torch.from_numpy(np_randn(f"h_{N}_{C}", (2, N, C)))
    keep = torch.ones(2, 77, dtype=torch.bool)
    keep[0, 60:] = False
    bias = torch.zeros(2, 77).masked_fill(~keep, -10000.0)
    mask = bias[:, None, None, :].expand(2, heads, 1, 77).reshape(2 * heads, 1, 77).contiguous()
    proc = ap.AudioAttnProcessor(level="mid", mode="add")
    _load(proc, psd)
    with torch.no_grad():
        ref = proc(attn, h, encoder_hidden_states=ehs, attention_mask=mask, audio={"mid": audio})
        ref_nomask = proc(attn, h, encoder_hidden_states=ehs, audio={"mid": audio})
    assert float((ref[0] - ref_nomask[0]).abs().max()) > 1e-3 and torch.equal(ref[1], ref_nomask[1])
    rows = np.arange(0, N, 4)
    np.savez_compressed(os.path.join(GOLD, "attn_processor_mask.npz"), seed=SEED, keep=keep.numpy(), rows=rows, out=ref.numpy()[:, rows])
    print("mask golden written")


def unet_goldens():
    """Oracle-generated (UNPINNED) UNet / pipeline vectors for the CUDA parity tests."""
    import time
    from oracle import pipeline as PL, sd15
    torch.set_num_threads(os.cpu_count())
    W = PL.build_weights(seed=0, with_vae=True)
    ctx_c = torch.from_numpy(PL.text_states("a beach"))[None]
    ctx_u = torch.from_numpy(PL.text_states(""))[None]
    clap = torch.from_numpy(PL.clap_embedding(0))[None]
    # (1) one UNet forward at a reduced 16x16 latent (fast CPU check) with the audio hook
    x = torch.from_numpy(PL.init_noise(5, 16, 16))[None]
    hier = __import__("oracle.audio", fromlist=["x"]).improved_hier_forward(W["hier"], clap)
    hook = PL.make_attn2_hook(W, hier["routed"], "add")
    taps = {}
    with torch.no_grad():
        t0 = time.time()
        eps = sd15.unet_forward(W["unet"], x, 500.0, ctx_c, hook, taps=taps)
        print(f"  unet 16x16 forward {time.time()-t0:.1f}s")
    np.savez_compressed(os.path.join(GOLD, "unet_16x16.npz"), eps=eps.numpy(), t=500.0,
                        conv_in=taps["conv_in"].numpy(), mid=taps["mid"].numpy())
    # (2) config 1: batch 1, 20 DDIM steps, CFG 7.5, 64x64, fp32 -> per-step latents
    noise = torch.from_numpy(PL.init_noise(0))[None]
    t0 = time.time()
    out = PL.sample(W, clap, ctx_c, ctx_u, noise, steps=20, guidance=7.5, decode=True)
    dt = time.time() - t0
    print(f"  config1 (20 steps + decode) {dt:.1f}s on {os.cpu_count()} cores")
    np.savez_compressed(os.path.join(GOLD, "pipeline_cfg1_20steps.npz"),
                        latents=torch.cat(out["latents"], 0).numpy(),
                        image=out["image"].numpy().astype(np.float16), seconds=dt, cores=os.cpu_count())
    # (3) config 2 prefix: 50-step schedule, first 6 steps + eps of step 0 (full 50 is covered at
    #     test time on the GPU by the oracle running on CUDA)
    out = PL.sample(W, clap, ctx_c, ctx_u, noise, steps=50, guidance=7.5, max_steps=6)
    np.savez_compressed(os.path.join(GOLD, "pipeline_cfg2_first6.npz"),
                        latents=torch.cat(out["latents"], 0).numpy(), eps0=out["eps"][0].numpy())
    print("unet goldens written")


def full_trajectory_goldens():
    """Oracle-generated (UNPINNED restatement, see oracle/__init__.py) full 50-step trajectories for the configurations
    the benchmark numbers are quoted on (BASELINE.json configs[1], configs[2]):
      * config 2: ONE image ("a beach", seed 0), 50 DDIM steps, CFG 7.5 -> all 50 per-step latents + decoded image;
      * config 3: two more members of a micro-batch-8 prompt x seed sweep (slots 2 and 5 of `config3_jobs()`), 50 steps each.
    The oracle treats every image independently (per-sample Norm-60, D3), so one image at a time is the batch's definition."""
    import time
    from oracle import pipeline as PL
    torch.set_num_threads(os.cpu_count())
    W = PL.build_weights(seed=0, with_vae=True)
    ctx_u = torch.from_numpy(PL.text_states(""))[None]
    t0 = time.time()
    out = PL.sample(W, torch.from_numpy(PL.clap_embedding(0))[None], torch.from_numpy(PL.text_states("a beach"))[None], ctx_u,
                    torch.from_numpy(PL.init_noise(0))[None], steps=50, guidance=7.5, decode=True)
    print(f"  config2 (50 steps + decode) {time.time()-t0:.1f}s on {os.cpu_count()} cores")
    np.savez_compressed(os.path.join(GOLD, "pipeline_cfg2_50steps.npz"), latents=torch.cat(out["latents"], 0).numpy(),
                        image=out["image"].numpy().astype(np.float16))
    jobs = PL.config3_jobs()
    res = {}
    for slot in (2, 5):
        prompt, seed = jobs[slot]
        t0 = time.time()
        o = PL.sample(W, torch.from_numpy(PL.clap_embedding(seed))[None], torch.from_numpy(PL.text_states(prompt))[None], ctx_u,
                      torch.from_numpy(PL.init_noise(seed))[None], steps=50, guidance=7.5)
        print(f"  config3 slot {slot} ({prompt!r}, seed {seed}) {time.time()-t0:.1f}s")
        res[f"latents_slot{slot}"] = torch.cat(o["latents"], 0).numpy()
    np.savez_compressed(os.path.join(GOLD, "pipeline_cfg3_mb8_slots.npz"), slots=np.array([2, 5]), **res)
    print("full-trajectory goldens written")


if __name__ == "__main__":
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--unet", action="store_true")
    ap_.add_argument("--only-unet", action="store_true")
    ap_.add_argument("--only-full", action="store_true", help="only the 50-step config-2 / config-3 trajectories")
    ap_.add_argument("--only-mask", action="store_true", help="only the attention_mask golden of the processor")
    a = ap_.parse_args()
    sys.path.insert(0, ROOT)
    if a.only_full:
        full_trajectory_goldens()
        sys.exit(0)
    if a.only_mask:
        mask_goldens()
        sys.exit(0)
    if not a.only_unet:
        audio_goldens()
    if a.unet or a.only_unet:
        unet_goldens()
