"""Pin oracle/clap.py against the UNMODIFIED Hugging Face CLAP (the third-party arithmetic behind
/root/reference/models/audio_encoder.py:47-48,164-171) and write tests/golden/clap_*.npz.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_clap

Build container only (needs `transformers`); the fixtures hold inputs / reference outputs, weights are regenerated
from (name, shape, kind, seed) by oracle/weights.py on any machine.
"""
from __future__ import annotations

import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 4321


def main():
    from transformers import ClapConfig, ClapFeatureExtractor
    from transformers.models.clap.modeling_clap import ClapAudioModel, ClapProjectionLayer

    from oracle import clap as C
    from oracle.pipeline import rel_l2 as _rel

    def rel_l2(a, b):
        return _rel(torch.as_tensor(np.asarray(a)) if not torch.is_tensor(a) else a,
                    torch.as_tensor(np.asarray(b)) if not torch.is_tensor(b) else b)
    from oracle.weights import count, synth_state_dict
    from clap2diffusion_b200.synthetic import synthetic_audio

    torch.manual_seed(0)
    cfg = ClapConfig()
    audio_model = ClapAudioModel(cfg.audio_config).eval()
    proj = ClapProjectionLayer(cfg.audio_config).eval()
    spec = C.clap_audio_spec()
    sd = synth_state_dict(spec, SEED)
    # the spec must be exactly the HF parameter / buffer set (minus derived buffers)
    hf = {f"audio_model.{k}": v for k, v in audio_model.state_dict().items()}
    hf.update({f"audio_projection.{k}": v for k, v in proj.state_dict().items()})
    derived = {k for k in hf if k.endswith("relative_position_index") or k.endswith("num_batches_tracked")}
    assert set(hf) - derived == set(sd), sorted((set(hf) - derived) ^ set(sd))[:10]
    for k, v in sd.items():
        assert tuple(hf[k].shape) == v.shape, (k, hf[k].shape, v.shape)
    n_params = sum(p.numel() for p in audio_model.parameters()) + sum(p.numel() for p in proj.parameters())
    print("CLAP audio tower + projection parameters:", n_params, "spec (incl. BN running stats):", count(spec))
    audio_model.load_state_dict({k[len("audio_model."):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("audio_model.")},
                                strict=False)
    proj.load_state_dict({k[len("audio_projection."):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("audio_projection.")})

    waves = np.stack([synthetic_audio(s) for s in (0, 1)])
    waves[1] *= np.linspace(0.05, 1.0, waves.shape[1], dtype=np.float32)          # non-stationary second clip
    fe = ClapFeatureExtractor(truncation="rand_trunc", padding="repeatpad")
    feats = fe(list(waves), sampling_rate=48000, return_tensors="pt")
    mel_hf = feats["input_features"].numpy()                                       # [2,1,1001,64]
    assert rel_l2(np.stack([C.log_mel(w) for w in waves])[:, None], mel_hf) < 1e-6
    assert np.allclose(C.mel_filters_slaney(), fe.mel_filters_slaney, rtol=1e-10, atol=1e-12)

    hidden = {}
    with torch.no_grad():
        out = audio_model(input_features=feats["input_features"], is_longer=feats["is_longer"], output_hidden_states=True)
        emb_hf = torch.nn.functional.normalize(proj(out.pooler_output), dim=-1)
        W = {k: torch.from_numpy(v) for k, v in sd.items()}
        taps = {}
        emb_or = C.tower_forward(W, torch.from_numpy(mel_hf), taps)
        enc_or = C.encode_audio(W, waves)
    print("oracle vs HF: embedding rel_l2 = %.2e, encode_audio rel_l2 = %.2e" % (rel_l2(emb_or, emb_hf), rel_l2(enc_or, emb_hf)))
    assert rel_l2(emb_or, emb_hf) < 5e-6 and rel_l2(enc_or, emb_hf) < 5e-6
    # HF hidden states: (patch embeddings, stage outputs after down-sampling ...) as [B,C,H,W]
    hs = out.hidden_states
    assert rel_l2(taps["patch_embed"], hs[0].flatten(2).transpose(1, 2)) < 5e-6
    np.savez_compressed(os.path.join(GOLD, "clap_audio.npz"), seeds=np.array([0, 1]), mel=mel_hf.astype(np.float32),
                        mel_sub=mel_hf[:, 0, ::50].astype(np.float32), patch_embed=taps["patch_embed"][:, ::64].numpy(),
                        stage0=taps["stage0"][:, ::64].numpy(), stage2=taps["stage2"][:, ::16].numpy(),
                        pooled=out.pooler_output.numpy(), embedding=emb_hf.numpy(), n_params=np.array(n_params))
    print("wrote clap_audio.npz")


if __name__ == "__main__":
    main()
