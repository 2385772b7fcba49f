"""Inference script for CLAP2Diffusion on B200 -- drop-in for the reference's ``scripts/inference.py``.

Same CLI (``--audio --text --output --checkpoint_dir --steps --cfg_scale --seed --no_hierarchical``) and
the same ``AudioToImageInference`` class surface (load_models / load_audio / extract_clap_embedding /
apply_normalization / generate / batch_generate; reference scripts/inference.py:21-180), but ``generate``
runs the real denoising loop on libc2d kernels instead of fabricating a random image
(reference :153-166): CLAP embedding -> hierarchical audio tokens -> audio attention processors on the
SD-1.5 UNet -> 50 x (UNet + CFG + DDIM) -> VAE decode.

Checkpoints (same file names and dict keys as the reference's scripts write / read):
  audio_projector_stage3_finetuned.pth | audio_projector_stage2.pth   {'hierarchical_state_dict': HierarchicalAudioV4,
        'adapter_state_dict': AudioAdapter, ...}          (train_stage3.py:262-271, train_stage2.py:183-189)
  hierarchical_v4_final.pth                               raw HierarchicalAudioV4 state dict (inference.py:53-59)
  unet_adapter_final.pth                                  {'processor_early' | 'processor_mid' | 'processor_late':
        AudioAttnProcessor state dicts[, 'hierarchical_state_dict': ImprovedHierarchicalAudioEncoder, 'mode']} -- the
        reference names the file (inference.py:62-65) but never writes it; this is the layout this package saves.
  unet.pth / vae_decoder.pth (diffusers key layout), clap_audio.pth (HF key layout).
The audio conditioning of the image is switched ON only when trained audio weights were found for BOTH the
hierarchical model and the attention processors: random-init processors (alpha = 0 -> gate 0.5) would add noise to
the text states at all 16 attn2 sites.  Otherwise the run is text-only and says so loudly.

No tokenizer / CLIP text encoder exists on the box (no network): ``--text`` is either a file with precomputed CLIP
text states ([77,768] or [1,77,768]; .npy / .pt) or a prompt string, which is hashed into the synthetic stand-in of
``clap2diffusion_b200.synthetic`` (with a warning).  ``--audio synthetic:<seed>`` generates the 10 s / 48 kHz test clip.
"""
from __future__ import annotations

import argparse
import os
import sys
import zlib
from pathlib import Path

import numpy as np
import torch

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from clap2diffusion_b200 import ops, synthetic  # noqa: E402
from clap2diffusion_b200.models.audio_adapter_v4 import AudioAdapter  # noqa: E402
from clap2diffusion_b200.models.hierarchical_audio_v4 import HierarchicalAudioV4  # noqa: E402
from clap2diffusion_b200.pipeline import AudioToImagePipeline  # noqa: E402


class AudioToImageInference:
    def __init__(self, checkpoint_dir="../checkpoints", device=None, dtype=torch.bfloat16):
        self.checkpoint_dir = Path(checkpoint_dir)
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("CLAP2Diffusion-B200 needs a CUDA device: the compute library has no CPU path")
            device = torch.device("cuda")
        self.device = torch.device(device)
        self.dtype = dtype
        print(f"Initializing inference pipeline on {self.device}")
        self.load_models()
        self.OPTIMAL_NORM = 60.0

    # ------------------------------------------------------------------ checkpoints
    def _load(self, name):
        path = self.checkpoint_dir / name
        if not path.exists():
            return None
        print(f"Loading {name} from {path}")
        return torch.load(path, map_location="cpu", weights_only=True)

    def load_models(self):
        """Loads whatever checkpoints exist (reference file names and dict keys, :34-71) and random-initialises the
        rest.  Sets ``self.audio_conditioning`` (see the module docstring)."""
        hier_sd, hier_src = None, None
        for name in ("audio_projector_stage3_finetuned.pth", "audio_projector_stage2.pth"):
            ck = self._load(name)
            if ck is None:
                continue
            if "adapter_state_dict" in ck and not hasattr(self, "audio_adapter"):
                self.audio_adapter = AudioAdapter().to(self.device).eval()
                self.audio_adapter.load_state_dict(ck["adapter_state_dict"])
            if "hierarchical_state_dict" in ck and hier_sd is None:
                hier_sd, hier_src = ck["hierarchical_state_dict"], name
        ck = self._load("hierarchical_v4_final.pth")
        if ck is not None and hier_sd is None:
            hier_sd, hier_src = ck, "hierarchical_v4_final.pth"
        if hier_sd is not None:
            self.hierarchical_model = HierarchicalAudioV4().to(self.device).eval()
            self.hierarchical_model.load_state_dict(hier_sd)
        unet_sd, vae_sd = self._load("unet.pth"), self._load("vae_decoder.pth")
        ck = self._load("unet_adapter_final.pth")
        mode = (ck or {}).get("mode", "add")
        if unet_sd is None or vae_sd is None:
            print("No SD-1.5 UNet / VAE checkpoint found: using random-init weights (synthetic run)")
            self.pipeline = AudioToImagePipeline.random_init(seed=0, device=self.device, dtype=self.dtype, mode=mode)
        else:
            self.pipeline = AudioToImagePipeline(unet_sd, vae_sd, device=self.device, dtype=self.dtype, mode=mode)
        have_hier = have_proc = False
        if ck is not None:
            if "hierarchical_state_dict" in ck:          # ImprovedHierarchicalAudioEncoder (soft decomposition + router)
                self.pipeline.hier.load_state_dict({k: v.to(self.device) for k, v in ck["hierarchical_state_dict"].items()})
                have_hier, hier_src = True, "unet_adapter_final.pth"
            levels = [lvl for lvl, names in self.pipeline.manager.level_mapping.items() if names]
            for lvl in levels:
                if f"processor_{lvl}" in ck:
                    names = self.pipeline.manager.level_mapping[lvl]
                    self.pipeline.unet.sites[names[0][:-len(".processor")]].processor.load_state_dict(ck[f"processor_{lvl}"])
            have_proc = all(f"processor_{lvl}" in ck for lvl in levels)
        if not have_hier and hasattr(self, "hierarchical_model"):
            # the reference's own checkpoints hold the legacy 5-3-2 model: it drives the processors through its
            # encode() (ambience -> early, background -> mid, foreground -> late)
            self.pipeline.sampler.hier = self.hierarchical_model
            have_hier = True
        self.audio_conditioning = have_hier and have_proc
        if self.audio_conditioning:
            print(f"Audio conditioning ON: hierarchical model from {hier_src}, attention processors from unet_adapter_final.pth")
        else:
            missing = [n for n, ok in (("hierarchical model", have_hier), ("attention processors (unet_adapter_final.pth)", have_proc)) if not ok]
            print("=" * 60 + f"\nWARNING: no trained weights for the {' and the '.join(missing)}: audio conditioning is OFF\n"
                  "         (text-only image; random-init processors would corrupt the text states).\n" + "=" * 60)

    def save_unet_adapter(self, path=None):
        """Writes ``unet_adapter_final.pth`` (the layout load_models reads): one state dict per processor level."""
        path = Path(path) if path is not None else self.checkpoint_dir / "unet_adapter_final.pth"
        ck = {"mode": "add"}
        for lvl, names in self.pipeline.manager.level_mapping.items():
            if names:
                proc = self.pipeline.unet.sites[names[0][:-len(".processor")]].processor
                ck["mode"] = proc.mode
                ck[f"processor_{lvl}"] = {k: v.detach().cpu() for k, v in proc.state_dict().items()}
        torch.save(ck, path)
        return path

    def text_states(self, text_prompt):
        """[1,77,768] text states: a .npy / .pt file of precomputed CLIP states, else the synthetic stand-in."""
        sp = str(text_prompt)
        if sp.endswith((".npy", ".pt")) and os.path.exists(sp):
            a = np.load(sp) if sp.endswith(".npy") else torch.load(sp, map_location="cpu", weights_only=True).float().numpy()
            a = np.asarray(a, dtype=np.float32).reshape(-1, 77, 768)[:1]
            return torch.from_numpy(a).to(self.device, self.dtype)
        if sp and not getattr(self, "_warned_text", False):
            print("  (no CLIP text encoder on this box: the prompt is hashed into synthetic text states; pass a .npy / .pt "
                  "file of [77,768] CLIP states for real conditioning)")
            self._warned_text = True
        return torch.from_numpy(synthetic.text_states(sp)[None]).to(self.device, self.dtype)

    # ------------------------------------------------------------------ audio
    def load_audio(self, audio_path, duration=10):
        """48 kHz mono, at most `duration` seconds, peak-normalised (reference :73-83, without librosa)."""
        print(f"Loading audio from {audio_path}")
        sp = str(audio_path)
        if sp.startswith("synthetic:"):
            return synthetic.synthetic_audio(int(sp.split(":", 1)[1]))
        from scipy.io import wavfile
        from scipy.signal import resample_poly
        sr, a = wavfile.read(sp)
        a = a.astype(np.float32) / (np.iinfo(a.dtype).max if np.issubdtype(a.dtype, np.integer) else 1.0)
        if a.ndim > 1:
            a = a.mean(axis=1)
        if sr != 48000:
            g = np.gcd(int(sr), 48000)
            a = resample_poly(a, 48000 // g, int(sr) // g).astype(np.float32)
        a = a[: 48000 * duration]
        return a / (np.abs(a).max() + 1e-8)

    def extract_clap_embedding(self, audio):
        """Unit-norm [1,512] CLAP embedding of the clip from the GPU HTSAT tower (the reference has a
        ``torch.randn(1, 512)`` placeholder here, :85-90).  Weights: ``clap_audio.pth`` (HF key layout) in the
        checkpoint directory, else ``laion/clap-htsat-unfused`` from the local HF cache, else random init."""
        if getattr(self, "clap_encoder", None) is None:
            from clap2diffusion_b200.models.audio_encoder import CLAPAudioEncoder
            sd = self._load("clap_audio.pth")
            if sd is not None:
                self.clap_encoder = CLAPAudioEncoder(device=str(self.device), state_dict=sd, dtype=self.dtype)
            else:
                try:
                    self.clap_encoder = CLAPAudioEncoder(device=str(self.device), dtype=self.dtype)
                except RuntimeError:
                    print("  no CLAP checkpoint available: random-initialised HTSAT tower")
                    self.clap_encoder = CLAPAudioEncoder.random_init(seed=0, device=str(self.device), dtype=self.dtype)
        return self.clap_encoder.encode_audio(np.asarray(audio, dtype=np.float32), 48000)

    def apply_normalization(self, audio_tokens, target_norm=60.0):
        """x * target / mean(||x||_2) (reference :92-99; batch-coupled mean, identical for batch 1)."""
        return ops.norm_scale(audio_tokens.contiguous(), target_norm, per_sample=False)

    # ------------------------------------------------------------------ generation
    @torch.no_grad()
    def generate(self, audio_path, text_prompt="", num_inference_steps=50, guidance_scale=7.5, seed=None,
                 use_hierarchical=True):
        from PIL import Image
        if seed is not None:
            torch.manual_seed(seed)
            np.random.seed(seed)
        audio = self.load_audio(audio_path)
        clap = self.extract_clap_embedding(audio)
        if hasattr(self, "audio_adapter"):
            tokens = self.apply_normalization(self.audio_adapter(clap), self.OPTIMAL_NORM)
            print(f"Audio tokens shape: {tokens.shape}")
            print(f"Audio tokens norm: {torch.norm(tokens.float()).item():.2f}")
        if use_hierarchical and hasattr(self, "hierarchical_model"):
            tokens_77, hierarchy = self.hierarchical_model(clap, return_intermediate=True)
            print("Hierarchical processing:")
            print(f"  - Output tokens shape: {tokens_77.shape}")
            for k in ("foreground", "background", "ambience"):
                print(f"  - {k.capitalize()} features: {hierarchy[k].shape}")
        print("\nGenerating image with:")
        print(f"  Audio: {Path(str(audio_path)).name}")
        print(f"  Text: {text_prompt}")
        print(f"  Steps: {num_inference_steps}")
        print(f"  CFG Scale: {guidance_scale}")
        noise_seed = seed if seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
        out = self.pipeline.sampler.sample(
            clap, self.text_states(text_prompt),
            torch.from_numpy(synthetic.text_states("")[None]).to(self.device, self.dtype),
            torch.from_numpy(synthetic.init_noise(noise_seed)[None]).to(self.device),
            steps=num_inference_steps, guidance=guidance_scale,
            use_audio=bool(use_hierarchical and self.audio_conditioning), decode=True)
        self.last_latents = out["latents"]
        img = out["image"][0].cpu().numpy().transpose(1, 2, 0)             # [-1,1] NCHW -> HWC uint8 on the host
        return Image.fromarray(((np.clip(img, -1.0, 1.0) + 1.0) * 127.5).astype(np.uint8))

    def batch_generate(self, audio_paths, text_prompts=None, **kwargs):
        if text_prompts is None:
            text_prompts = [""] * len(audio_paths)
        return [self.generate(a, t, **kwargs) for a, t in zip(audio_paths, text_prompts)]


def main(argv=None):
    parser = argparse.ArgumentParser(description="CLAP2Diffusion Inference")
    parser.add_argument("--audio", type=str, required=True, help="Path to audio file (or synthetic:<seed>)")
    parser.add_argument("--text", type=str, default="", help="Text prompt")
    parser.add_argument("--output", type=str, default="output.png", help="Output image path")
    parser.add_argument("--checkpoint_dir", type=str, default="../checkpoints", help="Checkpoint directory")
    parser.add_argument("--steps", type=int, default=50, help="Number of inference steps")
    parser.add_argument("--cfg_scale", type=float, default=7.5, help="Guidance scale")
    parser.add_argument("--seed", type=int, default=None, help="Random seed")
    parser.add_argument("--no_hierarchical", action="store_true", help="Disable hierarchical processing")
    args = parser.parse_args(argv)
    print("\n" + "=" * 60)
    print("CLAP2Diffusion Inference")
    print("=" * 60)
    pipeline = AudioToImageInference(checkpoint_dir=args.checkpoint_dir)
    image = pipeline.generate(audio_path=args.audio, text_prompt=args.text, num_inference_steps=args.steps,
                              guidance_scale=args.cfg_scale, seed=args.seed, use_hierarchical=not args.no_hierarchical)
    image.save(args.output)
    print(f"\n✓ Image saved to {args.output}")


if __name__ == "__main__":
    main()
