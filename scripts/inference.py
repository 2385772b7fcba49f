"""Inference script for CLAP2Diffusion on B200 -- drop-in for the reference's ``scripts/inference.py``.

Same CLI (``--audio --text --output --checkpoint_dir --steps --cfg_scale --seed --no_hierarchical``) and
the same ``AudioToImageInference`` class surface (load_models / load_audio / extract_clap_embedding /
apply_normalization / generate / batch_generate; reference scripts/inference.py:21-180), but ``generate``
runs the real denoising loop on libc2d kernels instead of fabricating a random image
(reference :153-166): CLAP embedding -> hierarchical audio tokens -> audio attention processors on the
SD-1.5 UNet -> 50 x (UNet + CFG + DDIM) -> VAE decode.

No checkpoints, tokenizer files or pretrained CLAP exist on the box (no network): missing weights are
random-initialised, the text states are the synthetic stand-in of ``clap2diffusion_b200.synthetic`` and the
CLAP embedding is a deterministic stand-in derived from the waveform (the reference itself uses
``torch.randn(1, 512)``, :85-90).  ``--audio synthetic:<seed>`` generates the 10 s / 48 kHz test clip.
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
        """Loads whatever checkpoints exist (same file names and dict keys as the reference, :34-71) and
        random-initialises the rest."""
        ck = self._load("audio_projector_stage2.pth")
        if ck is not None:
            self.audio_adapter = AudioAdapter().to(self.device).eval()
            if "adapter_state_dict" in ck:
                self.audio_adapter.load_state_dict(ck["adapter_state_dict"])
        ck = self._load("hierarchical_v4_final.pth")
        if ck is not None:
            self.hierarchical_model = HierarchicalAudioV4().to(self.device).eval()
            self.hierarchical_model.load_state_dict(ck)
        unet_sd, vae_sd = self._load("unet.pth"), self._load("vae_decoder.pth")
        if unet_sd is None or vae_sd is None:
            print("No SD-1.5 UNet / VAE checkpoint found: using random-init weights (synthetic run)")
            self.pipeline = AudioToImagePipeline.random_init(seed=0, device=self.device, dtype=self.dtype)
        else:
            self.pipeline = AudioToImagePipeline(unet_sd, vae_sd, device=self.device, dtype=self.dtype)
        ck = self._load("unet_adapter_final.pth")
        if ck is not None:
            if "hierarchical_state_dict" in ck:
                self.pipeline.hier.load_state_dict({k: v.to(self.device) for k, v in ck["hierarchical_state_dict"].items()})
            for lvl, names in self.pipeline.manager.level_mapping.items():
                if names and f"processor_{lvl}" in ck:
                    self.pipeline.unet.sites[names[0][:-len(".processor")]].processor.load_state_dict(ck[f"processor_{lvl}"])

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
            clap, torch.from_numpy(synthetic.text_states(text_prompt)[None]).to(self.device, self.dtype),
            torch.from_numpy(synthetic.text_states("")[None]).to(self.device, self.dtype),
            torch.from_numpy(synthetic.init_noise(noise_seed)[None]).to(self.device),
            steps=num_inference_steps, guidance=guidance_scale, use_audio=use_hierarchical, decode=True)
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
